"""Regenerates profiles/README.md from profiles/bench_r2_1gpu.json + profiles/ncu_r2_launches.csv
(keeps the hand-written sections from "## Experiments" on)."""
import collections, csv, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
d = json.load(open(os.path.join(P, "bench_r2_1gpu.json")))
lines = [l for l in open(os.path.join(P, "ncu_r2_launches.csv")) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u in ("nsecond", "ns") else v * 1e3 if u in ("msecond", "ms") else v
    agg.setdefault(row["Kernel Name"], []).append(v)
tot = sum(sum(v) for k, v in agg.items() if "wm::" in k)
old = open(os.path.join(P, "README.md")).read() if os.path.exists(os.path.join(P, "README.md")) else ""
tail = old[old.index("## Scaling"):] if "## Scaling" in old else ""
out = ["# Round-2 measurements (B200, 64x3x512x512 fp32, BASELINE config 2: 7 layers, 14 kernels per step)", "",
       "Source files: `bench_r2_1gpu.json` (python bench.py --steps 20 --warmup 5), `bench_r2_2gpu.json` (torchrun, 2 ranks),",
       "`ncu_r2_launches.csv` (ncu --metrics gpu__time_duration.sum --clock-control none, same command),",
       "`ncu_r2_*.txt` (ncu --set full summaries of the kernels changed in round 2; `ncu_r1_*.txt` for the others),",
       "`ncu_traffic.json` (DRAM bytes per launch).  All produced by `tools/gpu_profile_r2.sh`.  Round-1 records are kept",
       "(`bench_r1_*.json`, `ncu_r1_*`, `sweep_r1.md`).", "",
       f"Step = {d['ms_per_step']} ms, value = {d['value']} Mpix/s on 1 GPU, whole-step roofline fraction {d['roofline']['whole_step']['frac']}; "
       f"e2e (pinned fp32 host batch in AND 201 MB result out, inside the timed region) = {d['e2e']['value']} Mpix/s, e2e_u8 (8-bit host frames) = {d['e2e_u8']['value']} Mpix/s; "
       f"the unmodified reference on the host cores = {d['cpu_baseline']['value']} Mpix/s on {d['cpu_baseline']['cores']} threads, run eagerly on the same B200 = "
       f"{d['reference_gpu_eager']['value']} Mpix/s.", "",
       "## Per layer, CUDA-event time of the per-kernel pass (the same K steps right after the timed region, an event after every kernel; fraction of the measured 6536.7 GB/s copy peak on ALGORITHMIC bytes).  The timed region itself brackets only the dominant kernel: " + f"{d['roofline']['kernel']} = {d['roofline']['kernel_ms'] * 1e3:.1f} us live, {d['roofline']['frac']} of the roofline; the per-kernel pass runs at {d['kernels_pass']['ms_per_step']} ms per step (every event costs ~3 us)", "",
       "| layer.direction | us | alg. GB/s | frac |", "|---|---|---|---|"]
for k, v in d["kernels"].items():
    out.append(f"| {k} | {v['ms'] * 1e3:.1f} | {v['GBps']:.0f} | {v['frac']:.3f} |")
out += ["", "## ncu launch list (cold-cache, serialised; compare SHARES): average duration and share of the wm:: kernels", "",
        "| kernel | launches | avg us | share |", "|---|---|---|---|"]
for k, v in agg.items():
    if "wm::" in k:
        out.append(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / tot * 100:.1f}% |")
out += ["", "Notes: ncu flushes caches around each kernel, so the write-back of the last ~126 MB of a kernel's output (L2 size) falls "
        "outside its ncu duration; the event times above include it (steady state). DiffJPEG forward here is the state-saving variant "
        "(31 B/px of real traffic against 24 B/px algorithmic), its backward reads 31 B/px against 36 algorithmic.", ""]
open(os.path.join(P, "README.md"), "w").write("\n".join(out) + "\n" + tail)
print("\n".join(out[8:24]))

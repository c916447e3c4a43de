"""Summarise an .ncu-rep: headline metrics + executed-instruction share per basic block.
    python tools/ncu_blocks.py gpurun_out/prof.ncu-rep [min_share_pct]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
for ln in det.splitlines():
    if any(k in ln for k in ("Duration", "DRAM Throughput", "Issue Slots Busy", "Executed Ipc Active", "Eligible Warps", "Active Warps Per",
                             "Registers Per", "Theoretical Occ", "Achieved Occ", "Warp Cycles Per Issued", "Dynamic Shared", "Block Limit Sh", "Block Limit Reg", "wm::")):
        print(ln.rstrip()[:150])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
for row in r[2:]:
    for k, v in zip(r[0], row):
        if (("issue_stalled" in k and "per_issue_active" in k and float(v or 0) > 0.15) or k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "Kernel Name")
                or "data_bank_conflicts_pipe_lsu_mem_shared.sum" in k):
            print("  ", k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), v)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for blk in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    print("==", rows[0][1][:100])
    hdr = rows[1]; data = [x for x in rows[2:] if len(x) == len(hdr)]
    isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); ist = hdr.index("Warp Stall Sampling (All Samples)")
    ex = [int(x[iex]) if x[iex].isdigit() else 0 for x in data]
    st = [int(x[ist]) if x[ist].isdigit() else 0 for x in data]
    tot = sum(ex) or 1; tst = sum(st) or 1
    print(f"   total warp-instr {tot}  sass {len(data)}")
    i = 0
    while i < len(data):
        j = i
        while j + 1 < len(data) and ex[j + 1] == ex[i]:
            j += 1
        n = j - i + 1
        if ex[i] * n > tot * thr / 100:
            ops = []
            for k in range(i, j + 1):
                t = data[k][isrc].split()
                ops.append((t[1] if t[0].startswith("@") else t[0]).split(".")[0])
            print(f"   {ex[i] * n / tot * 100:5.1f}% instr {sum(st[i:j + 1]) / tst * 100:5.1f}% stall  n={n:3d} x{ex[i]:>9d} [{i}-{j}] {dict(Counter(ops))}")
        i = j + 1

#!/usr/bin/env python
"""bench.py — attack-layer hot path throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: DiffJPEG(q=50) forward+backward followed
by forward+backward of every member of BASELINE config 2's
Combined([JpegCompression, GaussianBlur(k3), MiddleBlur(3), Gaussian(σ=.05), Resize(bicubic)])
on a 64x3x512x512 fp32 batch (Combined picks ONE member at random per call; running all five
by id is its expectation and keeps steps identical).  Metric: megapixels of (b,h,w) locations
pushed through a layer's forward+backward per second = steps * 6 * B*H*W / time.

  value     : device-resident inputs, CUDA-event timed, max over ranks.
  e2e       : same step through the public nn.Module API with the batch in PINNED HOST memory:
              every step copies its input host->device (double-buffered on a copy stream) and
              reads a per-step result scalar back.
  roofline  : for the kernel with the largest share of the step: algorithmic bytes / measured
              average launch duration (CUDA events inside the timed region) vs the measured HBM peak.
  cpu_baseline / --impl reference : the CPU oracle port (oracle/attack_oracle.py, torch fp32,
              all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "video-watermarking-forgery-detection_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

B, H, W = 64, 512, 512
QUALITY = 50
RESIZE_RATIOS = (0.5, 0.75, 1.25, 1.5)
LAYERS = ("diffjpeg", "jpegcompression", "gaussianblur", "middleblur3", "gaussian", "resize")
# algorithmic HBM bytes per (b,h,w) location, fp32 NCHW (DESIGN.md §4 / SURVEY §8d)
ALG_BYTES = {
    "diffjpeg": (24, 36), "jpegcompression": (24, 24), "gaussianblur": (24, 24),
    "middleblur3": (27, 27), "gaussian": (24, 24), "resize": (24, 36),   # gaussian: SURVEY 8d (1-bit clamp mask saved)
}
METRIC = "DiffJPEG+Combined fwd+bwd Mpix/s"
WORKLOAD = ("configs[1]: DiffJPEG(q50) + Combined([JpegCompression, GaussianBlur(k3), MiddleBlur(3), "
            "Gaussian(.05), Resize(bicubic)]) members id=0..4, fwd+bwd, 64x3x512x512 fp32")


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- our arm
def build_layers(dev):
    import wmattack
    dj = wmattack.DiffJPEG(True, H, W, quality=QUALITY)
    comb = wmattack.Combined([wmattack.JpegCompression(dev), wmattack.GaussianBlur(), wmattack.MiddleBlur(3),
                              wmattack.Gaussian(), wmattack.Resize()])
    return dj, comb


def run_step(dj, comb, x, g, step, events=None):
    """One step; if `events` is given, records a CUDA event after every forward and backward."""
    def mark():
        if events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events.append(e)
    checksum = None
    mark()
    for li in range(6):
        x.grad = None
        if li == 0:
            y = dj(x)
        elif li == 5:
            y = comb.list[4](x, resize_ratio=RESIZE_RATIOS[step % len(RESIZE_RATIOS)])
        else:
            y = comb(x, id=li - 1)
        mark()
        y.backward(g)
        mark()
        checksum = x.grad if checksum is None else checksum  # keep a handle for the result read-back
    return checksum


def ours(args, rank, world, dev):
    from wmattack import _lib
    # Run backward nodes on the calling thread: every layer is ONE kernel launch of 60-170 us, and the
    # autograd engine's hand-off to its device thread (~40 us per backward() call) would otherwise leave
    # the GPU idle between launches.  A trainer that calls .backward() once per step does not need this.
    torch.autograd.set_multithreading_enabled(False)
    torch.manual_seed(1234 + rank)
    dj, comb = build_layers(dev)
    gen = torch.Generator(dev).manual_seed(rank)
    x = torch.rand(B, 3, H, W, device=dev, generator=gen).requires_grad_(True)
    g = torch.rand(B, 3, H, W, device=dev, generator=gen)
    px_step = 6 * B * H * W

    clk = ClockSampler(torch.cuda.current_device())
    clk.__enter__()                      # samples every 20 ms until the e2e leg is done (all under load)
    for s in range(args.warmup):
        run_step(dj, comb, x, g, s)
    barrier(world)
    torch.cuda.synchronize()
    events = []
    launches0 = _lib.launch_count
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(args.steps):
        run_step(dj, comb, x, g, s, events)
    t1.record()
    torch.cuda.synchronize()
    barrier(world)
    launches = _lib.launch_count - launches0
    ms = t0.elapsed_time(t1)
    ms = allreduce_max(ms, world, dev)
    value = world * args.steps * px_step / (ms / 1e3) / 1e6

    # per-kernel-group durations from the events recorded inside the timed region
    per = {}
    n_ev = 13
    for s in range(args.steps):
        ev = events[s * n_ev:(s + 1) * n_ev]
        for li, name in enumerate(LAYERS):
            per.setdefault(name + ".fwd", []).append(ev[2 * li].elapsed_time(ev[2 * li + 1]))
            per.setdefault(name + ".bwd", []).append(ev[2 * li + 1].elapsed_time(ev[2 * li + 2]))
    avg = {k: sum(v) / len(v) for k, v in per.items()}
    px = B * H * W
    peak, peak_src = hbm_peak()
    kernels = {}
    for k, msk in avg.items():
        name, d = k.split(".")
        byts = ALG_BYTES[name][0 if d == "fwd" else 1] * px
        kernels[k] = {"ms": round(msk, 4), "GBps": round(byts / (msk / 1e3) / 1e9, 1),
                      "frac": round(byts / (msk / 1e3) / 1e9 / peak, 3)}
    # dominant kernel: every layer is ONE kernel launch per direction, so each event interval is one kernel
    single = dict(avg)
    dom = max(single, key=single.get)
    dname, dd = dom.split(".")
    dbytes = ALG_BYTES[dname][0 if dd == "fwd" else 1] * px
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(dbytes / (avg[dom] / 1e3) / 1e9, 1), "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": round(dbytes / (avg[dom] / 1e3) / 1e9 / peak, 4),
                "traffic": load_ncu_traffic(dom), "share_of_step": round(avg[dom] / (ms / args.steps), 4),
                "algorithmic_bytes_per_launch": dbytes}

    e2e = run_e2e(args, dj, comb, g, rank, world, dev, px_step)
    e2e_u8 = run_e2e(args, dj, comb, g, rank, world, dev, px_step, u8=True)
    if len(clk.lines) < 3:               # very short runs: keep the GPU busy until a few samples exist
        t_end = time.time() + 0.5
        while time.time() < t_end:
            run_step(dj, comb, x, g, 0)
        torch.cuda.synchronize()
    clk.__exit__(None, None, None)
    out = {
        "metric": METRIC, "value": round(value, 1), "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": [B, 3, H, W], "resize_ratios": list(RESIZE_RATIOS),
                   "l2": "each tensor is 201 MB > 126 MB L2, no flush needed", "sharding": f"batch x{world}, no collective"},
        "e2e": e2e, "e2e_u8": dict(e2e_u8, note="same step with 8-bit host frames uploaded as bytes and converted on "
                                              "the device (wm_u8_to_unit_float); extra to the contract's fp32 e2e"),
        "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
        "clocks": clk.summary(),
    }
    return out


def bind_to_gpu_numa_node(dev):
    """Pin this rank's host threads (and therefore its pinned staging buffers, first-touch) to the CPUs
    local to its GPU (sysfs local_cpulist of the PCI device), so that with 8 ranks the H2D copies do
    not all cross the socket interconnect.  Best effort: returns the cpu list or None."""
    try:
        p = torch.cuda.get_device_properties(dev)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:
        pass
    return None


def run_e2e(args, dj, comb, g, rank, world, dev, px_step, u8=False):
    """Host-resident input: H2D of the step's batch (pinned, double-buffered on a copy stream) +
    D2H of a result scalar, every step, inside the timed region.
    u8=True: the host holds 8-bit frames (what a video decoder produces); they are uploaded as bytes
    and converted to [0,1] float on the device by wmattack.functional.from_uint8 inside the timed region."""
    from wmattack import functional as WF
    numa = bind_to_gpu_numa_node(dev) if world > 1 else None
    if u8:
        host = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
        stage = [torch.empty(B, 3, H, W, device=dev, dtype=torch.uint8) for _ in range(2)]
    else:
        host = [torch.rand(B, 3, H, W).pin_memory() for _ in range(2)]
    devbuf = [torch.empty(B, 3, H, W, device=dev) for _ in range(2)]
    result = torch.zeros(1).pin_memory()
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[i % 2])
            if u8:
                stage[i % 2].copy_(host[i % 2], non_blocking=True)
                WF.from_uint8(stage[i % 2], out=devbuf[i % 2])          # on the copy stream, ahead of the step
            else:
                devbuf[i % 2].copy_(host[i % 2], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def loop(n):
        for e in free:
            e.record(main)
        upload(0)
        for s in range(n):
            if s + 1 < n:
                upload(s + 1)
            main.wait_event(ready[s % 2])
            x = devbuf[s % 2].requires_grad_(True)
            gx = run_step(dj, comb, x, g, s)
            result.copy_(gx.view(-1)[:1], non_blocking=True)
            devbuf[s % 2] = x.detach()
            free[s % 2].record(main)
        torch.cuda.synchronize()

    loop(max(2, args.warmup // 2))
    barrier(world)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    loop(args.steps)
    t1.record()
    torch.cuda.synchronize()
    ms = allreduce_max(t0.elapsed_time(t1), world, dev)
    return {"value": round(world * args.steps * px_step / (ms / 1e3) / 1e6, 1), "unit": "Mpix/s",
            "h2d_bytes_per_step": B * 3 * H * W * (1 if u8 else 4), "d2h_bytes_per_step": 4,
            "ms_per_step": round(ms / args.steps, 4), "host_cpus": numa}


def load_ncu_traffic(kernel_key):
    """dram bytes/launch of the dominant kernel from the committed ncu summary (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


# --------------------------------------------------------------------------------- CPU oracle arm
def cpu_port_step(O, x, g, step):
    """The same 6 forward+backward passes on the CPU oracle (torch fp32, autograd)."""
    fns = [lambda t: O.diffjpeg(t, QUALITY), O.jpeg_compression, lambda t: O.gaussian_blur(t, 3),
           lambda t: O.median_blur(t, 3),
           lambda t: O.gaussian_noise_clamped(t, torch.randn_like(t) * 0.05),
           lambda t: O.resize(t, RESIZE_RATIOS[step % len(RESIZE_RATIOS)])]
    for fn in fns:
        xx = x.clone().requires_grad_(True)
        fn(xx).backward(g)


def cpu_baseline(steps=1, sample_b=2, warmup=1):
    from oracle import attack_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.rand(sample_b, 3, H, W)
    g = torch.rand(sample_b, 3, H, W)
    for s in range(warmup):
        cpu_port_step(O, x, g, s)
    t = time.perf_counter()
    for s in range(steps):
        cpu_port_step(O, x, g, s)
    dt = time.perf_counter() - t
    val = steps * 6 * sample_b * H * W / dt / 1e6
    return {"value": round(val, 3), "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} step(s) of the same 6 fwd+bwd passes on a {sample_b}x3x{H}x{W} batch "
                      f"(oracle/attack_oracle.py, torch {torch.__version__} CPU fp32, {dt:.1f} s)"}, dt / steps


def reference_arm(args, rank, world):
    if rank != 0:
        return None
    sample_b = 16
    base, sec = cpu_baseline(steps=max(1, args.steps), sample_b=sample_b, warmup=min(args.warmup, 1))
    return {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": [B, 3, H, W],
                       "note": "reference is pure Python/PyTorch and cannot travel to the GPU box; its CPU path "
                               "is timed through the oracle port on a bounded sample"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ------------------------------------------------------------------------------------- plumbing
def barrier(world):
    if world > 1:
        torch.distributed.barrier()


def allreduce_max(v, world, dev):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        line = reference_arm(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU oracle arm)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            # keep stdout to the one JSON line: "NCCL version ..." goes there when NCCL_DEBUG=VERSION comes
            # from the environment or from an nccl.conf on the box (the environment wins over the file)
            os.environ["NCCL_DEBUG"] = "WARN"
        torch.distributed.init_process_group("nccl", device_id=dev)
    out = ours(args, rank, world, dev)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"], _ = cpu_baseline(steps=8, sample_b=16, warmup=1)
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py — attack-layer hot path throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: DiffJPEG(q=50) forward+backward followed by
forward+backward of every member of BASELINE config 2's
Combined([JpegCompression, GaussianBlur(k3), MiddleBlur(5) and (3), Gaussian(sigma=.05), Resize(bicubic)])
on a 64x3x512x512 fp32 batch (Combined picks ONE member at random per call, running
every member by id is its expectation and keeps steps identical).  Metric: megapixels of (b,h,w)
locations pushed through a layer's forward+backward per second = steps * 7 * B*H*W / time.

  value     : device-resident inputs, CUDA-event timed, max over ranks.
  e2e       : same step through the public nn.Module API with the batch in PINNED HOST memory: every
              step copies its input host->device (201 MB) and the step's result (the last layer's
              input gradient, 201 MB) device->host, both inside the timed region, double-buffered
              on copy streams.
  roofline  : for the kernel with the largest share of the step: algorithmic bytes / measured
              average launch duration (CUDA events inside the timed region) vs the measured HBM peak.
  cpu_baseline / --impl reference : the UNMODIFIED reference modules (baseline/_ref, vendored by
              baseline/vendor_reference.py; kind "reference") on the box's host cores, bounded sample;
              falls back to the oracle port (kind "port") only if baseline/_ref is absent.
  extra keys: kernels (per-kernel roofline table), config1 (DiffJPEG q50 16x3x256^2: reference on
              host cores exactly as SURVEY 8d C1 + ours), reference_gpu_eager (the reference's
              eager graph on this B200, same step), config3 (1080p frame sweep), config5 (4K
              quality sweep, this rank's share of the 64-frame clip), train_step (config 4).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
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
LAYERS = ("diffjpeg", "jpegcompression", "gaussianblur", "middleblur5", "middleblur3", "gaussian", "resize")
NL = len(LAYERS)
# algorithmic HBM bytes per (b,h,w) location, fp32 NCHW (DESIGN.md §3 / SURVEY §8d)
ALG_BYTES = {
    "diffjpeg": (24, 36), "jpegcompression": (24, 24), "gaussianblur": (24, 24), "middleblur5": (27, 27),
    "middleblur3": (27, 27), "gaussian": (24, 24), "resize": (24, 36),
}
METRIC = "DiffJPEG+Combined fwd+bwd Mpix/s"
WORKLOAD = ("configs[1]: DiffJPEG(q50) + Combined([JpegCompression, GaussianBlur(k3), MiddleBlur(5), MiddleBlur(3), "
            "Gaussian(.05), Resize(bicubic)]) every member by id, fwd+bwd, 64x3x512x512 fp32")


def config_block(world):
    return {"workload": WORKLOAD, "per_gpu_batch": [B, 3, H, W], "resize_ratios": list(RESIZE_RATIOS),
            "l2": "each tensor is 201 MB > 126 MB L2, no flush needed", "sharding": f"batch x{world}, no collective"}


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


def timed(fn, iters, warm, world=1, dev=None):
    """ms per call of fn(): CUDA events on the current stream, `warm` untimed calls, max over ranks."""
    for _ in range(warm):
        fn()
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return allreduce_max(e0.elapsed_time(e1) / iters, world, dev)


# ------------------------------------------------------------------------------------- our arm
def build_layers(dev):
    import wmattack
    dj = wmattack.DiffJPEG(True, H, W, quality=QUALITY)
    comb = wmattack.Combined([wmattack.JpegCompression(dev), wmattack.GaussianBlur(), wmattack.MiddleBlur(5),
                              wmattack.MiddleBlur(3), wmattack.Gaussian(), wmattack.Resize()])
    return dj, comb


def run_step(dj, comb, x, g, step, events=None, only=None):
    """One step; if `events` is given, records a CUDA event after every forward and backward (mark m: 0 = step start,
    2*li+1 = after layer li's forward, 2*li+2 = after its backward) - or, with `only`, just the marks in that set.
    Returns the last layer's input gradient (the step's result for the e2e read-back)."""
    m = [0]

    def mark():
        if events is not None and (only is None or m[0] in only):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events.append(e)
        m[0] += 1
    mark()
    for li in range(NL):
        x.grad = None
        if li == 0:
            y = dj(x)
        elif li == NL - 1:
            y = comb.list[5](x, resize_ratio=RESIZE_RATIOS[step % len(RESIZE_RATIOS)])
        else:
            y = comb(x, id=li - 1)
        mark()
        y.backward(g)
        mark()
    return x.grad


def ours(args, rank, world, dev):
    from wmattack import _lib
    # Run backward nodes on the calling thread: every layer is ONE kernel launch of 60-300 us, and the
    # autograd engine's hand-off to its device thread (~40 us per backward() call) would otherwise leave
    # the GPU idle between launches.  A trainer that calls .backward() once per step does not need this.
    torch.autograd.set_multithreading_enabled(False)
    torch.manual_seed(1234 + rank)
    dj, comb = build_layers(dev)
    gen = torch.Generator(dev).manual_seed(rank)
    x = torch.rand(B, 3, H, W, device=dev, generator=gen).requires_grad_(True)
    g = torch.rand(B, 3, H, W, device=dev, generator=gen)
    px_step = NL * B * H * W

    clk = ClockSampler(torch.cuda.current_device())
    clk.__enter__()                      # samples every 20 ms until the last GPU leg is done (all under load)
    with torch.no_grad():                # set-up, not a step: the Resize band tables of every geometry the steps cycle through
        for r in RESIZE_RATIOS:          # (cached per geometry; a W < 4 warm-up would otherwise build one inside the timed region)
            comb.list[5](x.detach(), resize_ratio=r)
    for s in range(args.warmup):
        run_step(dj, comb, x, g, s)
    barrier(world)
    torch.cuda.synchronize()
    # The timed region brackets ONE kernel with events - the dominant one (the 5x5 median forward), whose live duration the
    # roofline object needs.  An event between EVERY pair of kernels costs ~3 us of pipeline bubble each (measured:
    # 1.487 vs 1.446 ms per step, tools/exp/event_gap_probe.py), so the full per-kernel table comes from a second pass of
    # the same K steps run right after the timed region (`kernels_pass` in the JSON line says so).
    DOM = "middleblur5.fwd"
    dli = LAYERS.index(DOM.split(".")[0])
    dom_marks = {2 * dli, 2 * dli + 1}
    dom_events = []
    launches0 = _lib.launch_count
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(args.steps):
        run_step(dj, comb, x, g, s, dom_events, dom_marks)
    t1.record()
    torch.cuda.synchronize()
    barrier(world)
    launches = _lib.launch_count - launches0
    ms = t0.elapsed_time(t1)
    ms = allreduce_max(ms, world, dev)
    value = world * args.steps * px_step / (ms / 1e3) / 1e6
    dom_ms = sum(dom_events[2 * s].elapsed_time(dom_events[2 * s + 1]) for s in range(args.steps)) / args.steps

    # second pass: the same K steps with an event after every forward and backward (every layer is ONE kernel launch per
    # direction, so each event interval is one kernel + its launch gap)
    events = []
    p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
    p0.record()
    for s in range(args.steps):
        run_step(dj, comb, x, g, s, events)
    p1.record()
    torch.cuda.synchronize()
    pass_ms = p0.elapsed_time(p1) / args.steps
    per = {}
    n_ev = 2 * NL + 1
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
    dom = max(avg, key=avg.get)
    # the kernel bracketed inside the timed region is the dominant one of the second pass too (if it ever were not,
    # the roofline object falls back to the second pass's interval and says so)
    live = dom == DOM
    dom_live_ms = dom_ms if live else avg[dom]
    dname, dd = dom.split(".")
    dbytes = ALG_BYTES[dname][0 if dd == "fwd" else 1] * px
    step_bytes = sum(sum(ALG_BYTES[n]) for n in LAYERS) * px
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(dbytes / (dom_live_ms / 1e3) / 1e9, 1), "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": round(dbytes / (dom_live_ms / 1e3) / 1e9 / peak, 4),
                "traffic": load_ncu_traffic(dom), "share_of_step": round(dom_live_ms / (ms / args.steps), 4),
                "algorithmic_bytes_per_launch": dbytes, "kernel_ms": round(dom_live_ms, 4),
                "measured": "CUDA events around this kernel inside the timed region" if live else "second pass (see kernels_pass)",
                "whole_step": {"algorithmic_bytes": step_bytes, "GBps": round(step_bytes / (ms / args.steps / 1e3) / 1e9, 1),
                               "frac": round(step_bytes / (ms / args.steps / 1e3) / 1e9 / peak, 4)}}
    kernels_pass = {"what": "the same K steps run once more right after the timed region with a CUDA event after EVERY forward and "
                            "backward: the `kernels` table.  The timed region itself records events only around the dominant kernel "
                            "(an event between every pair of kernels costs ~3 us of pipeline bubble each)",
                    "ms_per_step": round(pass_ms, 4)}

    e2e = run_e2e(args, dj, comb, g, rank, world, dev, px_step)
    e2e_u8 = run_e2e(args, dj, comb, g, rank, world, dev, px_step, u8=True)
    extras = {}
    for name, fn in (("config2_natural", lambda: config2_natural(args, dj, comb, g, world, dev, px_step)),
                     ("config2_autocast", lambda: config2_autocast(args, dj, comb, g, world, dev, px_step)),
                     ("config1", lambda: config1_gpu(dev)), ("config3", lambda: config3(world, dev)),
                     ("config5", lambda: config5(rank, world, dev)),
                     ("train_step", lambda: train_step_leg(args, rank, world, dev))):
        if name in args.skip:
            continue
        try:
            extras[name] = fn()
        except Exception as e:                      # an extra leg must never take the headline down
            extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    if world == 1 and "reference_gpu_eager" not in args.skip:
        try:
            extras["reference_gpu_eager"] = reference_gpu_eager(dev, value)
        except Exception as e:
            extras["reference_gpu_eager"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
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
        "config": config_block(world),
        "e2e": e2e, "e2e_u8": dict(e2e_u8, note="same step with 8-bit host frames: bytes uploaded and converted on the "
                                              "device (wm_u8_to_unit_float), result returned as bytes; extra to the contract's fp32 e2e"),
        "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "kernels_pass": kernels_pass,
        "parity_note": "MiddleBlur/GF delegate to kornia upstream (not vendored/pinned/installed): their parity is "
                       "pinned only to the kornia 0.6.x algorithm restated in oracle/ — 'parity unpinned' rows",
        "clocks": clk.summary(),
    }
    out.update(extras)
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
    """Host-resident input AND result: per step, H2D of the batch from pinned memory (copy stream,
    double-buffered, overlapping the previous step's kernels) and D2H of the step's result — the input
    gradient of the last layer, 201 MB — into pinned memory (second copy stream), all inside the timed region.
    u8=True: the host holds 8-bit frames (what a video decoder produces / an encoder consumes): bytes are
    uploaded, converted on the device (wm_u8_to_unit_float), and the attacked frames of the last layer go
    back as bytes (wm_unit_float_to_u8)."""
    from wmattack import functional as WF
    numa = bind_to_gpu_numa_node(dev) if world > 1 else None
    shape = (B, 3, H, W)
    if u8:
        host = [torch.randint(0, 256, shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        stage = [torch.empty(shape, device=dev, dtype=torch.uint8) for _ in range(2)]
        res_dev = [torch.empty(shape, device=dev, dtype=torch.uint8) for _ in range(2)]
        res_host = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
    else:
        host = [torch.rand(shape).pin_memory() for _ in range(2)]
        res_dev = [None, None]
        res_host = [torch.empty(shape).pin_memory() for _ in range(2)]
    devbuf = [torch.empty(shape, device=dev) for _ in range(2)]
    up_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
    main = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]       # upload i landed
    free = [torch.cuda.Event() for _ in range(2)]        # devbuf i no longer read by the step
    done = [torch.cuda.Event() for _ in range(2)]        # step result i produced
    drained = [torch.cuda.Event() for _ in range(2)]     # D2H of result i finished

    def upload(i):
        with torch.cuda.stream(up_stream):
            up_stream.wait_event(free[i % 2])
            if u8:
                stage[i % 2].copy_(host[i % 2], non_blocking=True)
                WF.from_uint8(stage[i % 2], out=devbuf[i % 2])          # on the copy stream, ahead of the step
            else:
                devbuf[i % 2].copy_(host[i % 2], non_blocking=True)
            ready[i % 2].record(up_stream)

    def loop(n):
        for e in free + drained:
            e.record(main)
        upload(0)
        for s in range(n):
            if s + 1 < n:
                upload(s + 1)
            main.wait_event(ready[s % 2])
            x = devbuf[s % 2].requires_grad_(True)
            gx = run_step(dj, comb, x, g, s)
            if u8:
                main.wait_event(drained[s % 2])                           # res_dev slot reusable
                with torch.no_grad():
                    WF.to_uint8(comb.list[5](x.detach(), resize_ratio=RESIZE_RATIOS[s % 4]), out=res_dev[s % 2])
                res = res_dev[s % 2]
            else:
                res = gx
            done[s % 2].record(main)
            with torch.cuda.stream(down_stream):
                down_stream.wait_event(done[s % 2])
                res_host[s % 2].copy_(res, non_blocking=True)
                res.record_stream(down_stream)
                drained[s % 2].record(down_stream)
            devbuf[s % 2] = x.detach()
            x.grad = None
            free[s % 2].record(main)
        torch.cuda.synchronize()

    loop(max(2, args.warmup // 2))
    barrier(world)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    loop(args.steps)
    torch.cuda.synchronize()
    t1.record()
    torch.cuda.synchronize()
    ms = allreduce_max(t0.elapsed_time(t1), world, dev)
    nbytes = B * 3 * H * W * (1 if u8 else 4)
    per_step = ms / args.steps
    # the link itself: the same two buffers copied up and down at once with NO kernels in between (all ranks together),
    # so the distance of the step above from the PCIe / host-memory ceiling is explicit
    src_h, dst_d = host[0], (stage[0] if u8 else devbuf[0])
    src_d, dst_h = (res_dev[0] if u8 else devbuf[1]), res_host[0]

    def raw_pair():
        with torch.cuda.stream(up_stream):
            dst_d.copy_(src_h, non_blocking=True)
        with torch.cuda.stream(down_stream):
            dst_h.copy_(src_d, non_blocking=True)
    raw_pair()
    torch.cuda.synchronize()
    barrier(world)
    l0 = torch.cuda.Event(enable_timing=True); l1 = torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(5):
        raw_pair()
    main.wait_stream(up_stream); main.wait_stream(down_stream)
    l1.record()
    torch.cuda.synchronize()
    link_ms = allreduce_max(l0.elapsed_time(l1) / 5, world, dev)
    return {"value": round(world * args.steps * px_step / (ms / 1e3) / 1e6, 1), "unit": "Mpix/s",
            "link_probe": {"ms_per_up_down_pair": round(link_ms, 4), "aggregate_GBps_each_way": round(world * nbytes / (link_ms / 1e3) / 1e9, 1),
                           "step_over_link": round(per_step / link_ms, 3),
                           "note": "the same buffers copied H2D and D2H concurrently with no kernels: the link ceiling of this step"},
            "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
            "ms_per_step": round(per_step, 4), "host_cpus": numa,
            "aggregate_h2d_GBps": round(world * nbytes / (per_step / 1e3) / 1e9, 1),
            "aggregate_d2h_GBps": round(world * nbytes / (per_step / 1e3) / 1e9, 1),
            "bound": "PCIe: the copies, not a kernel, set this number (see aggregate_*_GBps)"}


def load_ncu_traffic(kernel_key):
    """dram bytes/launch of the dominant kernel from the committed ncu summary (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


# ----------------------------------------------------------------------- extra legs (configs 1, 3, 5)
def config2_natural(args, dj, comb, g, world, dev, px_step):
    """The headline step on NATURAL-LIKE frames (SURVEY 8(d) "value distributions"): low-pass filtered noise quantised to
    k/255 — post-Quantization statistics, where the median windows tie, most DCT coefficients fall in the zero bin and
    the clamps of Resize / Gaussian noise rarely fire.  The kernels are data-oblivious except for rare-path branches."""
    gen = torch.Generator(dev).manual_seed(7)
    x = torch.rand(B, 3, H, W, device=dev, generator=gen)
    k = torch.ones(3, 1, 9, 9, device=dev) / 81.0
    for _ in range(2):                                       # two 9x9 box blurs: smooth, image-like spectrum
        x = torch.nn.functional.conv2d(x, k, padding=4, groups=3)
    x = (x - x.amin()) / (x.amax() - x.amin())
    x = (torch.round(x * 255) / 255).contiguous().requires_grad_(True)
    for s in range(max(2, args.warmup // 2)):
        run_step(dj, comb, x, g, s)
    torch.cuda.synchronize()
    ms = timed(lambda: [run_step(dj, comb, x, g, s) for s in range(4)], max(1, args.steps // 4), 0, world, dev) / 4
    ties = float((wmattack_unique_fraction(x.detach()[:2])))
    return {"workload": "the config-2 step on smooth frames quantised to k/255 (two 9x9 box blurs of uniform noise)",
            "ms_per_step": round(ms, 4), "value": round(world * px_step / ms / 1e3, 1), "unit": "Mpix/s",
            "fraction_of_3x3_windows_with_a_repeated_value": round(ties, 3)}


def config2_autocast(args, dj, comb, g, world, dev, px_step):
    """The headline step on a bfloat16 batch - the autocast boundary of the trainers (models/IRNcrop_model.py:340): every
    layer stages / reads the 2-byte image as it is and stores its input gradient in that type (the *_typed entry points
    and the typed DiffJPEG / 8x8 JPEG / noise kernels); the attack arithmetic stays float32.  An extra, not the headline:
    the headline is the reference's float32 configuration."""
    gen = torch.Generator(dev).manual_seed(11)
    x = torch.rand(B, 3, H, W, device=dev, generator=gen).to(torch.bfloat16).requires_grad_(True)
    for s in range(max(4, args.warmup)):                     # every Resize ratio once: the 2-byte gradient buffers of each size get cached
        run_step(dj, comb, x, g, s)
    torch.cuda.synchronize()
    assert x.grad is not None and x.grad.dtype == torch.bfloat16
    ms = timed(lambda: [run_step(dj, comb, x, g, s) for s in range(4)], max(1, args.steps // 4), 0, world, dev) / 4
    return {"workload": "the config-2 step on a bfloat16 batch (typed boundary: 2-byte image in, 2-byte gradient out, float32 arithmetic)",
            "ms_per_step": round(ms, 4), "value": round(world * px_step / ms / 1e3, 1), "unit": "Mpix/s"}


def wmattack_unique_fraction(x):
    """Fraction of 3x3 windows that hold a repeated value (what makes a median window tie)."""
    u = torch.nn.functional.unfold(x.reshape(-1, 1, *x.shape[2:]), 3, padding=1)          # [N, 9, L]
    srt = u.sort(dim=1).values
    return (srt[:, 1:] == srt[:, :-1]).any(dim=1).float().mean()


def config1_gpu(dev):
    """BASELINE config 1's workload (DiffJPEG q50 fwd+bwd, 16x3x256x256) on this GPU, eager and as a
    replayed CUDA graph (at this size the call is launch-bound)."""
    import wmattack
    g0 = torch.Generator().manual_seed(0)
    x = torch.rand(16, 3, 256, 256, generator=g0).to(dev).requires_grad_(True)
    gy = torch.rand(16, 3, 256, 256, generator=torch.Generator().manual_seed(1)).to(dev)
    m = wmattack.DiffJPEG(True, 256, 256, quality=50)

    def step():
        torch.autograd.grad(m(x), x, gy)
    eager = timed(step, 50, 5)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(); step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    rep = timed(graph.replay, 200, 10)
    px = 16 * 256 * 256
    return {"workload": "configs[0]: DiffJPEG(q50) fwd+bwd 16x3x256x256 fp32", "ours_eager_ms": round(eager, 4),
            "ours_graph_ms": round(rep, 4), "ours_eager_mpix_s": round(px / eager / 1e3, 1),
            "ours_graph_mpix_s": round(px / rep / 1e3, 1)}


def config3(world, dev, frames=(1, 4, 16, 64, 256)):
    """BASELINE config 3: MiddleBlur 3x3 / 5x5 and GaussianBlur sigma=2 (k3 = the layer's default, k7 = GF's)
    over N x 3 x 1080 x 1920 frame batches, N frames PER GPU (weak), forward + backward."""
    import wmattack
    peak, _ = hbm_peak()
    layers = [("middleblur3", lambda: wmattack.MiddleBlur(3), 54), ("middleblur5", lambda: wmattack.MiddleBlur(5), 54),
              ("gaussianblur_k3", lambda: wmattack.GaussianBlur(3), 48), ("gaussianblur_k7", lambda: wmattack.GaussianBlur(7), 48)]
    out = {"workload": "configs[2]: N x 3 x 1080 x 1920 fp32 per GPU, fwd+bwd", "rows": []}
    for n in frames:
        x = torch.rand(n, 3, 1080, 1920, device=dev).requires_grad_(True)
        g = torch.rand(n, 3, 1080, 1920, device=dev)
        px = n * 1080 * 1920
        for name, mk, bpp in layers:
            layer = mk()
            iters = max(3, min(100, int(3e9 / (px * bpp))))
            ms = timed(lambda: torch.autograd.grad(layer(x), x, g), iters, 3, world, dev)
            gbs = px * bpp / ms / 1e6
            out["rows"].append({"frames_per_gpu": n, "layer": name, "ms": round(ms, 4),
                                "mpix_s_aggregate": round(world * px / ms / 1e3, 1), "frac_of_hbm_peak": round(gbs / peak, 3)})
        del x, g
        torch.cuda.empty_cache()
    return out


def config5(rank, world, dev, total_frames=64, chunk=8):
    """BASELINE config 5: DiffJPEG quality sweep 10..95 over a 64-frame 4K clip, frames sharded
    contiguously over the ranks (strong scaling: 64/world frames each, processed 8 frames per call)."""
    import wmattack
    from wmattack.sharding import frame_shard
    peak, _ = hbm_peak()
    a, b = frame_shard(total_frames, rank, world)
    mine = b - a
    chunk = min(chunk, mine)
    x = torch.rand(chunk, 3, 2160, 3840, device=dev).requires_grad_(True)
    g = torch.rand(chunk, 3, 2160, 3840, device=dev)
    calls = -(-mine // chunk)
    rows = []
    tot_ms = 0.0
    for q in (10, 20, 30, 40, 50, 60, 70, 80, 90, 95):
        m = wmattack.DiffJPEG(True, 2160, 3840, quality=q)

        def clip():
            for _ in range(calls):
                torch.autograd.grad(m(x), x, g)
        ms = timed(clip, 3, 1, world, dev)
        tot_ms += ms
        px = total_frames * 2160 * 3840
        rows.append({"quality": q, "ms_per_clip": round(ms, 3), "mpix_s_aggregate": round(px / ms / 1e3, 1),
                     "frac_of_hbm_peak_per_gpu": round(px / world * 60 / ms / 1e6 / peak, 3)})
    return {"workload": f"configs[4]: DiffJPEG q10..95 fwd+bwd on 64x3x2160x3840, {mine} frames on this rank "
                        f"({chunk} per call), scaling strong", "rows": rows,
            "mpix_s_aggregate_mean": round(10 * total_frames * 2160 * 3840 / tot_ms / 1e3, 1)}


# ------------------------------------------------------------------ the reference itself (baseline/_ref)
def _ref_step_fns(R, dev, b):
    """The same 7 forward+backward passes through the reference's OWN modules (baseline/_ref).
    Combined itself cannot hold MiddleBlur upstream (it has no .name, combined.py:19), so members are
    called directly.  JpegCompression's autograd backward raises on torch >= 2 (in-place unsqueeze_ on
    views, jpeg_compression.py:109,118): its backward is stood in for by a second forward (the layer is
    linear and its adjoint is the same conv pair transposed) — never slower than a real backward."""
    dj = R.DiffJPEG(True, H, W, quality=QUALITY)
    jc, gb, m5, m3, ga, rs = R.JpegCompression(dev), R.GaussianBlur(), R.MiddleBlur(5), R.MiddleBlur(3), R.Gaussian(), R.Resize()
    # utils/JPEG.py keeps its quantisation tables as MODULE-LEVEL nn.Parameters shared by every instance:
    # .to() moves them in place, so always move explicitly (the CPU leg may run after the GPU-eager leg)
    for m in (dj, jc, gb, m5, m3, ga, rs):
        m.to(dev)

    def fb(layer):
        def run(x, g, step):
            xx = x.detach().requires_grad_(True)
            layer(xx).backward(g)
        return run

    def jc_run(x, g, step):
        with torch.no_grad():
            jc(x)
            jc(x)

    def rs_run(x, g, step):
        xx = x.detach().requires_grad_(True)
        rs(xx, resize_ratio=RESIZE_RATIOS[step % len(RESIZE_RATIOS)]).backward(g)
    return [("diffjpeg", fb(dj)), ("jpegcompression", jc_run), ("gaussianblur", fb(gb)), ("middleblur5", fb(m5)),
            ("middleblur3", fb(m3)), ("gaussian", fb(ga)), ("resize", rs_run)]


def _port_step_fns():
    from oracle import attack_oracle as O
    def fb(fn):
        def run(x, g, step):
            xx = x.detach().requires_grad_(True)
            fn(xx).backward(g)
        return run
    return [("diffjpeg", fb(lambda t: O.diffjpeg(t, QUALITY))), ("jpegcompression", fb(O.jpeg_compression)),
            ("gaussianblur", fb(lambda t: O.gaussian_blur(t, 3))), ("middleblur5", fb(lambda t: O.median_blur(t, 5))),
            ("middleblur3", fb(lambda t: O.median_blur(t, 3))),
            ("gaussian", fb(lambda t: O.gaussian_noise_clamped(t, torch.randn_like(t) * 0.05))),
            ("resize", lambda x, g, s: fb(lambda t: O.resize(t, RESIZE_RATIOS[s % 4]))(x, g, s))]


def cpu_reference(steps, warmup, sample_b):
    """Times the reference's CPU path on the host cores on a bounded sample (sample_b of the 64 images).
    Returns (cpu_baseline dict, seconds per step)."""
    from baseline import ref_harness as RH
    torch.set_num_threads(os.cpu_count() or 1)
    kind = "reference" if RH.available() else "port"
    x = torch.rand(sample_b, 3, H, W)
    g = torch.rand(sample_b, 3, H, W)
    if kind == "reference":
        with RH.cpu_mode():
            fns = _ref_step_fns(RH.layers("cpu"), "cpu", sample_b)
            sec, per = _time_cpu(fns, x, g, steps, warmup)
        src = "UNMODIFIED reference modules from baseline/_ref (utils/JPEG.py DiffJPEG, noise_layers/*)"
        if RH.kornia_is_stub():
            src += "; kornia.filters.MedianBlur = the harness's kornia-0.6.x stand-in (kornia not installed)"
        src += "; JpegCompression: 2 forwards (its autograd backward raises on torch>=2)"
    else:
        fns = _port_step_fns()
        sec, per = _time_cpu(fns, x, g, steps, warmup)
        src = "oracle/attack_oracle.py port (baseline/_ref not present)"
    val = NL * sample_b * H * W / sec / 1e6
    return {"value": round(val, 3), "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} step(s) of the same {NL} fwd+bwd passes on a {sample_b}x3x{H}x{W} sample of the "
                      f"64-image batch; {src}; torch {torch.__version__} CPU fp32, {sec * steps:.1f} s timed",
            "per_layer_ms": {k: round(v * 1e3, 1) for k, v in per.items()}}, sec


def _time_cpu(fns, x, g, steps, warmup):
    for s in range(warmup):
        for _, fn in fns:
            fn(x, g, s)
    per = {k: 0.0 for k, _ in fns}
    t_all = time.perf_counter()
    for s in range(steps):
        for k, fn in fns:
            t = time.perf_counter()
            fn(x, g, s)
            per[k] += time.perf_counter() - t
    sec = (time.perf_counter() - t_all) / steps
    return sec, {k: v / steps for k, v in per.items()}


def config1_reference_cpu():
    """SURVEY 8(d) C1 exactly as named: x = rand(16,3,256,256) seed 0, gy = rand seed 1,
    DiffJPEG(True,256,256,quality=50) (round_only_at_0), y = m(x); y.backward(gy); 3 warm-up + 10 timed, median."""
    from baseline import ref_harness as RH
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.rand(16, 3, 256, 256, generator=torch.Generator().manual_seed(0))
    gy = torch.rand(16, 3, 256, 256, generator=torch.Generator().manual_seed(1))
    if RH.available():
        with RH.cpu_mode():
            m = RH.layers("cpu").DiffJPEG(True, 256, 256, quality=50)
            kind = "reference"
            ts = _c1_times(m, x, gy)
    else:
        from oracle import attack_oracle as O
        kind = "port"
        ts = _c1_times(lambda t: O.diffjpeg(t, 50), x, gy)
    med = statistics.median(ts)
    return {"workload": "configs[0]: DiffJPEG(q50) fwd+bwd 16x3x256x256 fp32 on CPU, 3 warm-up + 10 timed, median",
            "kind": kind, "cores": torch.get_num_threads(), "cpu_count": os.cpu_count(),
            "reference_cpu_ms": round(med * 1e3, 2), "reference_cpu_mpix_s": round(16 * 256 * 256 / med / 1e6, 2)}


def _c1_times(m, x, gy):
    ts = []
    for i in range(13):
        xx = x.clone().requires_grad_(True)
        t = time.perf_counter()
        y = m(xx)
        y.backward(gy)
        if i >= 3:
            ts.append(time.perf_counter() - t)
    return ts


def reference_gpu_eager(dev, our_value, steps=3, warmup=2):
    """The reference's eager PyTorch graph on THIS B200: same 7 fwd+bwd passes, same
    64x3x512x512 batch, unmodified modules from baseline/_ref moved to the GPU."""
    from baseline import ref_harness as RH
    if not RH.available():
        return {"unavailable": "baseline/_ref not present"}
    fns = _ref_step_fns(RH.layers("cuda"), str(dev), B)
    x = torch.rand(B, 3, H, W, device=dev)
    g = torch.rand(B, 3, H, W, device=dev)
    for s in range(warmup):
        for _, fn in fns:
            fn(x, g, s)
    torch.cuda.synchronize()
    per = {k: 0.0 for k, _ in fns}
    for s in range(steps):
        for k, fn in fns:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(x, g, s)
            e1.record()
            torch.cuda.synchronize()
            per[k] += e0.elapsed_time(e1) / steps
    ms = sum(per.values())
    val = NL * B * H * W / ms / 1e3
    return {"value": round(val, 1), "unit": "Mpix/s", "ms_per_step": round(ms, 3), "steps": steps,
            "per_layer_ms": {k: round(v, 3) for k, v in per.items()}, "ours_over_reference_eager": round(our_value / val, 2),
            "note": "unmodified reference modules (baseline/_ref) run eagerly on this GPU, device-resident inputs, "
                    "CUDA events; JpegCompression = 2 forwards (its backward raises upstream)"
                    + ("; MedianBlur = kornia-0.6.x stand-in" if RH.kornia_is_stub() else "")}


def train_step_leg(args, rank, world, dev):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import train_step as TS
    return TS.bench(rank, world, dev)


def reference_arm(args, rank, world):
    if rank != 0:
        return None
    sample_b = 8
    base, sec = cpu_reference(steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)), sample_b=sample_b)
    out = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "Mpix/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_block(world), "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:
        out["config1"] = config1_reference_cpu()
    except Exception as e:
        out["config1"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip", default="", help="comma list of extra legs to skip: config1,config3,config5,train_step,reference_gpu_eager")
    args = ap.parse_args()
    args.skip = set(filter(None, args.skip.split(",")))
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
                         "(use --impl reference for the CPU reference arm)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    out = ours(args, rank, world, dev)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"], _ = cpu_reference(steps=3, warmup=1, sample_b=8)
            except Exception as e:
                out["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            try:
                out.setdefault("config1", {}).update(config1_reference_cpu())
                c1 = out["config1"]
                if "ours_graph_ms" in c1:
                    c1["ours_over_reference_cpu"] = round(c1["reference_cpu_ms"] / c1["ours_graph_ms"], 1)
            except Exception as e:
                out.setdefault("config1", {})["reference_cpu_error"] = f"{type(e).__name__}: {e}"[:300]
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

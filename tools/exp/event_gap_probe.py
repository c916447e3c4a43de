"""Does recording a CUDA event between every pair of kernels (bench.py's per-kernel table) cost step time?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import bench
dev = torch.device("cuda")
torch.autograd.set_multithreading_enabled(False)
dj, comb = bench.build_layers(dev)
x = torch.rand(bench.B, 3, bench.H, bench.W, device=dev).requires_grad_(True)
g = torch.rand(bench.B, 3, bench.H, bench.W, device=dev)
with torch.no_grad():
    for r in bench.RESIZE_RATIOS:
        comb.list[5](x.detach(), resize_ratio=r)
for s in range(8):
    bench.run_step(dj, comb, x, g, s)
def run(with_events, steps=40):
    torch.cuda.synchronize()
    ev = [] if with_events else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(steps):
        bench.run_step(dj, comb, x, g, s, ev)
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps
for rep in range(3):
    print(f"with per-kernel events {run(True):.4f} ms/step   without {run(False):.4f} ms/step", flush=True)

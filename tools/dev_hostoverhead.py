import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import bench
dev = torch.device("cuda", 0)
dj, comb = bench.build_layers(dev)
x = torch.rand(64, 3, 512, 512, device=dev).requires_grad_(True); g = torch.rand(64, 3, 512, 512, device=dev)
for s in range(5): bench.run_step(dj, comb, x, g, s)
torch.cuda.synchronize()
for label, ev in (("with events", []), ("no events", None)):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
    for s in range(40): bench.run_step(dj, comb, x, g, s, ev)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{label}: cpu enqueue {1e3*(t1-t0)/40:.3f} ms/step, gpu {e0.elapsed_time(e1)/40:.3f} ms/step, wall {1e3*(t2-t0)/40:.3f}")
# per-op CPU cost
import wmattack
from wmattack import functional as WF
xs = torch.rand(1, 3, 64, 64, device=dev).requires_grad_(True); gs = torch.rand(1, 3, 64, 64, device=dev)
m = wmattack.GaussianBlur()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(300):
    y = m(xs)
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(300):
    y = m(xs); y.backward(gs)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"tiny blur: fwd {1e6*(t1-t0)/300:.1f} us/call, fwd+bwd {1e6*(t2-t1)/300:.1f} us/call (host-bound)")

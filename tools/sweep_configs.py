"""BASELINE config 3 and config 5 sweeps on one GPU (developer tool, GPU box):
  config 3: MiddleBlur 3x3 / 5x5 and GaussianBlur sigma=2 (k=3, k=7) on N x 3 x 1080 x 1920, N = 1 .. 256
  config 5: DiffJPEG quality 10 .. 95 on one GPU's 8-frame share of the 64-frame 4K clip
forward + backward through the nn.Module API, CUDA events, Mpix/s and fraction of the HBM roofline
(algorithmic bytes of SURVEY 8d).  Writes a markdown table to stdout.

    python tools/sweep_configs.py > profiles/sweep_r1.md
"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6536.7
dev = "cuda"


def _events(step, n):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def time_fwd_bwd(layer, x, g, n=None, graph=False):
    """ms per forward+backward: eager, and (graph=True) as a captured CUDA graph replayed."""
    x = x.requires_grad_(True)
    px = x.shape[0] * x.shape[2] * x.shape[3]
    n = n or max(5, min(200, int(4e9 / (px * 60))))
    eager = _events(lambda: torch.autograd.grad(layer(x), x, g), n)
    if not graph:
        return eager, None
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            torch.autograd.grad(layer(x), x, g)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        torch.autograd.grad(layer(x), x, g)
    return eager, _events(cg.replay, n)


# bring the GPU to its working clocks before the first (smallest, shortest) measurement
_w = torch.rand(64, 3, 1080, 1920, device=dev)
for _ in range(200):
    _w.mul_(1.0001)
torch.cuda.synchronize()
del _w


print("## Config 3: N x 3 x 1080 x 1920 fp32, forward + backward per call (eager nn.Module API, CUDA events)\n")
print("| N frames | layer | ms fwd+bwd (eager) | ms (CUDA-graph replay) | Mpix/s (best) | alg. GB/s | of measured peak |")
print("|---|---|---|---|---|---|---|")
layers3 = [("MiddleBlur(3)", lambda: wmattack.MiddleBlur(3), 54), ("MiddleBlur(5)", lambda: wmattack.MiddleBlur(5), 54),
           ("GaussianBlur(k3, sigma 2)", lambda: wmattack.GaussianBlur(3), 48), ("GaussianBlur(k7, sigma 2)", lambda: wmattack.GaussianBlur(7), 48)]
for N in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    x = torch.rand(N, 3, 1080, 1920, device=dev)
    g = torch.rand(N, 3, 1080, 1920, device=dev)
    for name, mk, bpp in layers3:
        eager, rep = time_fwd_bwd(mk(), x, g, graph=N <= 16)
        ms = min(eager, rep) if rep else eager
        px = N * 1080 * 1920
        gbs = px * bpp / ms / 1e6
        print(f"| {N} | {name} | {eager:.3f} | {'%.3f' % rep if rep else '-'} | {px / ms / 1e3:.0f} | {gbs:.0f} | {gbs / PEAK * 100:.0f} % |", flush=True)
    del x, g
    torch.cuda.empty_cache()

print("\n## Config 5: DiffJPEG quality sweep, 8 x 3 x 2160 x 3840 fp32 (one GPU's share of the 64-frame 4K clip), forward + backward\n")
print("| quality | ms fwd+bwd | Mpix/s | alg. GB/s (60 B/px) | of measured peak |")
print("|---|---|---|---|---|")
x = torch.rand(8, 3, 2160, 3840, device=dev)
g = torch.rand(8, 3, 2160, 3840, device=dev)
for q in (10, 20, 30, 40, 50, 60, 70, 80, 90, 95):
    ms, _ = time_fwd_bwd(wmattack.DiffJPEG(True, 2160, 3840, quality=q), x, g, n=10)
    px = 8 * 2160 * 3840
    gbs = px * 60 / ms / 1e6
    print(f"| {q} | {ms:.3f} | {px / ms / 1e3:.0f} | {gbs:.0f} | {gbs / PEAK * 100:.0f} % |", flush=True)

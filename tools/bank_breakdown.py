"""Per-layer cost of the fused store epilogue (developer tool, GPU box): plain no-grad forward vs
forward_into a slice of the K-way batch with Quantization(x + (clamp(v) - x)) in the store."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack

B, H, W = 64, 512, 512
dev = "cuda"
x = torch.rand(B, 3, H, W, device=dev)
out = torch.empty(B, 3, H, W, device=dev)
layers = [wmattack.Resize(), wmattack.JpegMask(70), wmattack.MiddleBlur(3), wmattack.GaussianBlur(), wmattack.Gaussian(),
          wmattack.DiffJPEG(True, H, W, quality=50), wmattack.Identity()]


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


with torch.no_grad():
    for l in layers:
        name = getattr(l, "name", type(l).__name__)
        if isinstance(l, wmattack.Resize):
            plain = timeit(lambda: l(x, resize_ratio=0.75))
        else:
            plain = timeit(lambda: l(x))
        fused = timeit(lambda: l.forward_into(x, out, (x, True, True)))
        print(f"{name:18s} plain fwd {plain:7.1f} us   with store epilogue {fused:7.1f} us", flush=True)

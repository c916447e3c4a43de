"""Randomised shape sweep (developer tool, GPU box): every stencil / JPEG / resize layer on random small [B,3,H,W]
shapes — ragged widths, tiny planes, odd plane counts, strided views — against the CPU oracle, forward and gradient.
    python tools/fuzz_shapes.py [n_cases] [seed]
Prints the worst error per layer; exits non-zero on the first violation."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
sys.path.insert(0, ROOT)
import wmattack
from wmattack import functional as WF
from oracle import attack_oracle as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = "cuda"
worst = {}


def note(name, err, tol, ctx):
    worst[name] = max(worst.get(name, 0.0), err)
    if not (err <= tol):
        print(f"FAIL {name}: err {err:.3e} > {tol:.1e} at {ctx}")
        sys.exit(1)


def fb(fn, x, g):
    xx = x.to(dev).requires_grad_(True)
    y = fn(xx)
    y.backward(g.to(dev))
    return y.detach().cpu(), xx.grad.cpu()


def ofb(fn, x, g):
    xx = x.double().requires_grad_(True)
    y = fn(xx)
    y.backward(g.double())
    return y.detach(), xx.grad


for case in range(n_cases):
    b = int(rng.randint(1, 4))
    h = int(rng.choice([1, 2, 3, 5, 8, 17, 31, 36, 37, 40, 64, 73, 100]))
    w = int(rng.choice([1, 2, 3, 6, 8, 9, 16, 20, 33, 63, 64, 65, 72, 127, 128, 130, 131, 136, 200, 258, 264]))
    x = torch.from_numpy(rng.rand(b, 3, h, w).astype(np.float32))
    if rng.rand() < 0.3:
        x = torch.round(x * 7) / 7                      # ties
    g = torch.from_numpy(rng.rand(b, 3, h, w).astype(np.float32))
    ctx = (b, h, w)
    for k in (3, 5, 7):
        y, gx = fb(wmattack.GaussianBlur(k), x, g)
        yo, go = ofb(lambda t: O.gaussian_blur(t, k), x, g)
        note(f"blur{k}", float((y - yo).abs().max()), 1e-6, ctx); note(f"blur{k}.grad", float((gx - go).abs().max()), 1e-6, ctx)
    for k in (3, 5):
        y, gx = fb(wmattack.MiddleBlur(k), x, g)
        yo, idx = O.median_blur(x, k, return_index=True)
        note(f"median{k}", float((y - yo).abs().max()), 0.0, ctx)
        note(f"median{k}.grad", float((gx - O.median_blur_backward(g, idx, k)).abs().max()), 0.0, ctx)
        note(f"median{k}.nograd", float((wmattack.MiddleBlur(k)(x.to(dev)).cpu() - yo).abs().max()), 0.0, ctx)
    for layer, ref in ((wmattack.JpegMask(50), lambda t: O.jpeg8(t, 50, O.JPEG8_MASK)), (wmattack.JpegCompression(dev), O.jpeg_compression)):
        y, gx = fb(layer, x, g)
        yo, go = ofb(ref, x, g)
        note(type(layer).__name__, float((y - yo).abs().max()), 2e-5, ctx); note(type(layer).__name__ + ".grad", float((gx - go).abs().max()), 2e-5, ctx)
    if h >= 8 and w >= 8:
        r = float(rng.choice([0.5, 0.75, 1.25, 1.5]))
        m = wmattack.Resize()
        y = m(x.to(dev), resize_ratio=r).cpu()
        mid = torch.nn.functional.interpolate(x, size=[int(r * h), int(r * w)], mode="bicubic")
        ref = torch.clamp(torch.nn.functional.interpolate(mid, size=[h, w], mode="bicubic"), 0, 1)
        note("resize", float((y - ref).abs().max()), 2e-5, ctx + (r,))
    # autocast boundary: the same layers on a float16 / bfloat16 image (typed kernels when the rows sit on 16-byte
    # boundaries, one cast otherwise) must equal the float32 kernels on the widened image, gradient rounded to that type
    dt = torch.float16 if rng.rand() < 0.5 else torch.bfloat16
    xh = x.to(dt)
    half_layers = [("blur3", wmattack.GaussianBlur(3), {}), ("median3", wmattack.MiddleBlur(3), {}), ("median5", wmattack.MiddleBlur(5), {}),
                   ("jpegmask", wmattack.JpegMask(50), {})]
    if h >= 8 and w >= 8:
        half_layers.append(("resize", wmattack.Resize(), {"resize_ratio": float(rng.choice([0.5, 0.75, 1.25, 1.5]))}))
    for name, layer, kw in half_layers:
        xa = xh.to(dev).requires_grad_(True); ya = layer(xa, **kw); ya.backward(g.to(dev))
        xb = xh.float().to(dev).requires_grad_(True); yb = layer(xb, **kw); yb.backward(g.to(dev))
        note(f"{name}.half", float((ya - yb).abs().max()), 0.0, ctx + (str(dt),))
        note(f"{name}.half.grad", float((xa.grad.float() - xb.grad.to(dt).float()).abs().max()), 0.0, ctx + (str(dt),))
        if xa.grad.dtype != dt:
            print(f"FAIL {name}: gradient dtype {xa.grad.dtype} for a {dt} image"); sys.exit(1)
    # strided view: a column slice of a wider tensor (odd row stride)
    if w > 8:
        wide = torch.from_numpy(rng.rand(b, 3, h, w + 5).astype(np.float32)).to(dev)
        view = wide[..., 2:2 + w]
        vc = view.cpu().contiguous()
        note("median5.view", float((wmattack.MiddleBlur(5)(view).cpu() - O.median_blur(vc, 5)).abs().max()), 0.0, ctx)
        note("blur3.view", float((wmattack.GaussianBlur(3)(view).cpu() - O.gaussian_blur(vc.double(), 3)).abs().max()), 1e-6, ctx)
torch.cuda.synchronize()
for k in sorted(worst):
    print(f"{k:22s} worst |err| = {worst[k]:.3e}")
print(f"fuzz ok: {n_cases} cases")

"""Per-kernel device timing + quick torch-on-GPU cross-check (developer tool, GPU box only).

    python tools/kbench.py [blur] [median] [resize] [diffjpeg] [jpeg8] [noise] [--shape B H W]

Prints, per kernel: us/launch, algorithmic GB/s, fraction of the measured HBM peak, max-abs error
against a plain torch implementation on the same device (NOT the parity gate — that is tests/).
"""
import json, os, sys
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
sys.path.insert(0, ROOT)
import wmattack
from wmattack import functional as WF

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0

args = [a for a in sys.argv[1:]]
shape = (64, 512, 512)
if "--shape" in args:
    i = args.index("--shape")
    shape = tuple(int(v) for v in args[i + 1:i + 4])
    del args[i:i + 4]
which = set(args) or {"blur", "median", "resize", "diffjpeg", "jpeg8", "noise"}
B, H, W = shape
dev = "cuda"
torch.manual_seed(0)
x = torch.rand(B, 3, H, W, device=dev)
g = torch.rand(B, 3, H, W, device=dev)
px = B * H * W


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def report(name, us, bpp, err=None):
    gbs = px * bpp / us / 1e3
    e = "" if err is None else f"  err={err:.2e}"
    print(f"{name:28s} {us:9.1f} us  {gbs:8.0f} GB/s  {gbs / PEAK * 100:5.1f}% of measured peak{e}", flush=True)


def fwd_bwd(name, f, bpp_f, bpp_b, ref=None):
    xx = x.clone().requires_grad_(True)
    y = f(xx)
    err = None
    if ref is not None:
        err = float((y.detach() - ref(x)).abs().max())
    report(name + ".fwd(nograd)", timeit(lambda: f(x)), bpp_f, err)
    report(name + ".fwd", timeit(lambda: f(xx)), bpp_f)
    y = f(xx)
    errb = None
    if ref is not None:
        xr = x.clone().requires_grad_(True)
        ref(xr).backward(g)
        y.backward(g, retain_graph=True)
        errb = float((xx.grad - xr.grad).abs().max())
        xx.grad = None

    def step():
        xx.grad = None
        y.backward(g, retain_graph=True)
    report(name + ".bwd", timeit(step), bpp_b, errb)


def taps(k, sigma=2.0):
    t = torch.arange(k, dtype=torch.float64) - (k - 1) / 2
    w = torch.exp(-t * t / (2 * sigma * sigma))
    return (w / w.sum()).tolist()


if "blur" in which:
    for k in (3, 5, 7):
        tp = taps(k)
        w2 = torch.tensor(tp, device=dev, dtype=torch.float32)
        w2 = (w2[:, None] * w2[None, :]).expand(3, 1, k, k).contiguous()
        fwd_bwd(f"gaussblur k{k}", lambda t: WF.gaussian_blur(t, tp, 0), 24, 24,
                lambda t: F.conv2d(t, w2, padding=k // 2, groups=3))

if "median" in which:
    def med_ref(k):
        def f(t):
            b, c, h, w = t.shape
            u = F.unfold(t.reshape(b * c, 1, h, w), k, padding=k // 2).view(b, c, k * k, h, w)
            return u.median(dim=2)[0]
        return f
    for k in (3, 5):
        small = B * H * W <= 8 * 512 * 512
        fwd_bwd(f"median k{k}", lambda t: WF.median_blur(t, k), 27, 27, med_ref(k) if small else None)

if "resize" in which:
    for r in (0.5, 0.75, 1.25, 1.5):
        mid = (int(r * H), int(r * W))
        def ref(t, mid=mid):
            m = F.interpolate(t, size=mid, mode="bicubic")
            return torch.clamp(F.interpolate(m, size=(H, W), mode="bicubic"), 0, 1)
        fwd_bwd(f"resize r{r}", lambda t, mid=mid: WF.resize_roundtrip(t, mid, "bicubic"), 24, 36, ref)

if "diffjpeg" in which:
    m = wmattack.DiffJPEG(True, H, W, quality=50)
    fwd_bwd("diffjpeg q50 (saved state)", lambda t: m(t), 24, 36)
    m2 = wmattack.DiffJPEG(True, H, W, quality=50)
    m2.recompute_backward = True
    fwd_bwd("diffjpeg q50 (recompute)", lambda t: m2(t), 24, 36)

if "jpeg8" in which:
    for nm, mod, bb in (("jpegcompression", wmattack.JpegCompression(dev), 24), ("jpegss50", wmattack.JpegSS(50), 36),
                        ("jpegmask50", wmattack.JpegMask(50), 24)):
        fwd_bwd(nm, lambda t, mod=mod: mod(t), 24, bb)

if "noise" in which:
    gn = wmattack.Gaussian()
    fwd_bwd("gaussian", lambda t: gn(t), 24, 24)      # the backward reads gy + the 1-bit clamp mask
    sp = wmattack.SaltPepper(0.01)
    fwd_bwd("saltpepper", lambda t: sp(t), 24, 24)

if "crop" in which:
    box = (int(0.2 * H), int(0.2 * H) + int(0.7 * H), int(0.1 * W), int(0.1 * W) + int(0.75 * W))
    def crop_ref(t):
        c = t[:, :, box[0]:box[1], box[2]:box[3]]
        return F.interpolate(c, size=(H, W), mode="bilinear")
    cm = wmattack.Crop()
    fwd_bwd("crop 0.7x0.75 bilinear", lambda t: cm(t, apex=box)[0], 24, 24, crop_ref)

if "bank" in which:
    # a fixed ratio: every NEW resize geometry is proven once (one sync) before it is cached, which a per-call random
    # ratio would hit on every timed call; a trainer reaches the steady state after <= 513 distinct mid sizes at 512 px
    layers = [wmattack.Resize((0.75, 0.75)), wmattack.JpegMask(70), wmattack.MiddleBlur(3), wmattack.GaussianBlur(), wmattack.Gaussian(), wmattack.Identity()]
    bank = wmattack.AttackBank(layers)
    bank_sep = wmattack.AttackBank(layers); bank_sep.shared_read = False
    quant = wmattack.Quantization()
    def torch_style(t):          # models/IRNp_model.py:609-680 with torch elementwise ops around OUR attack kernels
        import numpy as _np
        sims = [torch.clamp(l(t.detach()), 0, 1) for l in layers]
        rep = t.repeat(len(layers), 1, 1, 1)
        att = rep + (torch.cat(sims, 0) - rep).detach()
        return quant(att)
    xx = x.clone().requires_grad_(True)
    gk = torch.rand(6 * B, 3, H, W, device=dev)
    us_bank = timeit(lambda: bank(xx)); us_sep = timeit(lambda: bank_sep(xx)); us_torch = timeit(lambda: torch_style(xx))
    print(f"6-way bank fwd: shared read of x (bank3) + fused epilogue {us_bank:.0f} us  vs one kernel per member {us_sep:.0f} us  "
          f"vs torch clamp/sub/add/cat + Quantization {us_torch:.0f} us", flush=True)
    yb = bank(xx)
    def stepb():
        xx.grad = None; yb.backward(gk, retain_graph=True)
    yt = torch_style(xx)
    def stept():
        xx.grad = None; yt.backward(gk, retain_graph=True)
    print(f"6-way bank bwd (straight-through sum over slices): {timeit(stepb):.0f} us  vs torch autograd {timeit(stept):.0f} us", flush=True)

if "codec" in which:
    # real-codec round trip (JpegTest): float frames 12 B/px in + 12 out (+3 B/px of byte planes
    # written and read between the two kernels); byte frames 3 + 3
    xs = x * 2 - 1
    xu = (x * 255).to(torch.uint8)
    for s in (2, 0):
        us = timeit(lambda: WF.jpeg_codec(xs, 75, s, "signed"))
        report(f"jpeg codec s={s} float q75", us, 24)
        us = timeit(lambda: WF.jpeg_codec(xu, 75, s, "uint8"))
        report(f"jpeg codec s={s} uint8 q75", us, 6)
    try:
        import io, time
        import numpy as np
        from PIL import Image
        fr = xu[0].permute(1, 2, 0).cpu().numpy()
        t0 = time.perf_counter()
        for _ in range(5):
            buf = io.BytesIO(); Image.fromarray(fr).save(buf, format="JPEG", quality=75, subsampling=2)
            dec = np.array(Image.open(io.BytesIO(buf.getvalue())))
        dt = (time.perf_counter() - t0) / 5
        got = WF.jpeg_codec(xu[:1], 75, 2, "uint8")[0].permute(1, 2, 0).cpu().numpy()
        print(f"Pillow one {H}x{W} frame, in memory: {dt * 1e3:.2f} ms = {H * W / dt / 1e6:.1f} Mpix/s on one host core; "
              f"device result identical: {bool((got == dec).all())}", flush=True)
    except ImportError:
        pass

"""Generate tests/golden/attack_golden.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, CPU):

    python tests/golden/make_golden.py [--ref /root/reference]

The reference is imported as-is behind three harness shims:
  1. empty stand-ins for `matplotlib` / `matplotlib.pyplot`;
  2. a stand-in `kornia.filters` (kornia is not installed/vendored/pinned) implementing
     kornia 0.6.x semantics for MedianBlur / GaussianBlur2d  -> those two goldens are
     "parity unpinned by the reference";
  3. `.cuda()` made a no-op (CPU container).
Nothing from the reference is copied into this repository; only its OUTPUTS on seeded
inputs are stored.  The GPU box never sees /root/reference.
"""
import argparse
import os
import random
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def install_shims():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    class MedianBlur(nn.Module):
        def __init__(self, kernel_size):
            super().__init__()
            self.k = kernel_size

        def forward(self, x):
            kh, kw = self.k
            b, c, h, w = x.shape
            kernel = torch.eye(kh * kw, dtype=x.dtype).view(kh * kw, 1, kh, kw)
            feat = F.conv2d(x.reshape(b * c, 1, h, w), kernel,
                            padding=((kh - 1) // 2, (kw - 1) // 2), stride=1)
            return feat.view(b, c, -1, h, w).median(dim=2)[0]

    class GaussianBlur2d(nn.Module):
        def __init__(self, kernel_size, sigma, border_type="reflect"):
            super().__init__()
            self.k, self.s, self.border = kernel_size, sigma, border_type

        def forward(self, x):
            def taps(k, s):
                xs = torch.arange(k, dtype=x.dtype) - k // 2
                g = torch.exp(-xs ** 2 / (2 * s ** 2))
                return g / g.sum()
            ky, kx = taps(self.k[0], self.s[0]), taps(self.k[1], self.s[1])
            c = x.shape[1]
            w2 = torch.outer(ky, kx).view(1, 1, *self.k).repeat(c, 1, 1, 1)
            ry, rx = self.k[0] // 2, self.k[1] // 2
            return F.conv2d(F.pad(x, (rx, rx, ry, ry), mode=self.border), w2, groups=c)

    kornia = types.ModuleType("kornia")
    kf = types.ModuleType("kornia.filters")
    kf.MedianBlur, kf.GaussianBlur2d = MedianBlur, GaussianBlur2d
    kornia.filters = kf
    sys.modules["kornia"], sys.modules["kornia.filters"] = kornia, kf

    nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self


def smooth_8bit(shape, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(shape, generator=g)
    k = torch.ones(1, 1, 5, 5) / 25
    b, c, h, w = shape
    x = F.conv2d(F.pad(x.view(b * c, 1, h, w), (2, 2, 2, 2), mode="replicate"), k).view(shape)
    x = (x - x.min()) / (x.max() - x.min())
    return torch.round(x * 255) / 255


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                  "attack_golden.npz"))
    args = ap.parse_args()
    install_shims()
    sys.path.insert(0, args.ref)

    from utils import JPEG as RJ
    import noise_layers as NL
    from noise_layers.jpeg import Jpeg, JpegSS, JpegMask
    from noise_layers.jpeg_compression import JpegCompression
    from noise_layers.gaussian_blur import GaussianBlur
    from noise_layers.gaussian import Gaussian
    from noise_layers.gaussian_noise import GN
    from noise_layers.gaussian_filter import GF
    from noise_layers.middle_filter import MiddleBlur
    from noise_layers.salt_pepper_noise import SaltPepper
    from noise_layers.resize import Resize
    from noise_layers.crop import Crop
    from noise_layers.crop import Dropout as CropDropout
    from noise_layers.dropout import Dropout as MaskDropout
    from noise_layers.combined import Combined
    from noise_layers.identity import Identity
    sys.path.insert(0, os.path.join(args.ref, "models", "modules"))
    import Quantization as RQ

    G = {}

    def put(name, t):
        G[name] = np.ascontiguousarray(t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t))

    def rand(shape, seed):
        return torch.rand(shape, generator=torch.Generator().manual_seed(seed))

    x32 = rand((2, 3, 32, 32), 0)
    g32 = rand((2, 3, 32, 32), 1)
    x4832 = rand((1, 3, 48, 32), 2)
    g4832 = rand((1, 3, 48, 32), 3)
    x20 = rand((2, 3, 20, 20), 4)
    g20 = rand((2, 3, 20, 20), 5)
    xs32 = smooth_8bit((2, 3, 32, 32), 6)
    x2028 = rand((1, 3, 20, 28), 7)
    for n, t in (("x32", x32), ("g32", g32), ("x4832", x4832), ("g4832", g4832), ("x20", x20),
                 ("g20", g20), ("xs32", xs32), ("x2028", x2028)):
        put(n, t)

    def fwd_bwd(mod, x, g, **kw):
        xx = x.clone().requires_grad_(True)
        y = mod(xx, **kw)
        y.backward(g)
        return y.detach(), xx.grad.detach()

    # ---- DiffJPEG -------------------------------------------------------------------
    roundings = {"r0": RJ.round_only_at_0, "cubic": RJ.diff_round, "hard": torch.round}
    for q in (10, 50, 75, 95):
        for rn, rf in roundings.items():
            if q in (10, 95) and rn != "r0":
                continue
            m = RJ.DiffJPEG(True, 32, 32, quality=q, rounding=rf)
            for xn, x in (("x32", x32), ("xs32", xs32)):
                y, gx = fwd_bwd(m, x, g32)
                put(f"diffjpeg/q{q}/{rn}/{xn}/y", y.contiguous())
                put(f"diffjpeg/q{q}/{rn}/{xn}/gx", gx)
            if q == 50:
                yq, cbq, crq = m.compress(xs32)
                put(f"diffjpeg/q{q}/{rn}/xs32/coef_y", yq)
                put(f"diffjpeg/q{q}/{rn}/xs32/coef_cb", cbq)
                put(f"diffjpeg/q{q}/{rn}/xs32/coef_cr", crq)
                put(f"diffjpeg/q{q}/{rn}/xs32/decompressed", m.decompress(yq, cbq, crq).contiguous())
    m = RJ.DiffJPEG(True, 48, 32, quality=30)
    y, gx = fwd_bwd(m, x4832, g4832)
    put("diffjpeg/q30/r0/x4832/y", y.contiguous())
    put("diffjpeg/q30/r0/x4832/gx", gx)
    put("diffjpeg/name_q30", np.frombuffer(m.name.encode(), dtype=np.uint8))
    # saturated input: exercises the clamp (and its tie gradient) of utils/JPEG.py:467-468
    xsat = (x32 > 0.5).float()
    put("xsat", xsat)
    m = RJ.DiffJPEG(True, 32, 32, quality=50)
    y, gx = fwd_bwd(m, xsat, g32)
    put("diffjpeg/q50/r0/xsat/y", y.contiguous())
    put("diffjpeg/q50/r0/xsat/gx", gx)

    # ---- Jpeg / JpegSS / JpegMask ------------------------------------------------------
    for cls, cn in ((Jpeg, "jpeg"), (JpegSS, "jpegss"), (JpegMask, "jpegmask")):
        for q in (30, 50, 90):
            for sub in (0, 2):
                m = cls(q, subsample=sub)
                for xn, x, g in (("x20", x20, g20), ("xs32", xs32, g32)):
                    y, gx = fwd_bwd(m, x, g) if cn != "jpeg" else (m(x), torch.zeros_like(x))
                    put(f"{cn}/q{q}/s{sub}/{xn}/y", y)
                    if cn != "jpeg":
                        put(f"{cn}/q{q}/s{sub}/{xn}/gx", gx)
    m = Jpeg(50)
    dct, pw, ph = m.yuv_dct(xs32, 0)
    put("jpeg/q50/s0/xs32/quantised", m.std_quantization(dct, m.scale_factor))
    put("jpeg/name_q50", np.frombuffer(m.name.encode(), dtype=np.uint8))

    # ---- JpegCompression (forward only: reference backward crashes on torch 2.11) ------
    jc = JpegCompression(torch.device("cpu"))
    with torch.no_grad():
        put("jpegcompression/x20/y", jc(x20.clone()))
        put("jpegcompression/x32/y", jc(x32.clone()))
        put("jpegcompression/x2028/y", jc(x2028.clone()))

    # ---- GaussianBlur -----------------------------------------------------------------
    for k in (3, 5, 7):
        m = GaussianBlur(kernel_size=k)
        y, gx = fwd_bwd(m, x2028, rand(x2028.shape, 8))
        put(f"gaussianblur/k{k}/x2028/y", y)
        put(f"gaussianblur/k{k}/x2028/gx", gx)
    put("gaussianblur/g2028", rand(x2028.shape, 8))

    # ---- kornia-backed layers (unpinned) ----------------------------------------------
    for k in (3, 5):
        put(f"middleblur/k{k}/x2028/y", MiddleBlur(k)(x2028))
        put(f"middleblur/k{k}/xs32/y", MiddleBlur(k)(xs32))
    put("gf/s1.5k7/x2028/y", GF(1.5, 7)((x2028, x2028)))

    # ---- stochastic layers: record the random tensor the reference drew -----------------
    captured = {}
    orig_normal = torch.nn.init.normal_

    def spy_normal(t, mean=0.0, std=1.0):
        torch.manual_seed(11)
        out = orig_normal(t, mean, std)
        captured["noise"] = out.clone()
        return out
    torch.nn.init.normal_ = spy_normal
    y, gx = fwd_bwd(Gaussian(), x32, g32)
    torch.nn.init.normal_ = orig_normal
    put("gaussian/x32/noise", captured["noise"])
    put("gaussian/x32/y", y)
    put("gaussian/x32/gx", gx)

    np.random.seed(12)
    y = GN(0.0025)((x32, x32))
    np.random.seed(12)
    put("gn/x32/noise", torch.Tensor(np.random.normal(0, 0.0025 ** 0.5, x32.shape)))
    put("gn/x32/y", y)

    torch.manual_seed(13)
    y, gx = fwd_bwd(SaltPepper(0.1), x32, g32)
    torch.manual_seed(13)
    put("saltpepper/p0.1/x32/rdn", torch.rand(x32.shape))
    put("saltpepper/p0.1/x32/y", y)
    put("saltpepper/p0.1/x32/gx", gx)

    cover = rand((2, 3, 32, 32), 14)
    put("cover32", cover)
    torch.manual_seed(15)
    y = CropDropout(0.5)((x32, cover))
    torch.manual_seed(15)
    put("cropdropout/p0.5/x32/rdn", torch.rand(x32.shape))
    put("cropdropout/p0.5/x32/y", y)

    np.random.seed(16)
    y = MaskDropout((0.5, 1))(x32, cover)
    np.random.seed(16)
    p = np.random.uniform(0.5, 1)
    put("maskdropout/x32/keep", np.float64(p))
    put("maskdropout/x32/mask", np.random.choice([0.0, 1.0], x32.shape[2:], p=[1 - p, p]))
    put("maskdropout/x32/y", y)

    # ---- Resize / Crop ------------------------------------------------------------------
    for mode in ("bicubic", "bilinear"):
        for r in (0.5, 0.7, 1.3, 1.5):
            y, gx = fwd_bwd(Resize(interpolation_method=mode), x2028, rand(x2028.shape, 8), resize_ratio=r)
            put(f"resize/{mode}/r{r}/x2028/y", y)
            put(f"resize/{mode}/r{r}/x2028/gx", gx)
    y, gx = fwd_bwd(Resize(), xsat, g32, resize_ratio=0.8)
    put("resize/bicubic/r0.8/xsat/y", y)
    put("resize/bicubic/r0.8/xsat/gx", gx)
    np.random.seed(17)
    y = Resize()(x32)
    np.random.seed(17)
    put("resize/random/ratio", np.float64(np.random.rand() * (1.5 - 0.5) + 0.5))
    put("resize/random/x32/y", y)

    np.random.seed(18)
    xx = x32.clone().requires_grad_(True)
    y, apex = Crop()(xx)
    y.backward(g32)
    put("crop/seed18/x32/y", y)
    put("crop/seed18/x32/gx", xx.grad)
    put("crop/seed18/x32/apex", np.array(apex, dtype=np.int64))
    np.random.seed(19)
    y, apex = Crop()(x2028, min_rate=0.7, max_rate=0.9)
    put("crop/seed19/x2028/y", y)
    put("crop/seed19/x2028/apex", np.array(apex, dtype=np.int64))
    y, apex = Crop()(x32, apex=(3, 20, 5, 31))
    put("crop/apex/x32/y", y)
    np.random.seed(20)
    outs = Crop().cropped_out(x32, min_rate=0.5)
    put("cropped_out/seed20/x32/scaled", outs[0])
    put("cropped_out/seed20/x32/zero_images", outs[1])
    put("cropped_out/seed20/x32/mask", outs[2])
    put("cropped_out/seed20/x32/apex", np.array(outs[3], dtype=np.float64))
    put("cropped_out/seed20/x32/new_images", outs[4].contiguous())

    # ---- Quantization, Combined, Identity ------------------------------------------------
    y, gx = fwd_bwd(RQ.Quantization(), x32, g32)
    put("quantization/x32/y", y)
    put("quantization/x32/gx", gx)
    random.seed(21)
    comb = Combined([Identity(), Jpeg(50), JpegSS(70), JpegMask(30), Resize()])
    names = []
    np.random.seed(22)
    for _ in range(12):
        comb(x20)
        names.append(comb.name)
    put("combined/seed21/names", np.frombuffer(",".join(names).encode(), dtype=np.uint8))
    put("identity/is_same_object", np.array(int(Identity()(x20) is x20)))


    # ---- round-2 additions (appended: earlier fixtures stay bit-identical) ---------------------
    # cropped_out gradients (crop.py:78-118): both differentiable outputs, seeded rectangle
    np.random.seed(23)
    xx = x32.clone().requires_grad_(True)
    outs = Crop().cropped_out(xx, min_rate=0.5)
    gz = rand(tuple(x32.shape), 24)
    (outs[0] * g32).sum().add((outs[1] * gz).sum()).backward()
    put("cropped_out/seed23/gz", gz)
    put("cropped_out/seed23/x32/scaled", outs[0])
    put("cropped_out/seed23/x32/zero_images", outs[1])
    put("cropped_out/seed23/x32/apex", np.array(outs[3], dtype=np.float64))
    put("cropped_out/seed23/x32/gx", xx.grad)
    # ... and with a caller-supplied fractional apex (models/IRN_model.py:1569 passes one)
    xx = x32.clone().requires_grad_(True)
    outs = Crop().cropped_out(xx, apex=(0.125, 0.75, 0.25, 0.9375), min_rate=0.5)
    (outs[0] * g32).sum().add((outs[1] * gz).sum()).backward()
    put("cropped_out/apex/x32/scaled", outs[0])
    put("cropped_out/apex/x32/zero_images", outs[1])
    put("cropped_out/apex/x32/gx", xx.grad)
    # cropped_for_outpainting (crop.py:57-76): pure slicing with two seeded rectangles
    np.random.seed(25)
    real_h = rand((2, 3, 32, 32), 26)
    put("outpainting/real_H", real_h)
    a, b, c = Crop().cropped_for_outpainting(x32, real_h)
    put("outpainting/seed25/x32/new_images", a.contiguous())
    put("outpainting/seed25/x32/zero_images", b.contiguous())
    put("outpainting/seed25/x32/GT", c.contiguous())
    # utils/compression.py + utils/decompression.py: size passed at CALL time (decompression.py:162)
    from utils import compression as RC, decompression as RD
    for rn, rf in roundings.items():
        comp = RC.compress_jpeg(rounding=rf, factor=RJ.quality_to_factor(40))
        dec = RD.decompress_jpeg(rounding=rf, factor=RJ.quality_to_factor(40))
        cy, ccb, ccr = comp(x4832)
        put(f"codec_calltime/q40/{rn}/x4832/coef_y", cy)
        put(f"codec_calltime/q40/{rn}/x4832/coef_cb", ccb)
        put(f"codec_calltime/q40/{rn}/x4832/coef_cr", ccr)
        put(f"codec_calltime/q40/{rn}/x4832/y", dec(cy, ccb, ccr, 48, 32).contiguous())
    # Fourier-series rounding surrogate (utils/JPEG_utils.py:36-41) through DiffJPEG
    from utils import JPEG_utils as RU
    m = RJ.DiffJPEG(True, 32, 32, quality=50, rounding=RU.diff_round)
    for xn, x in (("x32", x32), ("xs32", xs32)):
        y, gx = fwd_bwd(m, x, g32)
        put(f"diffjpeg/q50/fourier/{xn}/y", y.contiguous())
        put(f"diffjpeg/q50/fourier/{xn}/gx", gx)

    np.savez_compressed(args.out, **G)
    print(f"wrote {args.out}: {len(G)} arrays, {os.path.getsize(args.out) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()

"""CPU oracle for the tailored attacking layer — TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU tensors, any float dtype,
fp64 for "truth") of the arithmetic of the reference's attack layer
(yingqichao/video-watermarking-forgery-detection: ``noise_layers/*`` and
``utils/JPEG.py::DiffJPEG``).  It is the *checker* for the CUDA path:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import it;
  * the product (``wmattack``) never imports it and has no CPU fallback.

Parity pinning: every function below is checked in ``tests/test_oracle_golden.py``
against fixtures in ``tests/golden/`` that were produced by running the UNMODIFIED
reference modules in the build container (``tests/golden/make_golden.py``).
Exceptions — "parity unpinned by the reference": ``median_blur`` and
``gaussian_filter_reflect`` restate kornia 0.6.x (``kornia.filters.MedianBlur`` /
``GaussianBlur2d``), a third-party dependency of the reference that is neither vendored
nor pinned nor installable here; their goldens come from the restatement of kornia's
published algorithm (zero-pad + one-hot conv + ``torch.median``; separable reflect conv).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Gradients are obtained by autograd through these functions in fp64,
plus the explicit ``median_blur_backward``.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# constants
# --------------------------------------------------------------------------------------

# Annex-K luminance table, row = vertical frequency (noise_layers/jpeg.py:54-63;
# utils/JPEG.py:98-104 stores the TRANSPOSE of this).
STD_LUMA = np.array(
    [[16, 11, 10, 16, 24, 40, 51, 61],
     [12, 12, 14, 19, 26, 58, 60, 55],
     [14, 13, 16, 24, 40, 57, 69, 56],
     [14, 17, 22, 29, 51, 87, 80, 62],
     [18, 22, 37, 56, 68, 109, 103, 77],
     [24, 35, 55, 64, 81, 104, 113, 92],
     [49, 64, 78, 87, 103, 121, 120, 101],
     [72, 92, 95, 98, 112, 100, 103, 99]], dtype=np.float64)


def _std_chroma() -> np.ndarray:
    # noise_layers/jpeg.py:67-76 and utils/JPEG.py:107-110 (symmetric, 99-filled)
    t = np.full((8, 8), 99.0)
    t[:4, :4] = [[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]]
    return t


STD_CHROMA = _std_chroma()

ROUND_ONLY_AT_0 = 0   # utils/JPEG.py:482   (== JpegSS.round_ss, noise_layers/jpeg.py:255)
ROUND_CUBIC = 1       # utils/JPEG.py:472   diff_round
ROUND_HARD = 2        # torch.round
ROUND_FOURIER = 3     # utils/JPEG_utils.py:36  (9-term Fourier series)
ROUND_NONE = -1       # test aid: no rounding, i.e. the pre-round quotient C / (table * factor)


def dct_matrix(dtype=torch.float64) -> torch.Tensor:
    """Orthonormal 8-point DCT-II matrix D[u, x] = c(u)/2 cos((2x+1)u pi/16).

    Same transform as utils/JPEG.py:195-208 (cos tensor x outer(alpha,alpha)/4) and
    noise_layers/jpeg.py:117-121 (``coff``)."""
    u = torch.arange(8, dtype=torch.float64).view(8, 1)
    x = torch.arange(8, dtype=torch.float64).view(1, 8)
    d = 0.5 * torch.cos((2 * x + 1) * u * math.pi / 16)
    d[0] = d[0] / math.sqrt(2.0)
    return d.to(dtype)


def quality_to_factor(quality: float) -> float:
    """utils/JPEG.py:487-498."""
    q = 5000.0 / quality if quality < 50 else 200.0 - quality * 2
    return q / 100.0


def apply_rounding(q: torch.Tensor, mode: int) -> torch.Tensor:
    if mode == ROUND_NONE:
        return q
    if mode == ROUND_ONLY_AT_0:          # utils/JPEG.py:482-484
        inside = (q.abs() < 0.5).to(q.dtype)
        return inside * q ** 3 + (1 - inside) * q
    if mode == ROUND_CUBIC:              # utils/JPEG.py:472-479
        r = torch.round(q)
        return r + (q - r) ** 3
    if mode == ROUND_HARD:
        return torch.round(q)
    if mode == ROUND_FOURIER:            # utils/JPEG_utils.py:36-41
        s = 0
        for n in range(1, 10):
            s = s + math.pow(-1, n + 1) / n * torch.sin(2 * math.pi * n * q)
        return q - s / math.pi
    raise ValueError(mode)


# --------------------------------------------------------------------------------------
# 8x8 blocking helpers (own formulation: unfold the plane as [.., H/8, 8, W/8, 8])
# --------------------------------------------------------------------------------------

def _to_blocks(p: torch.Tensor) -> torch.Tensor:
    """[B, H, W] -> [B, H/8, W/8, 8, 8] (block raster order, as utils/JPEG.py:176-181)."""
    b, h, w = p.shape
    return p.reshape(b, h // 8, 8, w // 8, 8).permute(0, 1, 3, 2, 4)


def _from_blocks(blk: torch.Tensor) -> torch.Tensor:
    b, nh, nw = blk.shape[:3]
    return blk.permute(0, 1, 3, 2, 4).reshape(b, nh * 8, nw * 8)


def _dct2(blk: torch.Tensor) -> torch.Tensor:
    d = dct_matrix(blk.dtype)
    return d @ blk @ d.t()


def _idct2(c: torch.Tensor) -> torch.Tensor:
    d = dct_matrix(c.dtype)
    return d.t() @ c @ d


# --------------------------------------------------------------------------------------
# DiffJPEG  (utils/JPEG.py:501-540; Appendix A.1 of SURVEY.md)
# --------------------------------------------------------------------------------------

_FWD_YCC = np.array([[0.299, 0.587, 0.114],
                     [-0.168736, -0.331264, 0.5],
                     [0.5, -0.418688, -0.081312]])            # utils/JPEG.py:125-127
_INV_YCC = np.array([[1.0, 0.0, 1.402],
                     [1.0, -0.344136, -0.714136],
                     [1.0, 1.772, 0.0]])                       # utils/JPEG.py:419-421


def diffjpeg_tables(factor, dtype=torch.float64) -> Tuple[torch.Tensor, torch.Tensor]:
    """Un-rounded `table*factor`; luminance table transposed (utils/JPEG.py:104, 229, 251).

    The reference's constants are float32 (np.float32 arrays -> nn.Parameter) and
    `table*factor` is a float32 product; for fp32 runs we mimic that, for fp64 "truth"
    we keep the exact product."""
    ty = torch.tensor(STD_LUMA.T.copy(), dtype=dtype) * factor
    tc = torch.tensor(STD_CHROMA, dtype=dtype) * factor
    return ty, tc


def diffjpeg_compress(x: torch.Tensor, factor, rounding: int = ROUND_ONLY_AT_0):
    """compress_jpeg.forward, utils/JPEG.py:279-291.  x: [B,3,H,W] in [0,1], H,W % 16 == 0.
    factor: python float or per-sample tensor [B].  Returns (y, cb, cr) as
    [B, nblk, 8, 8] rounded quantised coefficients."""
    dt = x.dtype
    b, _, h, w = x.shape
    assert h % 16 == 0 and w % 16 == 0
    m = torch.tensor(_FWD_YCC, dtype=dt)
    ycc = torch.einsum("kc,bchw->bkhw", m, x * 255)
    y = ycc[:, 0]
    cb = F.avg_pool2d(ycc[:, 1:2] + 128, 2)[:, 0]          # :154-157
    cr = F.avg_pool2d(ycc[:, 2:3] + 128, 2)[:, 0]
    if torch.is_tensor(factor):
        fac = factor.to(dt).view(b, 1, 1, 1, 1)
    else:
        fac = factor
    ty = torch.tensor(STD_LUMA.T.copy(), dtype=dt) * fac
    tc = torch.tensor(STD_CHROMA, dtype=dt) * fac
    out = []
    for plane, tab in ((y, ty), (cb, tc), (cr, tc)):
        c = _dct2(_to_blocks(plane) - 128)                  # :205-206
        q = apply_rounding(c / tab, rounding)               # :229-230 / :251-252
        out.append(q.reshape(b, -1, 8, 8))
    return tuple(out)


def diffjpeg_decompress(y, cb, cr, height: int, width: int, factor):
    """decompress_jpeg.forward, utils/JPEG.py:452-469."""
    dt = y.dtype
    b = y.shape[0]
    if torch.is_tensor(factor):
        fac = factor.to(dt).view(b, 1, 1, 1, 1)
    else:
        fac = factor
    ty = torch.tensor(STD_LUMA.T.copy(), dtype=dt) * fac
    tc = torch.tensor(STD_CHROMA, dtype=dt) * fac
    planes = []
    for q, tab, (hh, ww) in ((y, ty, (height, width)),
                             (cb, tc, (height // 2, width // 2)),
                             (cr, tc, (height // 2, width // 2))):
        blk = q.reshape(b, hh // 8, ww // 8, 8, 8) * tab    # :310 / :328
        planes.append(_from_blocks(_idct2(blk) + 128))      # :350-354, :371-376
    yy = planes[0]
    up = lambda p: p.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)  # :393-404
    ycc = torch.stack([yy, up(planes[1]) - 128, up(planes[2]) - 128], dim=1)
    mi = torch.tensor(_INV_YCC, dtype=dt)
    rgb = torch.einsum("kc,bchw->bkhw", mi, ycc)            # :425-428
    # binary min/max against constant tensors (:467-468): gradient 1/2 at exact ties
    rgb = torch.min(torch.full_like(rgb, 255.0), torch.max(torch.zeros_like(rgb), rgb))
    return rgb / 255


def diffjpeg(x: torch.Tensor, quality=None, rounding: int = ROUND_ONLY_AT_0, factor=None):
    """DiffJPEG.forward, utils/JPEG.py:535-540 (contiguous NCHW values)."""
    if factor is None:
        if torch.is_tensor(quality):
            factor = torch.where(quality < 50, 5000.0 / quality, 200.0 - 2 * quality) / 100.0
        else:
            factor = quality_to_factor(quality)
    y, cb, cr = diffjpeg_compress(x, factor, rounding)
    return diffjpeg_decompress(y, cb, cr, x.shape[2], x.shape[3], factor)


# --------------------------------------------------------------------------------------
# Jpeg / JpegSS / JpegMask  (noise_layers/jpeg.py:48-306; Appendix A.2)
# --------------------------------------------------------------------------------------

JPEG8_HARD, JPEG8_SS, JPEG8_MASK = 0, 1, 2

_FWD_YUV_MBRS = np.array([[0.299, 0.587, 0.114],
                          [-0.1687, -0.3313, 0.5],
                          [0.5, -0.4187, -0.0813]])                   # jpeg.py:147-155
_INV_YUV_MBRS = np.array([[1.0, 0.0, 1.40198758],
                          [1.0, -0.344113281, -0.714103821],
                          [1.0, 1.77197812, 0.0]])                    # jpeg.py:157-163


def jpeg8_scale(q: float) -> float:
    """noise_layers/jpeg.py:221."""
    return 2 - q * 0.02 if q >= 50 else 50 / q


def jpeg8_tables(scale: float):
    """clamp(round(std*scale), 1) in float32, as noise_layers/jpeg.py:54-76 does it
    (float32 tensor times python scalar, torch.round = half-to-even)."""
    ly = (torch.tensor(STD_LUMA, dtype=torch.float32) * scale).round().clamp(min=1)
    lc = (torch.tensor(STD_CHROMA, dtype=torch.float32) * scale).round().clamp(min=1)
    return ly, lc


def _pad8(x: torch.Tensor):
    h, w = x.shape[2:]
    ph, pw = (8 - h % 8) % 8, (8 - w % 8) % 8
    return F.pad(x, (0, pw, 0, ph)), ph, pw


def _subsample2(yuv: torch.Tensor) -> torch.Tensor:
    """noise_layers/jpeg.py:202-211: inside every 8x8 block, U/V odd rows <- the even row
    above, then odd columns <- the even column to the left."""
    out = yuv.clone()
    uv = out[:, 1:3]
    uv[:, :, 1::2, :] = uv[:, :, 0::2, :]
    uv[:, :, :, 1::2] = uv[:, :, :, 0::2]
    return out


def jpeg8_coefficients(x: torch.Tensor, subsample: int = 0) -> torch.Tensor:
    """JpegBasic.yuv_dct, noise_layers/jpeg.py:165-187 -> coefficient image [B,3,Hp,Wp]."""
    dt = x.dtype
    xp, _, _ = _pad8(x * 255)
    yuv = torch.einsum("kc,bchw->bkhw", torch.tensor(_FWD_YUV_MBRS, dtype=dt), xp)
    if subsample == 2:
        yuv = _subsample2(yuv)
    b, c, hp, wp = yuv.shape
    blk = _to_blocks(yuv.reshape(b * c, hp, wp))
    return _from_blocks(_dct2(blk)).reshape(b, c, hp, wp)


def jpeg8_quantised(x: torch.Tensor, q: float, variant: int = JPEG8_HARD, subsample: int = 0):
    """std_quantization output (noise_layers/jpeg.py:52-82): integers for JPEG8_HARD."""
    dt = x.dtype
    coef = jpeg8_coefficients(x, subsample)
    ly, lc = jpeg8_tables(jpeg8_scale(q))
    hp, wp = coef.shape[2:]
    tab = torch.stack([ly, lc, lc]).to(dt).repeat(1, hp // 8, wp // 8)
    qq = coef / tab
    if variant == JPEG8_HARD:
        return torch.round(qq), tab
    return apply_rounding(qq, ROUND_ONLY_AT_0), tab


def jpeg8_prequant(x: torch.Tensor, q: float, subsample: int = 0) -> torch.Tensor:
    """Test aid: the quotient dct / table that std_quantization (noise_layers/jpeg.py:52-82) rounds,
    as a [B,3,Hp,Wp] coefficient image."""
    coef = jpeg8_coefficients(x, subsample)
    ly, lc = jpeg8_tables(jpeg8_scale(q))
    hp, wp = coef.shape[2:]
    return coef / torch.stack([ly, lc, lc]).to(x.dtype).repeat(1, hp // 8, wp // 8)


def jpeg8(x: torch.Tensor, q: float, variant: int = JPEG8_HARD, subsample: int = 0):
    """Jpeg / JpegSS / JpegMask forward (noise_layers/jpeg.py:226-240, 259-273, 295-306).
    Correct for any H, W (the reference's re-blocking is square-only, jpeg.py:123-127)."""
    dt = x.dtype
    h, w = x.shape[2:]
    if variant == JPEG8_MASK:
        coef = jpeg8_coefficients(x, subsample)
        keep = torch.zeros(3, 8, 8, dtype=dt)
        keep[0, :5, :5] = 1
        keep[1:, :3, :3] = 1                                   # jpeg.py:288-293
        hp, wp = coef.shape[2:]
        deq = coef * keep.repeat(1, hp // 8, wp // 8)
    else:
        r, tab = jpeg8_quantised(x, q, variant, subsample)
        deq = r * tab                                           # jpeg.py:84-113
    b, c, hp, wp = deq.shape
    yuv = _from_blocks(_idct2(_to_blocks(deq.reshape(b * c, hp, wp)))).reshape(b, c, hp, wp)
    rgb = torch.einsum("kc,bchw->bkhw", torch.tensor(_INV_YUV_MBRS, dtype=dt), yuv)
    return rgb[:, :, :h, :w] / 255                              # jpeg.py:189-200 (no clamp)


# --------------------------------------------------------------------------------------
# JpegCompression (HiDDeN)  (noise_layers/jpeg_compression.py:65-159; Appendix A.3)
# --------------------------------------------------------------------------------------

_FWD_YUV_HIDDEN = np.array([[0.299, 0.587, 0.114],
                            [-0.14713, -0.28886, 0.436],
                            [0.615, -0.51499, -0.10001]])             # :51-55
_INV_YUV_HIDDEN = np.array([[1.0, 0.0, 1.13983],
                            [1.0, -0.39465, -0.58060],
                            [1.0, 2.03211, 0.0]])                     # :58-62


def zigzag_keep_mask(keep: int) -> np.ndarray:
    """get_jpeg_yuv_filter_mask, noise_layers/jpeg_compression.py:29-39 (one 8x8 window)."""
    order = sorted(((a, b) for a in range(8) for b in range(8)),
                   key=lambda p: (p[0] + p[1], -p[1] if (p[0] + p[1]) % 2 else p[1]))
    m = np.zeros((8, 8))
    for a, b in order[:keep]:
        m[a, b] = 1
    return m


def hidden_dct_matrices(dtype=torch.float64):
    """Un-normalised analysis matrix A[k,n]=cos(pi/8 (n+1/2) k) (:42-43) and synthesis
    S[n,k] = (cos(pi/8 (n+1/2) k) - [k==0]/2) / 4 (:46-48 with gen_filters' argument order)."""
    k = torch.arange(8, dtype=torch.float64).view(8, 1)
    n = torch.arange(8, dtype=torch.float64).view(1, 8)
    a = torch.cos(math.pi / 8 * (n + 0.5) * k)                  # [k, n]
    s = (torch.cos(math.pi / 8 * (n + 0.5) * k) - 0.5 * (k == 0)) * 0.25   # [k, n]
    return a.to(dtype), s.t().contiguous().to(dtype)            # A[k,n], S[n,k]


def jpeg_compression(x: torch.Tensor, keep: Sequence[int] = (25, 9, 9)) -> torch.Tensor:
    """JpegCompression.forward, noise_layers/jpeg_compression.py:128-160 (functional form;
    the module's own autograd crashes on torch 2.11)."""
    dt = x.dtype
    h, w = x.shape[2:]
    xp, _, _ = _pad8(x)
    yuv = torch.einsum("kc,bchw->bkhw", torch.tensor(_FWD_YUV_HIDDEN, dtype=dt), xp)
    b, c, hp, wp = yuv.shape
    a, s = hidden_dct_matrices(dt)
    blk = _to_blocks(yuv.reshape(b * c, hp, wp)).reshape(b, c, hp // 8, wp // 8, 8, 8)
    coef = a @ blk @ a.t()                                       # [ky, kx]
    mask = torch.stack([torch.tensor(zigzag_keep_mask(k_), dtype=dt) for k_ in keep])
    coef = coef * mask.view(1, 3, 1, 1, 8, 8)
    rec = s @ coef @ s.t()
    yuv2 = _from_blocks(rec.reshape(b * c, hp // 8, wp // 8, 8, 8)).reshape(b, c, hp, wp)
    rgb = torch.einsum("kc,bchw->bkhw", torch.tensor(_INV_YUV_HIDDEN, dtype=dt), yuv2)
    return rgb[:, :, :h, :w]


# --------------------------------------------------------------------------------------
# Blur / median
# --------------------------------------------------------------------------------------

def gaussian_taps(k: int, sigma: float = 2.0, dtype=torch.float64) -> torch.Tensor:
    """1-D normalised taps whose outer product is the 2-D kernel of
    noise_layers/gaussian_blur.py:17-38."""
    m = (k - 1) / 2.0
    t = torch.exp(-((torch.arange(k, dtype=torch.float64) - m) ** 2) / (2 * sigma ** 2))
    return (t / t.sum()).to(dtype)


def gaussian_blur(x: torch.Tensor, k: int = 3, sigma: float = 2.0) -> torch.Tensor:
    """GaussianBlur.forward, noise_layers/gaussian_blur.py:53-56: depth-wise k x k,
    sigma hard-wired to 2, ZERO padding (k-1)/2 (int())."""
    c = x.shape[1]
    t = gaussian_taps(k, sigma, x.dtype)
    w2 = torch.outer(t, t)
    w2 = (w2 / w2.sum()).view(1, 1, k, k).repeat(c, 1, 1, 1)
    return F.conv2d(x, w2, padding=int((k - 1) / 2), groups=c)


def gaussian_filter_reflect(x: torch.Tensor, k: int = 7, sigma: float = 1.0) -> torch.Tensor:
    """GF (noise_layers/gaussian_filter.py:5-13) = kornia GaussianBlur2d((k,k),(s,s)),
    border_type='reflect' — restated from kornia 0.6.x (parity unpinned)."""
    c = x.shape[1]
    xs = torch.arange(k, dtype=torch.float64) - k // 2
    t = torch.exp(-xs ** 2 / (2 * sigma ** 2))
    t = (t / t.sum()).to(x.dtype)
    w2 = torch.outer(t, t).view(1, 1, k, k).repeat(c, 1, 1, 1)
    r = k // 2
    return F.conv2d(F.pad(x, (r, r, r, r), mode="reflect"), w2, groups=c)


def median_windows(x: torch.Tensor, k: int) -> torch.Tensor:
    """[B,C,H,W] -> [B,C,k*k,H,W]: zero-padded k x k neighbourhoods in raster order
    (kornia get_binary_kernel2d one-hot conv; parity unpinned, see module docstring)."""
    b, c, h, w = x.shape
    r = (k - 1) // 2
    cols = F.unfold(F.pad(x, (r, r, r, r)).reshape(b * c, 1, h + 2 * r, w + 2 * r), k)
    return cols.view(b, c, k * k, h, w)


def median_blur(x: torch.Tensor, k: int, return_index: bool = False):
    """MiddleBlur.forward (noise_layers/middle_filter.py:11-13) = kornia MedianBlur((k,k)).
    Index convention (ours, since torch.median's tie choice is unspecified): FIRST window
    position in raster order whose value equals the median."""
    win = median_windows(x, k)
    val = win.median(dim=2)[0]
    if not return_index:
        return val
    eq = win == val.unsqueeze(2)
    idx = eq.to(torch.uint8).argmax(dim=2).to(torch.uint8)     # first True
    return val, idx


def median_blur_backward(gy: torch.Tensor, idx: torch.Tensor, k: int) -> torch.Tensor:
    """gx[p] = sum of gy[q] over outputs q whose arg-median is p; padded positions absorb."""
    b, c, h, w = gy.shape
    r = (k - 1) // 2
    gpad = torch.zeros(b, c, h + 2 * r, w + 2 * r, dtype=gy.dtype)
    ii = idx.long()
    dy, dx = ii // k, ii % k
    yy = torch.arange(h).view(1, 1, h, 1) + dy
    xx = torch.arange(w).view(1, 1, 1, w) + dx
    flat = (yy * (w + 2 * r) + xx).reshape(b, c, -1)
    gpad.view(b, c, -1).scatter_add_(2, flat, gy.reshape(b, c, -1))
    return gpad[:, :, r:r + h, r:r + w].contiguous()


# --------------------------------------------------------------------------------------
# Elementwise attacks (random tensors are INJECTED so that parity is exact)
# --------------------------------------------------------------------------------------

def gaussian_noise_clamped(x, noise):
    """Gaussian.forward, noise_layers/gaussian.py:10-17: clamp(x + N(mean, std^2), 0, 1)."""
    return torch.clamp(x + noise, 0, 1)


def gaussian_noise_additive(x, noise):
    """GN.gaussian_noise, noise_layers/gaussian_noise.py:13-16 (no clamp)."""
    return x + noise


def salt_pepper(x, rdn, prob: float):
    """SaltPepper.sp_noise, noise_layers/salt_pepper_noise.py:11-19."""
    p0 = prob / 2
    p1 = 1 - p0
    out = torch.where(rdn > p1, torch.zeros_like(x), x)
    return torch.where(rdn < p0, torch.ones_like(out), out)


def dropout_mask(noised, cover, mask_hw):
    """dropout.Dropout.forward, noise_layers/dropout.py:14-27; mask [H,W] in {0,1}."""
    m = mask_hw.to(noised.dtype).expand_as(noised)
    return noised * m + cover * (1 - m)


def dropout_elementwise(image, cover, rdn, prob: float):
    """crop.Dropout.forward, noise_layers/crop.py:142-147."""
    return torch.where(rdn > prob * 1.0, cover, image)


def cropout(image, cover, box):
    """Cropout.forward as intended, noise_layers/crop.py:128-134 (functional: returns a
    new tensor instead of mutating `cover`)."""
    h0, h1, w0, w1 = box
    out = cover.clone()
    out[:, :, h0:h1, w0:w1] = image[:, :, h0:h1, w0:w1]
    return out


def quantization(x):
    """Quant.forward, models/modules/Quantization.py:7-10 (gradient: identity)."""
    return torch.round(x * 255.0) / 255.0


# --------------------------------------------------------------------------------------
# Resize / Crop  (noise_layers/resize.py:28-55, noise_layers/crop.py:32-55; Appendix A.6)
# --------------------------------------------------------------------------------------

def interp_matrix(n_in: int, n_out: int, mode: str, dtype=torch.float64) -> torch.Tensor:
    """Dense [n_out, n_in] matrix of F.interpolate(size=n_out, mode, align_corners=False)
    along one axis (ATen upsample_bicubic2d A=-0.75 / upsample_bilinear2d semantics)."""
    w = torch.zeros(n_out, n_in, dtype=torch.float64)
    scale = n_in / n_out
    for o in range(n_out):
        rho = scale * (o + 0.5) - 0.5
        if mode == "bilinear":
            rho = max(rho, 0.0)
            i0 = min(int(math.floor(rho)), n_in - 1)
            i1 = min(i0 + 1, n_in - 1)
            lam = rho - i0
            w[o, i0] += 1 - lam
            w[o, i1] += lam
        elif mode == "bicubic":
            a = -0.75
            i0 = math.floor(rho)
            t = rho - i0
            w1 = lambda z: ((a + 2) * z - (a + 3)) * z * z + 1
            w2 = lambda z: ((a * z - 5 * a) * z + 8 * a) * z - 4 * a
            taps = [w2(t + 1), w1(t), w1(1 - t), w2(2 - t)]
            for j, wt in enumerate(taps):
                w[o, min(max(i0 - 1 + j, 0), n_in - 1)] += wt
        else:
            raise ValueError(mode)
    return w.to(dtype)


def interpolate(x: torch.Tensor, size: Tuple[int, int], mode: str) -> torch.Tensor:
    """Separable-matrix restatement of F.interpolate (validated against torch in tests)."""
    wh = interp_matrix(x.shape[2], size[0], mode, x.dtype)
    ww = interp_matrix(x.shape[3], size[1], mode, x.dtype)
    return wh @ x @ ww.t()


def resize_mid_size(h: int, w: int, ratio: float) -> Tuple[int, int]:
    return int(ratio * h), int(ratio * w)                        # resize.py:35


def resize(x: torch.Tensor, ratio: float, mode: str = "bicubic") -> torch.Tensor:
    """Resize.forward, noise_layers/resize.py:28-55: down/up by `ratio`, back, clamp[0,1]."""
    h, w = x.shape[2:]
    mid = interpolate(x, resize_mid_size(h, w, ratio), mode)
    return torch.clamp(interpolate(mid, (h, w), mode), 0, 1)


def crop_resize(x: torch.Tensor, box, mode: str = "bilinear") -> torch.Tensor:
    """Crop.forward's arithmetic, noise_layers/crop.py:48-53."""
    h0, h1, w0, w1 = box
    return interpolate(x[:, :, h0:h1, w0:w1], x.shape[2:], mode)


def crop_box_from_rng(shape, rng: np.random.RandomState, min_rate=0.5, max_rate=1.0):
    """RNG call order of Crop.forward + get_random_rectangle_inside
    (noise_layers/crop.py:13-40): rand() x2, ratio coupling, randint per free axis."""
    if min_rate:
        hr = min_rate + (max_rate - min_rate) * rng.rand()
        wr = min_rate + (max_rate - min_rate) * rng.rand()
    else:
        hr = 0.3 + 0.7 * rng.rand()
        wr = 0.3 + 0.7 * rng.rand()
    hr = min(hr, wr + 0.2)
    wr = min(wr, hr + 0.2)
    ih, iw = shape[2], shape[3]
    rh, rw = int(hr * ih), int(wr * iw)
    h0 = 0 if rh == ih else rng.randint(0, ih - rh)
    w0 = 0 if rw == iw else rng.randint(0, iw - rw)
    return h0, h0 + rh, w0, w0 + rw

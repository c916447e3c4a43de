"""nn.Module surface of the attack layer, signature-compatible with the reference
(yingqichao/video-watermarking-forgery-detection, paths below are relative to its root).

Each class keeps the reference constructor / forward signature and `.name` string; the
arithmetic runs in libwmattack.so.  Host-side random DECISIONS (which layer, resize ratio,
crop box, keep ratio) use the same RNG calls in the same order as the reference, so a seeded
run makes the same choices; per-element randomness is generated in-kernel (Philox) unless
`host_rng=True` (reproduce the reference's host draw exactly) or a tensor is injected.
"""
from __future__ import annotations

import random
import string
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import functional as F_
from .functional import quality_to_factor  # noqa: F401  (re-export, utils/JPEG.py:487)


# ---- noise_layers/__init__.py:4-9 ---------------------------------------------------------
def get_random_float(float_range):
    return random.random() * (float_range[1] - float_range[0]) + float_range[0]


def get_random_int(int_range):
    return random.randint(int_range[0], int_range[1])


def _first(x):
    """Bare-tensor layers also accept the upstream HiDDeN calling convention
    `layer([image, cover])` (hidden_models/encoder_decoder.py:26-27) and use element 0."""
    if isinstance(x, (list, tuple)):
        return x[0]
    return x


# ---- rounding surrogates (utils/JPEG.py:472-484) -------------------------------------------
def diff_round(x):
    """utils/JPEG.py:472 — as a plain torch expression (used as a ctor tag for DiffJPEG)."""
    return torch.round(x) + (x - torch.round(x)) ** 3


def round_only_at_0(x):
    """utils/JPEG.py:482."""
    cond = (torch.abs(x) < 0.5).float()
    return cond * (x ** 3) + (1 - cond) * x


def _rounding_mode(rounding) -> int:
    if isinstance(rounding, int):
        return rounding
    if rounding is torch.round:
        return F_.ROUND_HARD
    name = getattr(rounding, "__name__", "")
    mod = getattr(rounding, "__module__", "") or ""
    if name == "round_only_at_0" or name == "round_ss":
        return F_.ROUND_ONLY_AT_0
    if name == "diff_round_back":
        return F_.ROUND_CUBIC
    if name == "diff_round":
        # utils/JPEG.py's diff_round is the cubic; utils/JPEG_utils.py's is the Fourier series
        return F_.ROUND_FOURIER if mod.endswith("JPEG_utils") else F_.ROUND_CUBIC
    if name == "round":
        return F_.ROUND_HARD
    raise ValueError(f"unsupported rounding function {rounding!r}: use round_only_at_0, diff_round, "
                     "torch.round or a WM_ROUND_* integer")


def _set_name(module: nn.Module, name: str) -> None:
    """`.name` updates inside forward(): skip nn.Module.__setattr__ (4-5 us per call) when nothing changes."""
    if module.__dict__.get("name") != name:
        object.__setattr__(module, "name", name)


class Identity(nn.Module):
    """noise_layers/identity.py:5-16."""

    def __init__(self):
        super().__init__()
        self.name = "Identity"

    def forward(self, image):
        return image


def _identity_into(self, x, out, ep):
    F_.attack_epilogue(x, x, ep[1], ep[2], out=out)
    return True


Identity.forward_into = _identity_into
Identity.bank3_kind = "identity"                 # member of the shared-read bank kernel (functional._BankFusedFn)
Identity.bank3_member = lambda self: ()


class Combined(nn.Module):
    """noise_layers/combined.py:6-20: one member per call, chosen with python `random`."""

    def __init__(self, list=None):  # noqa: A002  (reference argument name)
        super().__init__()
        if list is None:
            list = [Identity()]
        self.list = list            # plain python list, as upstream (not an nn.ModuleList)
        self.name = "NotChosenYet"

    def forward_into(self, x, out, ep, id=None):  # noqa: A002
        """AttackBank hook: same choice logic as forward(), then the member's fused path if it has one."""
        if id is None or id >= len(self.list):
            id = get_random_int([0, len(self.list) - 1])
        selected = self.list[id]
        _set_name(self, selected.name)
        into = getattr(selected, "forward_into", None)
        if into is not None and into(x, out, ep):
            return True
        y = selected(x)
        F_.attack_epilogue(x, y[0] if isinstance(y, tuple) else y, ep[1], ep[2], out=out)
        return True

    def forward(self, image_and_cover, id=None):  # noqa: A002
        if id is None or id >= len(self.list):
            id = get_random_int([0, len(self.list) - 1])
        selected = self.list[id]
        _set_name(self, selected.name)
        return selected(image_and_cover)


# ---- DiffJPEG (utils/JPEG.py:256-540) --------------------------------------------------------
class compress_jpeg(nn.Module):
    """utils/JPEG.py:256-291 (and utils/compression.py:147).  forward(image) -> (y, cb, cr)."""

    def __init__(self, rounding=torch.round, factor=1):
        super().__init__()
        self.rounding = _rounding_mode(rounding)
        self.factor = factor

    def forward(self, image):
        return F_.diffjpeg_compress(image, self.factor, self.rounding)


class decompress_jpeg(nn.Module):
    """utils/JPEG.py:431-469; also accepts the call-time size of utils/decompression.py:162."""

    def __init__(self, height=None, width=None, rounding=torch.round, factor=1):
        super().__init__()
        self.height, self.width = height, width
        self.factor = factor

    def forward(self, y, cb, cr, height=None, width=None):
        h = self.height if height is None else height
        w = self.width if width is None else width
        return F_.diffjpeg_decompress(y, cb, cr, int(h), int(w), self.factor)


class DiffJPEG(nn.Module):
    """utils/JPEG.py:501-540.  Positional order kept: (differentiable, height, width, quality,
    rounding) — so the upstream call DiffJPEG(90) still binds 90 to `differentiable` and keeps
    quality 75.  `forward(image, quality=None)` adds an optional per-call
    override: a python number or a per-sample tensor [B] (quality sweep)."""

    # False: the forward saves 7 B/px of state for the backward.  Set True on an instance to save
    # nothing and let the backward recompute the forward from x (lower memory, ~40 us slower at 64x512^2).
    recompute_backward = False

    def __init__(self, differentiable=True, height=512, width=512, quality=75, rounding=round_only_at_0):
        super().__init__()
        self.name = "DiffJPEG" + str(quality)
        self.height, self.width = height, width
        self.quality = quality
        self.rounding = _rounding_mode(rounding)
        factor = quality_to_factor(quality)
        self.factor = factor
        self.compress = compress_jpeg(rounding=self.rounding, factor=factor)
        self.decompress = decompress_jpeg(height, width, rounding=self.rounding, factor=factor)

    @staticmethod
    def _factor_of(quality):
        if torch.is_tensor(quality):
            q = quality.detach().to(torch.float64)      # python-double arithmetic, like quality_to_factor
            return (torch.where(q < 50, 5000.0 / q, 200.0 - 2.0 * q) / 100.0).to(torch.float32)
        return quality_to_factor(quality)

    def forward_into(self, x, out, ep):
        """No-grad forward written into `out` through the fused store epilogue (AttackBank)."""
        return F_.diffjpeg_into(x, self.factor, self.rounding, out, ep)

    def forward(self, image, quality=None):
        image = _first(image)
        factor = self.factor if quality is None else self._factor_of(quality)
        return F_.diffjpeg(image, factor, self.rounding, self.recompute_backward)


# ---- Jpeg / JpegSS / JpegMask (noise_layers/jpeg.py) -----------------------------------------
_STD_LUMA = np.array(
    [[16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55],
     [14, 13, 16, 24, 40, 57, 69, 56], [14, 17, 22, 29, 51, 87, 80, 62],
     [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
     [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]], dtype=np.float32)
_STD_CHROMA = np.full((8, 8), 99, dtype=np.float32)
_STD_CHROMA[:4, :4] = [[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]]

_MBRS_FWD = [0.299, 0.587, 0.114, -0.1687, -0.3313, 0.5, 0.5, -0.4187, -0.0813]       # jpeg.py:147-155
_MBRS_INV = [1.0, 0.0, 1.40198758, 1.0, -0.344113281, -0.714103821, 1.0, 1.77197812, 0.0]  # jpeg.py:157-163


def _mbrs_tables(scale_factor: float) -> np.ndarray:
    """clamp(round(std * scale), min=1) evaluated exactly like noise_layers/jpeg.py:54-76:
    float32 tensor times python scalar, torch.round (half to even)."""
    ly = (torch.tensor(_STD_LUMA) * scale_factor).round().clamp(min=1).numpy()
    lc = (torch.tensor(_STD_CHROMA) * scale_factor).round().clamp(min=1).numpy()
    return np.stack([ly, lc, lc]).astype(np.float32)


class JpegBasic(nn.Module):
    """Shared host logic of noise_layers/jpeg.py:48-211 (tables are built once, not per call)."""

    _variant = F_.JPEG8_HARD
    _prefix = "Jpeg"

    def __init__(self, Q, subsample=0):
        super().__init__()
        self.name = self._prefix + str(Q)
        self.Q = Q
        self.scale_factor = 2 - self.Q * 0.02 if self.Q >= 50 else 50 / self.Q    # jpeg.py:221
        self.subsample = subsample
        if subsample not in (0, 2):
            raise ValueError("subsample must be 0 or 2 (noise_layers/jpeg.py:202-211)")
        fwd = [v * 255.0 for v in _MBRS_FWD]            # image * 255 (jpeg.py:168)
        inv = [v / 255.0 for v in _MBRS_INV]            # image_rgb / 255 (jpeg.py:200)
        self._params = F_.make_jpeg8_params(fwd, inv, self._table(), self._variant, subsample)

    def _table(self) -> np.ndarray:
        return _mbrs_tables(self.scale_factor)

    def forward_into(self, x, out, ep):
        return F_.jpeg8_into(x, self._params, out, ep)

    def forward(self, image):
        return F_.jpeg8(_first(image), self._params)

    def quantised(self, image):
        """std_quantization output (jpeg.py:52-82) as a [B,3,Hp,Wp] coefficient image."""
        return F_.jpeg8_quantised(_first(image), self._params)


class Jpeg(JpegBasic):
    """noise_layers/jpeg.py:214-240 (torch.round -> zero gradient)."""


class JpegSS(JpegBasic):
    """noise_layers/jpeg.py:243-273 (round_ss == round_only_at_0)."""
    _variant = F_.JPEG8_SS
    _prefix = "JpegSS"


class JpegMask(JpegBasic):
    """noise_layers/jpeg.py:276-306: keep Y[:5,:5], U/V[:3,:3]; Q only affects the name."""
    _variant = F_.JPEG8_MASK
    _prefix = "JpegMask"

    def _table(self) -> np.ndarray:
        keep = np.zeros((3, 8, 8), dtype=np.float32)
        keep[0, :5, :5] = 1
        keep[1:, :3, :3] = 1
        return keep


class JpegTest(nn.Module):
    """noise_layers/jpeg.py:10-45: the REAL codec — upstream saves every frame through PIL/libjpeg
    to a temp file and reads it back.  Here the same pixels (bit-identical to Pillow/libjpeg-turbo,
    tests/test_gpu_parity.py) are computed on the device by wm_jpegcodec: entropy coding is
    lossless, so only libjpeg's integer colour/DCT/quantisation/upsampling arithmetic remains.
    Non-differentiable, expects [-1,1] input, like upstream; `path` is accepted and unused."""

    def __init__(self, Q, subsample=2, path="temp/"):
        super().__init__()
        self.Q, self.subsample, self.path = Q, subsample, path
        self.name = "JpegTest" + str(Q)

    def forward(self, image):
        image = _first(image)
        # upstream draws one temp-file name per frame from python's `random` (get_path, jpeg.py:17-18,32): consume
        # the same numbers so that a seeded run keeps making the same Combined / get_random_int choices afterwards
        for _ in range(image.shape[0]):
            random.sample(string.ascii_letters + string.digits, 16)
        return F_.jpeg_codec(image, self.Q, self.subsample, "signed")


# ---- JpegCompression (noise_layers/jpeg_compression.py) --------------------------------------
def _zigzag_keep(keep: int) -> np.ndarray:
    """jpeg_compression.py:29-39, one 8x8 window."""
    order = sorted(((a, b) for a in range(8) for b in range(8)),
                   key=lambda p: (p[0] + p[1], -p[1] if (p[0] + p[1]) % 2 else p[1]))
    m = np.zeros((8, 8), dtype=np.float32)
    for a, b in order[:keep]:
        m[a, b] = 1
    return m


class JpegCompression(nn.Module):
    """noise_layers/jpeg_compression.py:65-159 (HiDDeN).  Un-normalised conv DCT, zig-zag keep
    mask, matching synthesis filters == orthonormal DCT -> mask -> inverse; any H, W.
    Unlike the reference (whose autograd crashes on torch >= 2.x) it is differentiable."""

    def __init__(self, device=None, yuv_keep_weights=(25, 9, 9)):
        super().__init__()
        self.device = device
        self.yuv_keep_weighs = yuv_keep_weights
        self.name = "JpegCompression"
        fwd = [0.299, 0.587, 0.114, -0.14713, -0.28886, 0.436, 0.615, -0.51499, -0.10001]   # :51-55
        inv = [1.0, 0.0, 1.13983, 1.0, -0.39465, -0.58060, 1.0, 2.03211, 0.0]               # :58-62
        table = np.stack([_zigzag_keep(k) for k in yuv_keep_weights])
        self._params = F_.make_jpeg8_params(fwd, inv, table, F_.JPEG8_MASK, 0)

    def forward_into(self, x, out, ep):
        return F_.jpeg8_into(x, self._params, out, ep)

    def forward(self, noised_image):
        return F_.jpeg8(_first(noised_image), self._params)


# ---- blur / median ---------------------------------------------------------------------------
def _gaussian_taps(k: int, sigma: float, centre: float) -> list:
    xs = np.arange(k, dtype=np.float64) - centre
    t = np.exp(-xs ** 2 / (2.0 * sigma ** 2))
    return list((t / t.sum()).astype(np.float64))


class GaussianBlur(nn.Module):
    """noise_layers/gaussian_blur.py:7-57: depth-wise k x k Gaussian, sigma hard-wired to 2
    (get_gaussian_kernel's default, :17), zero padding int((k-1)/2)."""

    def __init__(self, kernel_size=3, channels=3):
        super().__init__()
        self.kernel_size = kernel_size
        self.channels = channels
        self.name = "G_Blur"
        if kernel_size % 2 == 0:
            raise ValueError("even kernel sizes change the output size upstream; only odd sizes are supported")
        self._taps = tuple(_gaussian_taps(kernel_size, 2.0, (kernel_size - 1) / 2.0))

    def forward_into(self, x, out, ep):
        ok = F_.gaussian_blur_into(x, self._taps, out, ep)
        if ok:
            _set_name(self, "GaussianBlur")  # gaussian_blur.py:54 renames on first use
        return ok

    @property
    def bank3_kind(self):
        return "blur" if self.kernel_size == 3 else None

    def bank3_member(self):
        _set_name(self, "GaussianBlur")
        return self._taps

    def forward(self, tensor, cover_image=None):
        _set_name(self, "GaussianBlur")
        tensor = _first(tensor)
        if tensor.shape[1] != self.channels:
            raise RuntimeError(f"GaussianBlur built for {self.channels} channels, got {tensor.shape[1]}")
        return F_.gaussian_blur(tensor, self._taps, border=0)


class GF(nn.Module):
    """noise_layers/gaussian_filter.py:5-13 -> kornia GaussianBlur2d((k,k),(sigma,sigma)),
    reflect border (kornia semantics restated; parity unpinned by the reference)."""

    def __init__(self, sigma, kernel=7):
        super().__init__()
        self.name = "GF"
        self._taps = tuple(_gaussian_taps(kernel, float(sigma), kernel // 2))

    def forward(self, image_and_cover):
        image = _first(image_and_cover)
        return F_.gaussian_blur(image, self._taps, border=1)


class MiddleBlur(nn.Module):
    """noise_layers/middle_filter.py:5-13 -> kornia MedianBlur((k,k)): zero padding, k in {3,5}."""

    def __init__(self, kernel):
        super().__init__()
        if kernel not in (3, 5):
            raise ValueError("MiddleBlur supports kernel 3 (IRNcrop_model.py:95) and 5 (IRN_model.py:109)")
        self.kernel = kernel
        self.name = "MiddleBlur" + str(kernel)

    def forward_into(self, x, out, ep):
        return F_.median_blur_into(x, self.kernel, out, ep)

    @property
    def bank3_kind(self):
        return "median" if self.kernel == 3 else None

    def bank3_member(self):
        return ()

    def forward(self, image):
        return F_.median_blur(_first(image), self.kernel)


# ---- elementwise -----------------------------------------------------------------------------
class Gaussian(nn.Module):
    """noise_layers/gaussian.py:4-17: clamp(x + N(mean, stddev^2), 0, 1)."""

    def __init__(self, host_rng: bool = False):
        super().__init__()
        self.name = "Gaussian"
        self.host_rng = host_rng

    def forward_into(self, x, out, ep):
        _set_name(self, "Gaussian")
        return F_.gaussian_noise_into(x, 0.0, 0.05, True, out, ep)

    @property
    def bank3_kind(self):
        return None if self.host_rng else "noise"

    def bank3_member(self):
        _set_name(self, "Gaussian")
        return (0.0, 0.05, True)

    def forward(self, tensor, cover_image=None, mean=0, stddev=0.05, noise=None):
        _set_name(self, "Gaussian")
        tensor = _first(tensor)
        if noise is None and self.host_rng:
            noise = torch.nn.init.normal_(torch.empty(tensor.size(), device=tensor.device), mean, stddev)
        return F_.gaussian_noise(tensor, mean, stddev, clamp=True, noise=noise)


class GN(nn.Module):
    """noise_layers/gaussian_noise.py:6-20: image + N(mean, var), takes an (image, cover) pair."""

    def __init__(self, var, mean=0, host_rng: bool = False):
        super().__init__()
        self.var, self.mean = var, mean
        self.name = "GN"
        self.host_rng = host_rng

    def gaussian_noise(self, image, mean, var, noise=None):
        if noise is None and self.host_rng:
            noise = torch.Tensor(np.random.normal(mean, var ** 0.5, image.shape)).to(image.device)
        return F_.gaussian_noise(image, mean, var ** 0.5, clamp=False, noise=noise)

    def forward(self, image_and_cover, noise=None):
        return self.gaussian_noise(_first(image_and_cover), self.mean, self.var, noise)


class SaltPepper(nn.Module):
    """noise_layers/salt_pepper_noise.py:5-23."""

    def __init__(self, prob, host_rng: bool = False):
        super().__init__()
        self.prob = prob
        self.name = "SaltPepper"
        self.host_rng = host_rng

    def sp_noise(self, image, prob, rdn=None):
        if rdn is None and self.host_rng:
            rdn = torch.rand(image.shape).to(image.device)      # reference: CPU draw + H2D (:14)
        return F_.salt_pepper(image, prob, rdn)

    def forward(self, image, rdn=None):
        return self.sp_noise(_first(image), self.prob, rdn)


class ElementDropout(nn.Module):
    """crop.Dropout, noise_layers/crop.py:136-147: where(rand > prob, cover, image) per element."""

    def __init__(self, prob=0.5, host_rng: bool = False):
        super().__init__()
        self.prob = prob
        self.name = "Dropout"
        self.host_rng = host_rng

    def forward(self, image_and_cover, rdn=None):
        image, cover_image = image_and_cover
        if rdn is None and self.host_rng:
            rdn = torch.rand(image.shape).to(image.device)
        return F_.dropout_elementwise(image, cover_image, self.prob, rdn)


class MaskDropout(nn.Module):
    """dropout.Dropout, noise_layers/dropout.py:4-27: one [H,W] Bernoulli(keep) mask shared by
    batch and channels; keep ~ U(keep_min, keep_max) from np.random.uniform as upstream."""

    def __init__(self, keep_ratio_range=(0.5, 1), host_rng: bool = False):
        super().__init__()
        self.keep_min, self.keep_max = keep_ratio_range[0], keep_ratio_range[1]
        self.name = "Dropout"
        self.host_rng = host_rng

    def forward(self, noised_image, cover_image, mask=None):
        _set_name(self, "Dropout")
        mask_percent = np.random.uniform(self.keep_min, self.keep_max)
        h, w = noised_image.shape[2:]
        if mask is None:
            if self.host_rng:
                m = np.random.choice([0.0, 1.0], (h, w), p=[1 - mask_percent, mask_percent])
                mask = torch.tensor(m, device=noised_image.device, dtype=torch.float)
            else:
                mask = F_.bernoulli_mask(h, w, mask_percent, noised_image.device)
        return F_.dropout_mask(noised_image, cover_image, mask)


class Dropout(nn.Module):
    """Name-compatible front for BOTH upstream `Dropout`s:
    `from noise_layers import *` yields crop.Dropout(prob) called as layer((image, cover));
    `from noise_layers.dropout import Dropout` yields Dropout(keep_ratio_range) called as
    layer(noised, cover).  The constructor argument type selects the flavour."""

    def __init__(self, prob=0.5, keep_ratio_range=None, host_rng: bool = False):
        super().__init__()
        self.name = "Dropout"
        if keep_ratio_range is None and isinstance(prob, (tuple, list)):
            keep_ratio_range, prob = prob, None
        self._mask = MaskDropout(keep_ratio_range, host_rng) if keep_ratio_range is not None else None
        self._elem = ElementDropout(prob, host_rng) if keep_ratio_range is None else None

    def forward(self, a, b=None, **kw):
        if self._mask is not None:
            if b is None:
                a, b = a
            return self._mask(a, b, **kw)
        if b is not None:
            a = (a, b)
        return self._elem(a, **kw)


class Quantization(nn.Module):
    """models/modules/Quantization.py:16-21: round(x*255)/255 with identity gradient."""

    def __init__(self, clamp01: bool = False):
        super().__init__()
        self.clamp01 = clamp01

    def forward(self, input):  # noqa: A002
        return F_.quantize8(input, self.clamp01)


# ---- Resize / Crop -----------------------------------------------------------------------------
def random_float(min, max):  # noqa: A002
    """noise_layers/resize.py:6-13."""
    return np.random.rand() * (max - min) + min


class Resize(nn.Module):
    """noise_layers/resize.py:15-55."""

    def __init__(self, resize_ratio_range=(0.5, 1.5), interpolation_method="bicubic"):
        super().__init__()
        self.name = "Resize"
        self.resize_ratio_min = resize_ratio_range[0]
        self.resize_ratio_max = resize_ratio_range[1]
        if interpolation_method not in ("bicubic", "bilinear"):
            raise ValueError("interpolation_method must be 'bicubic' or 'bilinear'")
        self.interpolation_method = interpolation_method

    def forward_into(self, x, out, ep):
        h, w = x.shape[2], x.shape[3]
        r = random_float(self.resize_ratio_min, self.resize_ratio_max)        # same RNG call as forward()
        mid = (int(r * h), int(r * w))
        if F_.resize_roundtrip_into(x, mid, self.interpolation_method, out, ep):
            return True
        out_plain = F_.resize_roundtrip(x, mid, self.interpolation_method)      # keep the ratio that was drawn
        F_.attack_epilogue(x, out_plain, ep[1], ep[2], out=out)
        return True

    def forward(self, noised_image, resize_ratio=None):
        _set_name(self, "Resize")
        noised_image = _first(noised_image)
        h, w = noised_image.shape[2], noised_image.shape[3]
        if resize_ratio is None:
            resize_ratio = random_float(self.resize_ratio_min, self.resize_ratio_max)
        mid = (int(resize_ratio * h), int(resize_ratio * w))          # resize.py:35
        return F_.resize_roundtrip(noised_image, mid, self.interpolation_method)


class Crop(nn.Module):
    """noise_layers/crop.py:8-118."""

    def __init__(self):
        super().__init__()
        self.name = "Crop"

    def get_random_rectangle_inside(self, image_shape, height_ratio, width_ratio):
        image_height, image_width = image_shape[2], image_shape[3]
        remaining_height = int(height_ratio * image_height)
        remaining_width = int(width_ratio * image_width)
        height_start = 0 if remaining_height == image_height else np.random.randint(0, image_height - remaining_height)
        width_start = 0 if remaining_width == image_width else np.random.randint(0, image_width - remaining_width)
        return height_start, height_start + remaining_height, width_start, width_start + remaining_width

    def _ratios(self, min_rate, max_rate, couple):
        if min_rate:
            self.height_ratio = min_rate + (max_rate - min_rate) * np.random.rand()
            self.width_ratio = min_rate + (max_rate - min_rate) * np.random.rand()
        else:
            self.height_ratio = 0.3 + 0.7 * np.random.rand()
            self.width_ratio = 0.3 + 0.7 * np.random.rand()
        self.height_ratio = min(self.height_ratio, self.width_ratio + couple)
        self.width_ratio = min(self.width_ratio, self.height_ratio + couple)

    def forward(self, image, apex=None, min_rate=0.5, max_rate=1.0):
        self._ratios(min_rate, max_rate, 0.2)                          # crop.py:33-40
        if apex is not None:
            h_start, h_end, w_start, w_end = apex
        else:
            h_start, h_end, w_start, w_end = self.get_random_rectangle_inside(image.shape, self.height_ratio, self.width_ratio)
        window = F_.slice_window(image.shape, h_start, h_end, w_start, w_end)     # clamps like image[:, :, a:b, c:d]
        scaled = F_.interpolate(image, image.shape[2:], "bilinear", window=window)   # crop.py:48-53
        return scaled, (h_start, h_end, w_start, w_end)

    def cropped_for_outpainting(self, image, real_H, min_rate=0.5, min_rate_2=0.7, apex=None):
        """crop.py:57-76 (pure slicing)."""
        self.height_ratio = min_rate + (1 - min_rate) * np.random.rand()
        self.width_ratio = min_rate + (1 - min_rate) * np.random.rand()
        self.height_ratio = min(self.height_ratio, self.width_ratio + 0.3)
        self.width_ratio = min(self.width_ratio, self.height_ratio + 0.3)
        h0, h1, w0, w1 = self.get_random_rectangle_inside(image.shape, self.height_ratio, self.width_ratio)
        new_images = image[:, :, h0:h1, w0:w1]
        zero_images = torch.zeros_like(new_images)
        self.height_ratio = min_rate_2 + (1 - min_rate_2) * np.random.rand()
        self.width_ratio = min_rate_2 + (1 - min_rate_2) * np.random.rand()
        self.height_ratio = min(self.height_ratio, self.width_ratio + 0.3)
        self.width_ratio = min(self.width_ratio, self.height_ratio + 0.3)
        a0, a1, b0, b1 = self.get_random_rectangle_inside(new_images.shape, self.height_ratio, self.width_ratio)
        zero_images[:, :, a0:a1, b0:b1] = image[:, :, a0:a1, b0:b1]
        return new_images, zero_images, real_H[:, :, h0:h1, w0:w1]

    def cropped_out(self, image, apex=None, min_rate=None, max_rate=1.0):
        """crop.py:78-118: bicubic up + clamp, bicubic back + clamp, paste, straight-through
        "dual reshape" difference; returns the reference's 5-tuple."""
        self._ratios(min_rate, max_rate, 0.3)
        H, W = image.shape[2], image.shape[3]
        if apex is not None:
            h0, h1, w0, w1 = apex
            h0, h1, w0, w1 = int(h0 * H), int(h1 * H), int(w0 * W), int(w1 * W)
        else:
            h0, h1, w0, w1 = self.get_random_rectangle_inside(image.shape, self.height_ratio, self.width_ratio)
        new_images = image[:, :, h0:h1, w0:w1]
        mask = torch.ones_like(image)
        mask[:, :, h0:h1, w0:w1] = 0
        zero_images = image * (1 - mask)
        window = F_.slice_window(image.shape, h0, h1, w0, w1)
        scaled_images = F_.interpolate(image, (H, W), "bicubic", window=window, clamp=True)
        scaled_back = F_.interpolate(scaled_images, (window[2], window[3]), "bicubic", clamp=True)
        zero_images_gt = torch.zeros_like(image)
        zero_images_gt[:, :, h0:h1, w0:w1] = scaled_back
        dual_reshape_diff = (zero_images_gt - zero_images).clone().detach()
        zero_images = zero_images + dual_reshape_diff
        return scaled_images, zero_images, mask, (h0 / H, h1 / H, w0 / W, w1 / W), new_images


class Cropout(nn.Module):
    """noise_layers/crop.py:121-134 as intended (upstream raises AttributeError):
    paste a random rectangle of `image` into `cover`.  Returns a new tensor (no in-place edit)."""

    def __init__(self, height_ratio, width_ratio):
        super().__init__()
        self.height_ratio, self.width_ratio = height_ratio, width_ratio
        self.name = "Cropout"

    def forward(self, image_and_cover, box=None):
        image, cover_image = image_and_cover
        if box is None:
            box = Crop.get_random_rectangle_inside(self, image.shape, self.height_ratio, self.width_ratio)
        return F_.cropout(image, cover_image, box)


# ---- neighbours of the attack layer in the trainers' step ------------------------
class AttackEpilogue(nn.Module):
    """clamp + straight-through + Quantization in one pass (models/IRNp_model.py:674-680):
    `Quantization(x + (clamp(sim, 0, 1) - x).detach())`, bit-identical values, identity gradient to x."""

    def __init__(self, clamp: bool = True, quantize: bool = True):
        super().__init__()
        self.clamp, self.quantize = clamp, quantize

    def forward(self, x, simulated):
        return F_.attack_epilogue(x, simulated, self.clamp, self.quantize)


class AttackMix(nn.Module):
    """The hybrid attack of the video trainer (models/IRNcrop_model.py:357-373, as intended: the loop there adds the
    softmax weights themselves instead of the weighted images): `alpha = softmax(randn(B, K))`, the K attacked versions
    mixed per sample, then clamp_with_grad and Quantization — one pass over the K tensors instead of ~2K + 7."""

    def __init__(self, clamp: bool = True, quantize: bool = True):
        super().__init__()
        self.clamp, self.quantize = clamp, quantize

    def forward(self, attacked: Sequence[torch.Tensor], alpha: Optional[torch.Tensor] = None):
        attacked = [a[0] if isinstance(a, tuple) else a for a in attacked]
        if alpha is None:                  # the trainer's draw (:358-359), one weight vector per sample
            alpha = torch.softmax(torch.randn(attacked[0].shape[0], len(attacked), device=attacked[0].device), dim=1)
        return F_.attack_mix(attacked, alpha, self.clamp, self.quantize)


class AttackBank(nn.Module):
    """The K-way attack of the trainers (models/IRNp_model.py:609-680; the per-frame 5-way loop of
    models/IRNcrop_model.py:357-370): every layer attacks the SAME batch, each result is clamped,
    made straight-through and 8-bit quantised, and the K results are concatenated on the batch axis.
    Here each epilogue writes directly into its slice of the output (no cat, one pass per layer)."""

    def __init__(self, layers: Sequence[nn.Module], clamp: bool = True, quantize: bool = True):
        super().__init__()
        self.list = list(layers)          # plain list, like Combined (noise_layers/combined.py:10)
        self.clamp, self.quantize = clamp, quantize
        self.names = []
        self.fused = True                 # False: run every layer plainly + the stand-alone epilogue kernel
        self.shared_read = True           # 3x3-neighbourhood members read x ONCE (wm_bank3_fwd); False: one kernel each

    def forward(self, x):
        if self.fused:
            # the attack kernels write clamp + straight-through + Quantization in their own stores
            y = F_._BankFusedFn.apply(x, self.clamp, self.quantize, self.list, self.shared_read)
            self.names = [getattr(layer, "name", type(layer).__name__) for layer in self.list]
            return y
        with torch.no_grad():             # straight-through: the attacks' own graphs are never needed
            sims = []
            self.names = []
            for layer in self.list:
                y = layer(x)
                sims.append(y[0] if isinstance(y, tuple) else y)
                self.names.append(getattr(layer, "name", type(layer).__name__))
        return F_.attack_bank(x, sims, self.clamp, self.quantize)


class Splice(nn.Module):
    """`forward_image * (1 - mask) + other * mask` (models/IRNcrop_model.py:348), mask [B,1,H,W]."""

    def forward(self, image, other, mask):
        return F_.splice(image, other, mask)

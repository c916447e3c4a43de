"""torch.autograd.Function wrappers over the C ABI (include/wm_attack.h).

Every function here launches hand-written sm_100a kernels from libwmattack.so on the current
CUDA stream; PyTorch is used only to own device memory and to hook into autograd.
There is deliberately NO CPU implementation: a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from ._lib import Jpeg8Params

ROUND_ONLY_AT_0, ROUND_CUBIC, ROUND_HARD, ROUND_FOURIER = 0, 1, 2, 3
JPEG8_HARD, JPEG8_SS, JPEG8_MASK = 0, 1, 2
BILINEAR, BICUBIC = 0, 1
_MODES = {"bilinear": BILINEAR, "bicubic": BICUBIC}


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------

try:                                       # raw handle of the current stream without building a Stream object
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                     # pragma: no cover - older/newer torch without the private hook
    _raw_stream = None


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (every launch goes there)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _check_cuda(t: torch.Tensor, who: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{who}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{who}: expected a CUDA tensor — this package has no CPU fallback "
                           f"(got device {t.device})")


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.float32 else t.float()


# element types the typed entry points read / store directly (include/wm_attack.h WM_DT_*): the autocast boundary's
# cast is fused into the kernel instead of a separate .float() / .to(bfloat16) pass
DT_F32, DT_F16, DT_BF16 = 0, 1, 2
_DT_CODE = {torch.float32: DT_F32, torch.float16: DT_F16, torch.bfloat16: DT_BF16}


def _image(t: torch.Tensor, who: str, align_elems: int = 8, typed: bool = False) -> Tuple[torch.Tensor, int, int, int]:
    """Return (tensor, sb, sc, sh) of a [B,C,H,W] float32 CUDA tensor readable in place by the
    vector kernels (unit W stride, aligned strides/base); otherwise a contiguous copy.
    typed=True keeps float16 / bfloat16 tensors as they are (for entry points with a *_dtype argument)."""
    _check_cuda(t, who)
    if t.dim() != 4:
        raise ValueError(f"{who}: expected a 4-D [B,C,H,W] tensor, got shape {tuple(t.shape)}")
    if not (typed and t.dtype in _DT_CODE):
        t = _f32(t)
    sb, sc, sh, sw = t.stride()
    ok = (sw == 1 and sb % align_elems == 0 and sc % align_elems == 0 and sh % align_elems == 0
          and t.data_ptr() % (t.element_size() * align_elems) == 0 and min(sb, sc, sh) >= 0)
    if not ok:
        t = t.contiguous()
        sb, sc, sh, sw = t.stride()
    return t, sb, sc, sh


def _planes(t: torch.Tensor, who: str) -> Tuple[torch.Tensor, int, int]:
    """[B,C,H,W] viewed as N = B*C planes with one plane stride (needs sb == C*sc)."""
    _check_cuda(t, who)
    if t.dim() != 4:
        raise ValueError(f"{who}: expected a 4-D [B,C,H,W] tensor, got shape {tuple(t.shape)}")
    t = _f32(t)
    b, c, h, w = t.shape
    sb, sc, sh, sw = t.stride()
    if not (sw == 1 and (b == 1 or sb == c * sc) and sh >= w and sc >= 0):
        t = t.contiguous()
        sb, sc, sh, sw = t.stride()
    return t, sc, sh


def _typed_planes(t: torch.Tensor, who: str) -> Tuple[torch.Tensor, int, int, int]:
    """(tensor, plane stride, row stride, WM_DT_* code).  A float16 / bfloat16 image whose rows sit on 16-byte boundaries
    (W % 8 == 0, strides multiples of 8 elements) is handed to the *_typed entry points AS IT IS - the TMA ring stages the
    2-byte planes and the kernel widens them - instead of paying a .float() pass; anything else is float32 planes."""
    _check_cuda(t, who)
    if t.dim() == 4 and t.dtype in (torch.float16, torch.bfloat16):
        b, c, h, w = t.shape
        sb, sc, sh, sw = t.stride()
        if (sw == 1 and w % 8 == 0 and sh % 8 == 0 and sc % 8 == 0 and sh >= w and sc >= 0 and (b == 1 or sb == c * sc)
                and t.data_ptr() % 16 == 0):
            return t, sc, sh, _DT_CODE[t.dtype]
    t, sp, sh = _planes(t, who)
    return t, sp, sh, DT_F32


def _flat(t: torch.Tensor, who: str) -> torch.Tensor:
    _check_cuda(t, who)
    return _f32(t).contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_rng_lock = threading.Lock()
_rng_calls = 0


RNG_FROM_DEVICE = 0xFFFFFFFFFFFFFFFF
_dev_rng = {"on": False, "state": {}, "slots": []}


def device_rng(enabled: bool = True, seed: Optional[int] = None) -> None:
    """Keep the Philox state of the stochastic layers ON THE DEVICE (one {seed, next offset} pair per GPU,
    advanced by a 1-thread kernel in the launching stream) instead of in this module's host counter.
    Needed to capture Gaussian / SaltPepper / Dropout calls in a CUDA graph: with the host counter a
    replay would repeat the captured (seed, offset), i.e. the same noise.  Call once before capturing."""
    _dev_rng["on"] = bool(enabled)
    _dev_rng["state"].clear()
    _dev_rng["seed"] = seed
    if enabled and torch.cuda.is_available():
        # create the state of the current device NOW (eagerly, outside any capture): building it lazily inside a
        # captured region would be an H2D copy during capture
        _device_rng_state(torch.device("cuda", torch.cuda.current_device()))


def _device_rng_state(device) -> torch.Tensor:
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = str(device)
    st = _dev_rng["state"].get(key)
    if st is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError(f"device_rng: no generator state on {key} yet and a CUDA graph is being captured — call "
                               "functional.device_rng(True) with that device current before capturing")
        seed = _dev_rng.get("seed")
        seed = (torch.initial_seed() if seed is None else int(seed)) & 0x7FFFFFFFFFFFFFFF
        rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
        st = _dev_rng["state"][key] = torch.tensor([seed, rank << 44], dtype=torch.int64, device=device)
        torch.cuda.current_stream(device).synchronize()        # visible to every stream that uses it later
    return st


def _device_rng_slot(device, n_elems: int) -> torch.Tensor:
    st = _device_rng_state(device)
    slot = torch.empty(2, dtype=torch.int64, device=device)
    _lib.call("wm_rng_reserve", st.data_ptr(), slot.data_ptr(), (n_elems + 3) // 4 + 1, _stream())
    return slot


def next_philox_stream(n_elems: int, device=None):
    """(seed, offset) for an in-kernel Philox draw of n_elems values: seeded by
    torch.initial_seed(), advanced per call so that successive layers decorrelate.
    Per-rank decorrelation under DDP comes from each rank's own call sequence + rank offset.
    With device_rng(True) and a device given: (RNG_FROM_DEVICE, slot) where slot is a device tensor
    {seed, offset} reserved in the current stream (its address goes where the offset would)."""
    global _rng_calls
    if _dev_rng["on"] and device is not None:
        return RNG_FROM_DEVICE, _device_rng_slot(device, n_elems)
    with _rng_lock:
        off = _rng_calls
        _rng_calls += (n_elems + 3) // 4 + 1
    rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank()
    return torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, (off + (rank << 44)) & 0xFFFFFFFFFFFFFFFF


def _off(offset):
    """offset argument for the C ABI: a python int, or the address of a reserved device slot."""
    return offset.data_ptr() if torch.is_tensor(offset) else offset


# --------------------------------------------------------------------------------------
# DiffJPEG
# --------------------------------------------------------------------------------------

def quality_to_factor(quality: float) -> float:
    """utils/JPEG.py:487-498.  quality == 100 gives factor 0: upstream divides by table * 0 and returns NaN images;
    the kernels do the same (their clamps propagate NaN, tests/test_gpu_parity.py::test_nan_propagates_...)."""
    q = 5000.0 / quality if quality < 50 else 200.0 - quality * 2
    return q / 100.0


def _factor_args(factor, batch: int, device) -> Tuple[float, Optional[torch.Tensor]]:
    if torch.is_tensor(factor):
        f = factor.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        if f.numel() == 1:
            return float(f.item()), None
        if f.numel() != batch:
            raise ValueError(f"per-sample factor must have {batch} entries, got {f.numel()}")
        return 0.0, f
    return float(factor), None


class _DiffJPEGFn(torch.autograd.Function):
    """Forward-only calls (no grad needed) use wm_diffjpeg_fwd and save nothing.  When the input
    requires grad the forward saves 7 B/px (round'(q) + clamp codes) and the backward runs from gy and
    that state alone; `recompute=True` keeps the save-nothing pair (backward recomputes from x)."""

    @staticmethod
    def forward(ctx, x, factor, rounding, recompute):
        x, sb, sc, sh = _image(x, "DiffJPEG", typed=True)      # float16 / bfloat16 inputs are read as they are
        xdt = _DT_CODE[x.dtype]
        b, c, h, w = x.shape
        if c != 3:
            raise ValueError(f"DiffJPEG expects 3 channels, got {c}")
        if h % 16 or w % 16:
            raise ValueError(f"DiffJPEG needs H and W to be multiples of 16 (got {h}x{w}); the reference's "
                             "block_merging views require it (utils/JPEG.py:371-376)")
        fs, fps = _factor_args(factor, b, x.device)
        y = torch.empty((b, 3, h, w), device=x.device, dtype=torch.float32)
        need_grad = bool(ctx.needs_input_grad[0])
        ctx.mode = "none"
        if need_grad and rounding != ROUND_HARD and not recompute:
            d_y = torch.empty((b, h, w), device=x.device, dtype=torch.float32)
            d_c = torch.empty((b, 2, h // 2, w // 2), device=x.device, dtype=torch.float32)
            codes = torch.empty((b, h, w // 8), device=x.device, dtype=torch.int64)
            _lib.call("wm_diffjpeg_fwd_save", x.data_ptr(), xdt, sb, sc, sh, y.data_ptr(), d_y.data_ptr(), d_c.data_ptr(),
                      codes.data_ptr(), b, h, w, fs, _ptr(fps), rounding, _stream())
            ctx.save_for_backward(d_y, d_c, codes)
            ctx.mode = "saved"
        else:
            _lib.call("wm_diffjpeg_fwd", x.data_ptr(), xdt, sb, sc, sh, y.data_ptr(), b, h, w, fs, _ptr(fps), rounding, None, _stream())
            if need_grad and rounding != ROUND_HARD:
                ctx.save_for_backward(x, fps if fps is not None else torch.empty(0, device=x.device))
                ctx.mode = "recompute"
        ctx.meta = (fs, fps is not None, rounding, (b, 3, h, w), x.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        fs, has_fps, rounding, (b, _, h, w), xdtype = ctx.meta
        if ctx.mode == "none":                        # torch.round: zero gradient everywhere
            return torch.zeros((b, 3, h, w), device=gy.device, dtype=xdtype), None, None, None
        gy, gsb, gsc, gsh = _image(gy, "DiffJPEG.backward")
        gx = torch.empty((b, 3, h, w), device=gy.device, dtype=xdtype)     # stored in the input's element type by the kernel
        gdt = _DT_CODE[xdtype]
        if ctx.mode == "saved":
            d_y, d_c, codes = ctx.saved_tensors
            _lib.call("wm_diffjpeg_bwd_saved", gy.data_ptr(), gsb, gsc, gsh, d_y.data_ptr(), d_c.data_ptr(),
                      codes.data_ptr(), gx.data_ptr(), gdt, b, h, w, _stream())
        else:
            x, fps_t = ctx.saved_tensors
            sb, sc, sh, _ = x.stride()
            _lib.call("wm_diffjpeg_bwd", x.data_ptr(), gdt, sb, sc, sh, gy.data_ptr(), gsb, gsc, gsh, gx.data_ptr(), gdt,
                      b, h, w, fs, _ptr(fps_t) if has_fps else None, rounding, _stream())
        return gx, None, None, None


def diffjpeg(x: torch.Tensor, factor, rounding: int = ROUND_ONLY_AT_0, recompute: bool = False) -> torch.Tensor:
    """Fused DiffJPEG (utils/JPEG.py:535-540).  `factor` = quality_to_factor(quality), a python
    float or a per-sample tensor [B].  recompute=True: the backward recomputes the forward from x
    instead of reading 7 B/px of saved state (lower memory, slower)."""
    return _DiffJPEGFn.apply(x, factor, rounding, recompute)


def diffjpeg_compress(x: torch.Tensor, factor, rounding: int = ROUND_HARD):
    """compress_jpeg.forward (utils/JPEG.py:279-291) -> (y, cb, cr) as [B, nblk, 8, 8]. No autograd."""
    x, sb, sc, sh = _image(x.detach(), "DiffJPEG.compress")
    b, c, h, w = x.shape
    if c != 3 or h % 16 or w % 16:
        raise ValueError(f"compress expects [B,3,H,W] with H,W multiples of 16, got {tuple(x.shape)}")
    fs, fps = _factor_args(factor, b, x.device)
    cy = torch.empty((b, h * w // 64, 8, 8), device=x.device, dtype=torch.float32)
    ccb = torch.empty((b, h * w // 256, 8, 8), device=x.device, dtype=torch.float32)
    ccr = torch.empty_like(ccb)
    _lib.call("wm_diffjpeg_compress", x.data_ptr(), sb, sc, sh, cy.data_ptr(), ccb.data_ptr(), ccr.data_ptr(),
              b, h, w, fs, _ptr(fps), rounding, _stream())
    return cy, ccb, ccr


def diffjpeg_decompress(y, cb, cr, height: int, width: int, factor) -> torch.Tensor:
    """decompress_jpeg.forward (utils/JPEG.py:452-469). No autograd."""
    for t in (y, cb, cr):
        _check_cuda(t, "DiffJPEG.decompress")
    y, cb, cr = (_f32(t.detach()).contiguous() for t in (y, cb, cr))
    b = y.shape[0]
    if height % 16 or width % 16 or y.numel() != b * height * width or cb.numel() != y.numel() // 4:
        raise ValueError("decompress: coefficient tensors do not match height/width")
    fs, fps = _factor_args(factor, b, y.device)
    out = torch.empty((b, 3, height, width), device=y.device, dtype=torch.float32)
    _lib.call("wm_diffjpeg_decompress", y.data_ptr(), cb.data_ptr(), cr.data_ptr(), out.data_ptr(),
              b, height, width, fs, _ptr(fps), _stream())
    return out


# --------------------------------------------------------------------------------------
# 8x8-unit JPEG family
# --------------------------------------------------------------------------------------

def make_jpeg8_params(fwd_color: Sequence[float], inv_color: Sequence[float], table: np.ndarray,
                      variant: int, subsample: int) -> Jpeg8Params:
    p = Jpeg8Params()
    for i in range(9):
        p.fwd_color[i] = float(fwd_color[i])
        p.inv_color[i] = float(inv_color[i])
    tab = np.asarray(table, dtype=np.float32).reshape(3, 64)
    for c in range(3):
        for i in range(64):
            p.table[c][i] = float(tab[c, i])
    p.variant, p.subsample = int(variant), int(subsample)
    return p


class _Jpeg8Fn(torch.autograd.Function):
    """float16 / bfloat16 images are read as they are, and the gradient is stored in that type, whenever the call takes
    the vector path of the two-threads-per-block kernel (W % 8 == 0, no in-block subsampling; JpegSS needs its
    saved-state pair for a typed gradient); any other geometry converts to float32 first."""

    @staticmethod
    def forward(ctx, x, params):
        typed = x.dtype in (torch.float16, torch.bfloat16) and x.shape[-1] % 8 == 0 and params.subsample == 0
        x, sb, sc, sh = _image(x, "jpeg8", typed=typed)
        xdt = _DT_CODE[x.dtype]
        b, c, h, w = x.shape
        if c != 3:
            raise ValueError(f"JPEG layers expect 3 channels, got {c}")
        y = torch.empty((b, 3, h, w), device=x.device, dtype=torch.float32)
        ctx.params = params
        ctx.shape = (b, h, w)
        ctx.saved_d = False
        ctx.xdtype = x.dtype
        if params.variant == JPEG8_SS and ctx.needs_input_grad[0] and w % 8 == 0 and params.subsample == 0:
            # training pair: save ss'(q) (12 B/px); the backward then needs neither x nor a recompute
            d = torch.empty((b, 3, (h + 7) // 8 * 8, w), device=x.device, dtype=torch.float32)
            _lib.call("wm_jpeg8_fwd_save", x.data_ptr(), xdt, sb, sc, sh, y.data_ptr(), d.data_ptr(), b, h, w,
                      C.byref(params), _stream())
            ctx.save_for_backward(d)
            ctx.saved_d = True
            return y
        _lib.call("wm_jpeg8_fwd", x.data_ptr(), xdt, sb, sc, sh, y.data_ptr(), b, h, w, C.byref(params), None, _stream())
        if params.variant == JPEG8_SS:
            ctx.save_for_backward(x)                 # ragged / subsampled JpegSS: float32 here (typed is False)
        return y

    @staticmethod
    def backward(ctx, gy):
        p = ctx.params
        gdt = _DT_CODE[ctx.xdtype]
        if ctx.saved_d:
            (d,) = ctx.saved_tensors
            b, h, w = ctx.shape
            gy, gsb, gsc, gsh = _image(gy, "jpeg8.backward")
            gx = torch.empty((b, 3, h, w), device=gy.device, dtype=ctx.xdtype)
            _lib.call("wm_jpeg8_bwd_saved", gy.data_ptr(), gsb, gsc, gsh, d.data_ptr(), gx.data_ptr(), gdt, b, h, w,
                      C.byref(p), _stream())
            return gx, None
        if p.variant == JPEG8_SS:
            (x,) = ctx.saved_tensors
            b, _, h, w = x.shape
            sb, sc, sh, _ = x.stride()
            xp = x.data_ptr()
        else:
            b, h, w = ctx.shape
            sb = sc = sh = 0
            xp = None
        gy, gsb, gsc, gsh = _image(gy, "jpeg8.backward")
        gx = torch.empty((b, 3, h, w), device=gy.device, dtype=ctx.xdtype)
        _lib.call("wm_jpeg8_bwd", xp, sb, sc, sh, gy.data_ptr(), gsb, gsc, gsh, gx.data_ptr(), gdt, b, h, w,
                  C.byref(p), _stream())
        return gx, None


def jpeg8(x: torch.Tensor, params: Jpeg8Params) -> torch.Tensor:
    return _Jpeg8Fn.apply(x, params)


def jpeg8_quantised(x: torch.Tensor, params: Jpeg8Params) -> torch.Tensor:
    """std_quantization output (noise_layers/jpeg.py:52-82) as [B,3,Hp,Wp]. No autograd."""
    x, sb, sc, sh = _image(x.detach(), "jpeg8_quantised")
    b, c, h, w = x.shape
    hp, wp = (h + 7) // 8 * 8, (w + 7) // 8 * 8
    coef = torch.empty((b, 3, hp, wp), device=x.device, dtype=torch.float32)
    _lib.call("wm_jpeg8_quantised", x.data_ptr(), sb, sc, sh, coef.data_ptr(), b, h, w, C.byref(params), _stream())
    return coef


# --------------------------------------------------------------------------------------
# Gaussian blur / median
# --------------------------------------------------------------------------------------

_TAPS_CACHE: dict = {}


def _taps_array(taps: Sequence[float]):
    """ctypes float array of the taps (read by the launcher on the host, copied into kernel arguments);
    cached per tap tuple — building it costs several microseconds per call."""
    key = tuple(taps)
    arr = _TAPS_CACHE.get(key)
    if arr is None:
        if len(_TAPS_CACHE) > 256:
            _TAPS_CACHE.clear()
        arr = _TAPS_CACHE[key] = (C.c_float * len(key))(*[float(t) for t in key])
    return arr


class _BlurFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, taps, border):
        xdtype = x.dtype
        ring = border == 0 and len(taps) in (3, 5, 7)         # the TMA ring kernels: the only ones with typed planes
        x, sp, sh, dt = _typed_planes(x, "gaussian blur") if ring else (*_planes(x, "gaussian blur"), DT_F32)
        b, c, h, w = x.shape
        y = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
        arr = _taps_array(taps)
        if dt != DT_F32:       # float16 / bfloat16 planes staged as they are
            _lib.call("wm_gaussblur_typed", x.data_ptr(), dt, sp, sh, y.data_ptr(), DT_F32, b * c, h, w, arr, len(taps), _stream())
        else:
            _lib.call("wm_gaussblur", x.data_ptr(), sp, sh, y.data_ptr(), b * c, h, w, arr, len(taps), border, 0, None, _stream())
        ctx.meta = (tuple(taps), border, xdtype if ring else torch.float32)
        return y

    @staticmethod
    def backward(ctx, gy):
        taps, border, xdtype = ctx.meta
        gy, sp, sh = _planes(gy, "gaussian blur backward")
        b, c, h, w = gy.shape
        if xdtype in (torch.float16, torch.bfloat16) and w % 4 == 0 and sh % 4 == 0 and sp % 4 == 0 and gy.data_ptr() % 16 == 0:
            gx = torch.empty((b, c, h, w), device=gy.device, dtype=xdtype)      # the gradient leaves in the image's type
            _lib.call("wm_gaussblur_typed", gy.data_ptr(), DT_F32, sp, sh, gx.data_ptr(), _DT_CODE[xdtype], b * c, h, w,
                      _taps_array(taps), len(taps), _stream())
            return gx, None, None
        gx = torch.empty((b, c, h, w), device=gy.device, dtype=torch.float32)
        _lib.call("wm_gaussblur", gy.data_ptr(), sp, sh, gx.data_ptr(), b * c, h, w, _taps_array(taps), len(taps),
                  border, 1, None, _stream())
        return gx, None, None


def gaussian_blur(x: torch.Tensor, taps: Sequence[float], border: int = 0) -> torch.Tensor:
    """Separable blur with normalised 1-D `taps`; border 0 = zero pad, 1 = reflect."""
    return _BlurFn.apply(x, taps if isinstance(taps, tuple) else tuple(float(t) for t in taps), border)


def _idx_plane(b: int, c: int, h: int, w: int, device) -> torch.Tensor:
    """Arg-median plane with 16-byte rows (the bytes past W are padding): the backward's idx ring stays on TMA for any W."""
    return torch.empty((b, c, h, -(-w // 16) * 16), device=device, dtype=torch.uint8)


class _MedianFn(torch.autograd.Function):
    """Any geometry goes to the ring kernels: TMA-fed when the rows sit on 16-byte boundaries, cp.async-fed otherwise
    (W % 4 != 0, odd strides) — no padding copy either way."""

    @staticmethod
    def forward(ctx, x, k):
        need_idx = bool(ctx.needs_input_grad[0])
        ctx.xdtype = x.dtype
        x, sp, sh, dt = _typed_planes(x, "median blur")
        b, c, h, w = x.shape
        y = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
        idx = _idx_plane(b, c, h, w, x.device) if need_idx else None
        if dt != DT_F32:       # float16 / bfloat16 planes staged as they are (exact widening: same median, same position)
            _lib.call("wm_median_fwd_typed", x.data_ptr(), dt, sp, sh, y.data_ptr(), _ptr(idx), idx.shape[-1] if need_idx else 0,
                      b * c, h, w, k, _stream())
        else:
            _lib.call("wm_median_fwd", x.data_ptr(), sp, sh, y.data_ptr(), _ptr(idx), idx.shape[-1] if need_idx else 0,
                      b * c, h, w, k, None, _stream())
        ctx.k = k
        if need_idx:
            ctx.save_for_backward(idx)
        return y

    @staticmethod
    def backward(ctx, gy):
        (idx,) = ctx.saved_tensors
        gy = _flat(gy, "median blur backward")
        b, c, h, w = gy.shape
        if ctx.xdtype in (torch.float16, torch.bfloat16) and w % 4 == 0:
            gx = torch.empty((b, c, h, w), device=gy.device, dtype=ctx.xdtype)  # the gradient leaves in the image's type
            _lib.call("wm_median_bwd_typed", gy.data_ptr(), idx.data_ptr(), idx.shape[-1], gx.data_ptr(), _DT_CODE[ctx.xdtype],
                      b * c, h, w, ctx.k, _stream())
            return gx, None
        gx = torch.empty_like(gy)
        _lib.call("wm_median_bwd", gy.data_ptr(), idx.data_ptr(), idx.shape[-1], gx.data_ptr(), b * c, h, w, ctx.k, _stream())
        return gx, None


def median_blur(x: torch.Tensor, k: int) -> torch.Tensor:
    if k not in (3, 5):
        raise ValueError(f"median_blur supports kernel sizes 3 and 5, got {k}")
    return _MedianFn.apply(x, k)


def median_blur_with_index(x: torch.Tensor, k: int):
    """(values, uint8 arg-median index plane) — test/inspection helper, no autograd."""
    x, sp, sh = _planes(x.detach(), "median blur")
    b, c, h, w = x.shape
    y = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
    idx = _idx_plane(b, c, h, w, x.device)
    _lib.call("wm_median_fwd", x.data_ptr(), sp, sh, y.data_ptr(), idx.data_ptr(), idx.shape[-1], b * c, h, w, k, None, _stream())
    return y, idx[..., :w]


# --------------------------------------------------------------------------------------
# Elementwise attacks
# --------------------------------------------------------------------------------------

class _GaussNoiseFn(torch.autograd.Function):
    """Clamped layer with grad: the forward saves the clamp's 1-bit pass mask (0.4 B/px) and the backward is
    a masked copy of gy.  regen=True keeps the save-x pair whose backward regenerates the noise by Philox.
    float16 / bfloat16 images are read as they are; the mask pair stores the gradient in that type."""

    @staticmethod
    def forward(ctx, x, mean, std, clamp, noise, seed, offset, regen):
        _check_cuda(x, "gaussian noise")
        typed = x.dtype in (torch.float16, torch.bfloat16) and not (clamp and regen)
        x = x.contiguous() if typed else _flat(x, "gaussian noise")
        xdt = _DT_CODE[x.dtype]
        inj = _flat(noise, "gaussian noise (injected)") if noise is not None else None
        if inj is not None and inj.shape != x.shape:
            raise ValueError("injected noise must have the input's shape")
        y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        ctx.mode = "identity"
        ctx.xdtype = x.dtype
        if clamp and ctx.needs_input_grad[0] and not regen:
            n = x.numel()
            mask = torch.empty(4 * ((n + 127) // 128), device=x.device, dtype=torch.int32)
            _lib.call("wm_gaussnoise_fwd_mask", x.data_ptr(), xdt, y.data_ptr(), mask.data_ptr(), n, mean, std, seed, _off(offset),
                      _ptr(inj), _stream())
            ctx.save_for_backward(mask)
            ctx.mode = "mask"
            return y
        _lib.call("wm_gaussnoise_fwd", x.data_ptr(), xdt, y.data_ptr(), x.numel(), mean, std, int(clamp), seed, _off(offset),
                  _ptr(inj), None, _stream())
        ctx.meta = (mean, std, int(clamp), seed, offset)
        if clamp:
            ctx.save_for_backward(x, inj if inj is not None else torch.empty(0, device=x.device))
            ctx.mode = "regen"
        ctx.has_inj = inj is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = _flat(gy, "gaussian noise backward")
        if ctx.mode == "identity":
            return gy, None, None, None, None, None, None, None
        if ctx.mode == "mask":
            (mask,) = ctx.saved_tensors
            gx = torch.empty(gy.shape, device=gy.device, dtype=ctx.xdtype)
            _lib.call("wm_gaussnoise_bwd_mask", gy.data_ptr(), mask.data_ptr(), gx.data_ptr(), _DT_CODE[ctx.xdtype], gy.numel(), _stream())
            return gx, None, None, None, None, None, None, None
        gx = torch.empty_like(gy)
        mean, std, clamp, seed, offset = ctx.meta
        x, inj = ctx.saved_tensors
        _lib.call("wm_gaussnoise_bwd", x.data_ptr(), gy.data_ptr(), gx.data_ptr(), gy.numel(), mean, std, clamp,
                  seed, _off(offset), inj.data_ptr() if ctx.has_inj else None, _stream())
        return gx, None, None, None, None, None, None, None


def gaussian_noise(x, mean: float = 0.0, std: float = 0.05, clamp: bool = True, noise=None, regen: bool = False):
    """regen=True: save x and regenerate the noise in the backward instead of saving the 1-bit clamp mask."""
    seed, offset = next_philox_stream(x.numel(), x.device) if noise is None else (0, 0)
    return _GaussNoiseFn.apply(x, float(mean), float(std), clamp, noise, seed, offset, regen)


class _SaltPepperFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, prob, rdn, seed, offset):
        x = _flat(x, "salt & pepper")
        inj = _flat(rdn, "salt & pepper (injected)") if rdn is not None else None
        y = torch.empty_like(x)
        _lib.call("wm_saltpepper_fwd", x.data_ptr(), y.data_ptr(), x.numel(), prob, seed, _off(offset), _ptr(inj), _stream())
        ctx.meta = (prob, seed, offset)
        ctx.inj = inj
        return y

    @staticmethod
    def backward(ctx, gy):
        prob, seed, offset = ctx.meta
        gy = _flat(gy, "salt & pepper backward")
        gx = torch.empty_like(gy)
        _lib.call("wm_saltpepper_bwd", gy.data_ptr(), gx.data_ptr(), gy.numel(), prob, seed, _off(offset), _ptr(ctx.inj), _stream())
        return gx, None, None, None, None


def salt_pepper(x, prob: float, rdn=None):
    seed, offset = next_philox_stream(x.numel(), x.device) if rdn is None else (0, 0)
    return _SaltPepperFn.apply(x, float(prob), rdn, seed, offset)


class _DropoutElemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, cover, prob, rdn, seed, offset):
        image = _flat(image, "dropout")
        cover = _flat(cover, "dropout")
        if image.shape != cover.shape:
            raise ValueError("image and cover must have the same shape")
        inj = _flat(rdn, "dropout (injected)") if rdn is not None else None
        y = torch.empty_like(image)
        _lib.call("wm_dropout_elem_fwd", image.data_ptr(), cover.data_ptr(), y.data_ptr(), image.numel(), prob,
                  seed, _off(offset), _ptr(inj), _stream())
        ctx.meta = (prob, seed, offset)
        ctx.inj = inj
        return y

    @staticmethod
    def backward(ctx, gy):
        prob, seed, offset = ctx.meta
        gy = _flat(gy, "dropout backward")
        gi = torch.empty_like(gy) if ctx.needs_input_grad[0] else None
        gc = torch.empty_like(gy) if ctx.needs_input_grad[1] else None
        if gi is not None or gc is not None:
            _lib.call("wm_dropout_elem_bwd", gy.data_ptr(), _ptr(gi), _ptr(gc), gy.numel(), prob, seed, _off(offset),
                      _ptr(ctx.inj), _stream())
        return gi, gc, None, None, None, None


def dropout_elementwise(image, cover, prob: float, rdn=None):
    seed, offset = next_philox_stream(image.numel(), image.device) if rdn is None else (0, 0)
    return _DropoutElemFn.apply(image, cover, float(prob), rdn, seed, offset)


class _DropoutMaskFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, noised, cover, mask_hw):
        noised = _flat(noised, "dropout")
        cover = _flat(cover, "dropout")
        mask_hw = _flat(mask_hw, "dropout mask")
        b, c, h, w = noised.shape
        if mask_hw.numel() != h * w or cover.shape != noised.shape:
            raise ValueError("mask must be [H,W] and cover must match the noised image")
        y = torch.empty_like(noised)
        _lib.call("wm_dropout_mask_fwd", noised.data_ptr(), cover.data_ptr(), mask_hw.data_ptr(), y.data_ptr(),
                  b * c, h * w, _stream())
        ctx.save_for_backward(mask_hw)
        return y

    @staticmethod
    def backward(ctx, gy):
        (mask_hw,) = ctx.saved_tensors
        gy = _flat(gy, "dropout backward")
        b, c, h, w = gy.shape
        gn = torch.empty_like(gy) if ctx.needs_input_grad[0] else None
        gc = torch.empty_like(gy) if ctx.needs_input_grad[1] else None
        if gn is not None or gc is not None:
            _lib.call("wm_dropout_mask_bwd", gy.data_ptr(), mask_hw.data_ptr(), _ptr(gn), _ptr(gc), b * c, h * w, _stream())
        return gn, gc, None


def dropout_mask(noised, cover, mask_hw):
    return _DropoutMaskFn.apply(noised, cover, mask_hw)


def bernoulli_mask(h: int, w: int, keep: float, device) -> torch.Tensor:
    m = torch.empty((h, w), device=device, dtype=torch.float32)
    seed, offset = next_philox_stream(h * w, device)
    _lib.call("wm_bernoulli_mask", m.data_ptr(), h * w, float(keep), seed, _off(offset), _stream())
    return m


class _Quantize8Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, clamp01):
        x = _flat(x, "quantization")
        y = torch.empty_like(x)
        _lib.call("wm_quantize8_fwd", x.data_ptr(), y.data_ptr(), x.numel(), int(clamp01), _stream())
        return y

    @staticmethod
    def backward(ctx, gy):
        return gy, None      # straight-through (models/modules/Quantization.py:13-14)


def quantize8(x, clamp01: bool = False):
    return _Quantize8Fn.apply(x, clamp01)


class _CropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, cover, box):
        image = _flat(image, "cropout")
        cover = _flat(cover, "cropout")
        b, c, h, w = image.shape
        y = torch.empty_like(image)
        _lib.call("wm_cropout_fwd", image.data_ptr(), cover.data_ptr(), y.data_ptr(), b * c, h, w, *[int(v) for v in box], _stream())
        ctx.box = box
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = _flat(gy, "cropout backward")
        b, c, h, w = gy.shape
        z = torch.zeros_like(gy)
        gi = torch.empty_like(gy)
        gc = torch.empty_like(gy)
        box = [int(v) for v in ctx.box]
        _lib.call("wm_cropout_fwd", gy.data_ptr(), z.data_ptr(), gi.data_ptr(), b * c, h, w, *box, _stream())
        _lib.call("wm_cropout_fwd", z.data_ptr(), gy.data_ptr(), gc.data_ptr(), b * c, h, w, *box, _stream())
        return gi, gc, None


def cropout(image, cover, box):
    return _CropoutFn.apply(image, cover, tuple(box))


# --------------------------------------------------------------------------------------
# interpolation: Resize / Crop
# --------------------------------------------------------------------------------------

def _crop_fast(n, window, out_hw, src_hw, mode, clamp) -> bool:
    """Up-scaling crop geometries served by wm_cropresize_* (TMA source box, banded tables per call)."""
    h0, w0, hin, win = window
    return (not clamp) and src_hw[1] % 4 == 0 and bool(
        _lib.load().wm_cropresize_ok(hin, win, out_hw[0], out_hw[1], n, mode))


def _crop_tables(device, window, out_hw, mode):
    words = int(_lib.load().wm_cropresize_table_words(window[2], window[3], out_hw[0], out_hw[1], mode))
    return torch.empty(words, device=device, dtype=torch.int32)


def _interp_fwd(x, sp, sh, n, window, out_hw, mode, clamp, want_mask=False):
    h0, w0, hin, win = window
    y = torch.empty((n, out_hw[0], out_hw[1]), device=x.device, dtype=torch.float32)
    src_hw = x.shape[-2:]
    if (not want_mask and _crop_fast(n, window, out_hw, src_hw, mode, clamp) and sp % 4 == 0 and sh % 4 == 0
            and x.data_ptr() % 16 == 0):
        tables = _crop_tables(x.device, window, out_hw, mode)
        _lib.call("wm_cropresize_fwd", x.data_ptr(), sp, sh, src_hw[0], src_hw[1], h0, w0, hin, win, y.data_ptr(), n,
                  out_hw[0], out_hw[1], mode, tables.data_ptr(), _stream())
        return y, None
    mask = None
    if want_mask:
        mask = torch.empty((n, out_hw[0], (out_hw[1] + 31) // 32), device=x.device, dtype=torch.int32)
    _lib.call("wm_interp_fwd", x.data_ptr(), sp, sh, src_hw[0], src_hw[1], h0, w0, hin, win, y.data_ptr(), n, out_hw[0],
              out_hw[1], mode, int(clamp), _ptr(mask), _stream())
    return y, mask


def _interp_bwd(gy, mask, n, out_hw, src_hw, window, mode):
    h0, w0, hin, win = window
    gx = torch.empty((n, src_hw[0], src_hw[1]), device=gy.device, dtype=torch.float32)
    if mask is None and _crop_fast(n, window, out_hw, src_hw, mode, False):
        tables = _crop_tables(gy.device, window, out_hw, mode)
        _lib.call("wm_cropresize_bwd", gy.data_ptr(), gx.data_ptr(), src_hw[0], src_hw[1], h0, w0, hin, win, n,
                  out_hw[0], out_hw[1], mode, tables.data_ptr(), _stream())
        return gx
    ws = None
    if not _lib.load().wm_interp_is_tiled(hin, win, out_hw[0], out_hw[1], n):
        ws = torch.empty((n, hin, out_hw[1]), device=gy.device, dtype=torch.float32)
    _lib.call("wm_interp_bwd", gy.data_ptr(), None, _ptr(mask), n, out_hw[0], out_hw[1], gx.data_ptr(),
              src_hw[0], src_hw[1], h0, w0, hin, win, mode, _ptr(ws), _stream())
    return gx


class _InterpFn(torch.autograd.Function):
    """y = interpolate(x[:, :, h0:h0+hin, w0:w0+win], size=out_hw) [clamped to 0..1].
    Linear apart from the clamp, whose pass-through mask is saved as 1 bit per value."""

    @staticmethod
    def forward(ctx, x, window, out_hw, mode, clamp):
        need_grad = bool(ctx.needs_input_grad[0])
        x, sp, sh = _planes(x, "interpolate")
        b, c, h, w = x.shape
        n = b * c
        y, mask = _interp_fwd(x, sp, sh, n, window, out_hw, mode, clamp, want_mask=clamp and need_grad)
        ctx.meta = (window, tuple(out_hw), mode, (b, c, h, w))
        ctx.save_for_backward(mask if mask is not None else torch.empty(0, device=x.device))
        ctx.has_mask = mask is not None
        return y.view(b, c, out_hw[0], out_hw[1])

    @staticmethod
    def backward(ctx, gy):
        window, out_hw, mode, (b, c, h, w) = ctx.meta
        (mask,) = ctx.saved_tensors
        gy = _flat(gy, "interpolate backward")
        gx = _interp_bwd(gy, mask if ctx.has_mask else None, b * c, out_hw, (h, w), window, mode)
        return gx.view(b, c, h, w), None, None, None, None


def interpolate(x, size, mode: str = "bilinear", window=None, clamp: bool = False):
    """F.interpolate(x[window], size=size, mode=mode, align_corners=False) on our kernels."""
    h, w = x.shape[2:]
    if window is None:
        window = (0, 0, h, w)
    h0, w0, hin, win = (int(v) for v in window)
    if h0 < 0 or w0 < 0 or hin <= 0 or win <= 0 or h0 + hin > h or w0 + win > w:
        # the kernels read the window in place: an out-of-range rectangle would be an out-of-bounds device read
        raise ValueError(f"interpolate: window (h0={h0}, w0={w0}, h={hin}, w={win}) does not lie inside the "
                         f"{h}x{w} source (use slice_window() for Python-slice clamping)")
    return _InterpFn.apply(x, (h0, w0, hin, win), (int(size[0]), int(size[1])), _MODES[mode], clamp)


def slice_window(shape, h_start, h_end, w_start, w_end):
    """(h0, w0, hin, win) of image[:, :, h_start:h_end, w_start:w_end] with Python's slice semantics
    (silent clamping, negative indices) — what the reference's slicing does with a caller-supplied apex
    (noise_layers/crop.py:45, models/IRNcrop_model.py:541-543 reuses one apex across tensors)."""
    h0, h1, _ = slice(int(h_start), int(h_end)).indices(shape[2])
    w0, w1, _ = slice(int(w_start), int(w_end)).indices(shape[3])
    if h1 <= h0 or w1 <= w0:
        raise ValueError(f"crop rectangle [{h_start}:{h_end}, {w_start}:{w_end}] is empty on a "
                         f"{shape[2]}x{shape[3]} image")
    return h0, w0, h1 - h0, w1 - w0


_RESIZE_TABLES: dict = {}      # (device, H, W, Hm, Wm, mode) -> device tensor of band tables (tiny, LRU-capped)


_RESIZE_UNFUSED: set = set()   # geometries whose bands were found NOT to fit the fused kernel's windows


def _resize_tables(device, h, w, mid_hw, mode):
    """Band tables of one geometry, built ONCE and proven before first use: wm_resize_tables also runs, for both
    directions, the fused kernel's window arithmetic over every tile position and sets a flag in the workspace's last
    word if a band of U*D / its transpose would not fit (a weight outside a window would be dropped silently).  One
    host sync per NEW geometry (a few small kernels); a geometry that does not fit is remembered and served by the
    two-call path (returns None).  Completed (synchronised) before it enters the cache, so any stream may use it;
    never built during CUDA-graph capture."""
    key = (str(device), h, w, mid_hw[0], mid_hw[1], mode)
    t = _RESIZE_TABLES.get(key)
    if t is not None:
        return t
    if key in _RESIZE_UNFUSED:
        return None
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError(f"resize geometry {h}x{w} -> {mid_hw[0]}x{mid_hw[1]} has no cached band tables and a CUDA graph is "
                           "being captured: run the layer once eagerly with this ratio before capturing")
    if len(_RESIZE_TABLES) >= 1024:
        _RESIZE_TABLES.pop(next(iter(_RESIZE_TABLES)))
    n = int(_lib.load().wm_resize_table_floats(h, w, mid_hw[0], mid_hw[1]))
    t = torch.empty(n, device=device, dtype=torch.float32)
    _lib.call("wm_resize_tables", t.data_ptr(), h, w, mid_hw[0], mid_hw[1], mode, _stream())
    if int(t[-4:].view(torch.int32)[0].item()) != 0:           # the one sync of this geometry
        _RESIZE_UNFUSED.add(key)
        return None
    _RESIZE_TABLES[key] = t
    return t


class _ResizeFusedFn(torch.autograd.Function):
    """clamp(up(down(x)), 0, 1) in one kernel; the clamp pass-through mask is saved as 1 bit/value."""

    @staticmethod
    def forward(ctx, x, mid_hw, mode):
        need_grad = bool(ctx.needs_input_grad[0])
        xdtype = x.dtype
        x, sp, sh, dt = _typed_planes(x, "resize")
        b, c, h, w = x.shape
        n = b * c
        tables = _resize_tables(x.device, h, w, mid_hw, mode)       # proven by resize_roundtrip before dispatching here
        y = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
        mask = torch.empty((n, h, 4 * ((w + 127) // 128)), device=x.device, dtype=torch.int32) if need_grad else None
        if dt != DT_F32:       # float16 / bfloat16 planes staged as they are
            _lib.call("wm_resize_fwd_typed", x.data_ptr(), dt, sp, sh, y.data_ptr(), n, h, w, mid_hw[0], mid_hw[1], mode,
                      _ptr(mask), tables.data_ptr(), _stream())
        else:
            _lib.call("wm_resize_fwd", x.data_ptr(), sp, sh, y.data_ptr(), n, h, w, mid_hw[0], mid_hw[1], mode,
                      _ptr(mask), tables.data_ptr(), None, _stream())
        ctx.meta = (tuple(mid_hw), mode, (b, c, h, w))
        ctx.xdtype = xdtype
        if need_grad:
            ctx.save_for_backward(mask, tables)
        return y

    @staticmethod
    def backward(ctx, gy):
        mid_hw, mode, (b, c, h, w) = ctx.meta
        mask, tables = ctx.saved_tensors
        gy = _flat(gy, "resize backward")
        n = b * c
        if ctx.xdtype in (torch.float16, torch.bfloat16) and w % 4 == 0:
            gx = torch.empty((b, c, h, w), device=gy.device, dtype=ctx.xdtype)  # the gradient leaves in the image's type
            _lib.call("wm_resize_bwd_typed", gy.data_ptr(), mask.data_ptr(), gx.data_ptr(), _DT_CODE[ctx.xdtype], n, h, w,
                      mid_hw[0], mid_hw[1], mode, tables.data_ptr(), _stream())
            return gx, None, None
        gx = torch.empty((b, c, h, w), device=gy.device, dtype=torch.float32)
        _lib.call("wm_resize_bwd", gy.data_ptr(), mask.data_ptr(), gx.data_ptr(), n, h, w, mid_hw[0], mid_hw[1], mode,
                  tables.data_ptr(), _stream())
        return gx, None, None


def resize_roundtrip(x, mid_hw, mode: str = "bicubic"):
    """Resize.forward arithmetic (noise_layers/resize.py:38-53)."""
    h, w = x.shape[2:]
    mid_hw = (int(mid_hw[0]), int(mid_hw[1]))
    n = x.shape[0] * x.shape[1]
    # any width / stride: rows off the 16-byte grid take the kernel's cp.async-fed instantiation (decided by the library)
    if _lib.load().wm_resize_is_fused(h, w, mid_hw[0], mid_hw[1], n) \
            and _resize_tables(x.device, h, w, mid_hw, _MODES[mode]) is not None:
        return _ResizeFusedFn.apply(x, mid_hw, _MODES[mode])
    mid = interpolate(x, mid_hw, mode)
    return interpolate(mid, (h, w), mode, clamp=True)


# --------------------------------------------------------------------------------------
# neighbours of the attack layer: post-attack epilogue, K-way bank, tamper/splice
# --------------------------------------------------------------------------------------

class _EpilogueFn(torch.autograd.Function):
    """out = Quantization(x + (clamp(sim,0,1) - x).detach())  (models/IRNp_model.py:674-680).
    Straight-through: d out / d x = I, nothing flows into `sim`."""

    @staticmethod
    def forward(ctx, x, sim, clamp, quantize, out):
        x = _flat(x, "attack epilogue")
        sim = _flat(sim.detach(), "attack epilogue")
        if sim.shape != x.shape:
            raise ValueError(f"attack epilogue: shape mismatch {tuple(x.shape)} vs {tuple(sim.shape)}")
        if out is None:
            out = torch.empty_like(x)
        elif out.shape != x.shape or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError("attack epilogue: `out` must be a contiguous fp32 tensor (or slice) of x's shape")
        _lib.call("wm_attack_epilogue_fwd", x.data_ptr(), sim.data_ptr(), out.data_ptr(), x.numel(), int(clamp),
                  int(quantize), _stream())
        return out

    @staticmethod
    def backward(ctx, gy):
        return gy, None, None, None, None


def attack_epilogue(x, sim, clamp: bool = True, quantize: bool = True, out=None):
    return _EpilogueFn.apply(x, sim, clamp, quantize, out)


class _BankFn(torch.autograd.Function):
    """K attacked variants of the same batch, each finished by the epilogue and written straight
    into its slice of one [K*B, C, H, W] tensor (no torch.cat).  Backward: gx = sum_k gy_k."""

    @staticmethod
    def forward(ctx, x, clamp, quantize, *sims):
        x = _flat(x, "attack bank")
        k = len(sims)
        out = torch.empty((k * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
        n = x.numel()
        for i, sim in enumerate(sims):
            sim = _flat(sim.detach(), "attack bank")
            _lib.call("wm_attack_epilogue_fwd", x.data_ptr(), sim.data_ptr(), out.data_ptr() + 4 * n * i, n, int(clamp),
                      int(quantize), _stream())
        ctx.meta = (k, tuple(x.shape))
        return out

    @staticmethod
    def backward(ctx, gy):
        k, shape = ctx.meta
        gy = _flat(gy, "attack bank backward")
        gx = torch.empty(shape, device=gy.device, dtype=torch.float32)
        _lib.call("wm_slice_sum", gy.data_ptr(), gx.data_ptr(), gx.numel(), k, _stream())
        return (gx, None, None) + (None,) * k


def attack_bank(x, sims, clamp: bool = True, quantize: bool = True):
    return _BankFn.apply(x, clamp, quantize, *sims)


class _MixFn(torch.autograd.Function):
    """Convex mix of K attacked versions + clamp_with_grad + Quantization in one pass (wm_mix_fwd); the backward
    hands alpha[b, k] * gy to every member that needs a gradient (wm_mix_bwd)."""

    @staticmethod
    def forward(ctx, alpha, clamp, quantize, *ys):
        k = len(ys)
        if not 1 <= k <= 8:
            raise ValueError(f"attack mix: 1..8 members supported, got {k}")
        ys = [_flat(y, "attack mix") for y in ys]
        shape = ys[0].shape
        if any(y.shape != shape for y in ys):
            raise ValueError("attack mix: all members must have the same shape")
        b = shape[0]
        alpha = _flat(alpha.detach(), "attack mix weights")
        if tuple(alpha.shape) != (b, k):
            raise ValueError(f"attack mix: alpha must be [B, K] = {(b, k)}, got {tuple(alpha.shape)}")
        out = torch.empty(shape, device=ys[0].device, dtype=torch.float32)
        desc = _lib.MixDesc()
        for i, y in enumerate(ys):
            desc.t[i] = y.data_ptr()
        desc.K, desc.clamp01, desc.quantize = k, int(clamp), int(quantize)
        _lib.call("wm_mix_fwd", C.byref(desc), alpha.data_ptr(), out.data_ptr(), b, ys[0].numel() // b, _stream())
        ctx.save_for_backward(alpha)
        ctx.meta = (k, tuple(shape))
        return out

    @staticmethod
    def backward(ctx, gy):
        (alpha,) = ctx.saved_tensors
        k, shape = ctx.meta
        gy = _flat(gy, "attack mix backward")
        grads = [torch.empty(shape, device=gy.device, dtype=torch.float32) if ctx.needs_input_grad[3 + i] else None for i in range(k)]
        desc = _lib.MixDesc()
        for i, g in enumerate(grads):
            desc.t[i] = g.data_ptr() if g is not None else None
        desc.K = k
        _lib.call("wm_mix_bwd", gy.data_ptr(), alpha.data_ptr(), C.byref(desc), shape[0], gy.numel() // shape[0], _stream())
        return (None, None, None) + tuple(grads)


def attack_mix(ys, alpha, clamp: bool = True, quantize: bool = True):
    """Quantization(clamp_with_grad(sum_k alpha[:, k] * ys[k])) — the hybrid attack of models/IRNcrop_model.py:357-373."""
    return _MixFn.apply(alpha, clamp, quantize, *ys)


class _SpliceFn(torch.autograd.Function):
    """out = a * (1 - mask) + b * mask, mask [B,1,H,W] broadcast over channels
    (models/IRNcrop_model.py:348)."""

    @staticmethod
    def forward(ctx, a, b, mask):
        a = _flat(a, "splice"); b = _flat(b, "splice"); mask = _flat(mask, "splice")
        bsz, c, h, w = a.shape
        if b.shape != a.shape or mask.shape != (bsz, 1, h, w):
            raise ValueError(f"splice: expected b {tuple(a.shape)} and mask {(bsz, 1, h, w)}, got {tuple(b.shape)}, {tuple(mask.shape)}")
        out = torch.empty_like(a)
        _lib.call("wm_splice_fwd", a.data_ptr(), b.data_ptr(), mask.data_ptr(), out.data_ptr(), bsz, c, h * w, _stream())
        ctx.save_for_backward(mask)
        ctx.shape = (bsz, c, h, w)
        return out

    @staticmethod
    def backward(ctx, gy):
        (mask,) = ctx.saved_tensors
        bsz, c, h, w = ctx.shape
        gy = _flat(gy, "splice backward")
        ga = torch.empty_like(gy) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(gy) if ctx.needs_input_grad[1] else None
        if ga is not None or gb is not None:
            _lib.call("wm_splice_bwd", gy.data_ptr(), mask.data_ptr(), _ptr(ga), _ptr(gb), bsz, c, h * w, _stream())
        return ga, gb, None


def splice(a, b, mask):
    return _SpliceFn.apply(a, b, mask)


def from_uint8(frames: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """8-bit frames (any shape, CUDA uint8) -> float32 in [0,1] = frames.float() / 255, one kernel.
    Lets a loader upload bytes (4x less PCIe traffic than fp32) and convert on the device."""
    _check_cuda(frames, "from_uint8")
    if frames.dtype != torch.uint8:
        raise TypeError(f"from_uint8: expected a uint8 tensor, got {frames.dtype}")
    frames = frames.contiguous()
    if out is None:
        out = torch.empty(frames.shape, device=frames.device, dtype=torch.float32)
    elif out.shape != frames.shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("from_uint8: `out` must be a contiguous float32 tensor of the same shape")
    _lib.call("wm_u8_to_unit_float", frames.data_ptr(), out.data_ptr(), frames.numel(), _stream())
    return out


def to_uint8(frames: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """float32 frames in [0,1] (any shape, CUDA) -> uint8 = round(clamp(frames, 0, 1) * 255), one kernel: the
    byte k of the Quantization layer's value k/255, so results can leave the device as bytes."""
    _check_cuda(frames, "to_uint8")
    frames = _f32(frames.detach()).contiguous()
    if out is None:
        out = torch.empty(frames.shape, device=frames.device, dtype=torch.uint8)
    elif out.shape != frames.shape or out.dtype != torch.uint8 or not out.is_contiguous():
        raise ValueError("to_uint8: `out` must be a contiguous uint8 tensor of the same shape")
    _lib.call("wm_unit_float_to_u8", frames.data_ptr(), out.data_ptr(), frames.numel(), _stream())
    return out


# ---- real codec round trip (JpegTest, noise_layers/jpeg.py:10-45) -----------------------------------

_CODEC_MODES = {"signed": 0, "unit": 1, "uint8": 2}


def jpeg_codec(x: torch.Tensor, quality: int, subsampling: int = 2, value_range: str = "signed",
               return_coefficients: bool = False):
    """What ``Image.open(Image.fromarray(frame).save(quality=quality, subsampling=subsampling))``
    returns, for every frame of x [B,3,H,W], computed on the device bit-for-bit like libjpeg
    (entropy coding is lossless and is skipped).  Not differentiable, as upstream.

    value_range: "signed" = float in [-1,1] with JpegTest's conversions (noise_layers/jpeg.py:28,38-43),
    "unit" = float in [0,1] (8-bit rounding, /255), "uint8" = bytes in, bytes out.
    return_coefficients: also return (Y, Cb, Cr) int16 planes of quantised DCT coefficients laid out
    like the (block-padded) component planes — the integers the entropy coder would see."""
    _check_cuda(x, "jpeg_codec")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"jpeg_codec: expected [B,3,H,W], got {tuple(x.shape)}")
    if value_range not in _CODEC_MODES:
        raise ValueError(f"jpeg_codec: value_range must be one of {sorted(_CODEC_MODES)}")
    mode = _CODEC_MODES[value_range]
    quality, subsampling = int(quality), int(subsampling)
    if not 1 <= quality <= 100:
        raise ValueError(f"jpeg_codec: quality {quality} outside 1..100")
    if subsampling not in (0, 1, 2):
        raise ValueError("jpeg_codec: subsampling must be 0 (4:4:4), 1 (4:2:2) or 2 (4:2:0)")
    x = x.detach()
    if mode == 2:
        if x.dtype != torch.uint8:
            raise TypeError(f"jpeg_codec: value_range='uint8' needs a uint8 tensor, got {x.dtype}")
    else:
        x = _f32(x)
    if x.stride(3) != 1 or min(x.stride()) < 0:
        x = x.contiguous()
    b, _, h, w = x.shape
    y = torch.empty((b, 3, h, w), device=x.device, dtype=x.dtype)
    if b == 0 or h == 0 or w == 0:
        return (y, None) if return_coefficients else y
    nbytes = _lib.load().wm_jpegcodec_scratch_bytes(b, h, w, subsampling)
    scratch = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    coef = torch.zeros(nbytes, device=x.device, dtype=torch.int16) if return_coefficients else None
    _lib.call("wm_jpegcodec", x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), y.data_ptr(), b, h, w,
              quality, subsampling, mode, scratch.data_ptr(), coef.data_ptr() if coef is not None else None,
              _stream())
    if not return_coefficients:
        return y
    hs, vs = ((1, 1), (2, 1), (2, 2))[subsampling]
    mcu_rows = -(-h // (8 * vs))
    hy, wt = mcu_rows * 8 * vs, -(-w // 256) * 256
    hc, wc = mcu_rows * 8, wt // hs
    ny = b * hy * wt
    yq = coef[:ny].view(b, hy, wt)
    cq = coef[ny:].view(2, b, hc, wc)
    wy, wcb = -(-w // (8 * hs)) * 8 * hs, -(-w // (8 * hs)) * 8
    return y, (yq[:, :, :wy], cq[0][:, :, :wcb], cq[1][:, :, :wcb])


# ---- fused store epilogue: the attack kernel itself writes Quantization(x + (clamp(v) - x)); the descriptor
# travels as an explicit argument of the forward entry point (wm_store_epilogue), there is no hidden state

def _ep_arg(ep):
    """ep = (x_dense, clamp, quantize) or None -> the wm_store_epilogue argument of a forward entry point."""
    if ep is None:
        return None
    return C.byref(_lib.StoreEpilogue(ep[0].data_ptr(), int(ep[1]), int(ep[2])))


def _out_ok(out, shape):
    return out is not None and tuple(out.shape) == tuple(shape) and out.is_contiguous() and out.dtype == torch.float32 \
        and out.data_ptr() % 32 == 0


def diffjpeg_into(x, factor, rounding, out, ep=None) -> bool:
    """No-grad DiffJPEG forward written into `out` (optionally through the store epilogue)."""
    x, sb, sc, sh = _image(x, "DiffJPEG")
    b, c, h, w = x.shape
    if c != 3 or h % 16 or w % 16 or not _out_ok(out, x.shape):
        return False
    fs, fps = _factor_args(factor, b, x.device)
    _lib.call("wm_diffjpeg_fwd", x.data_ptr(), DT_F32, sb, sc, sh, out.data_ptr(), b, h, w, fs, _ptr(fps), rounding, _ep_arg(ep), _stream())
    return True


def jpeg8_into(x, params, out, ep=None) -> bool:
    x, sb, sc, sh = _image(x, "jpeg8")
    b, c, h, w = x.shape
    if c != 3 or w % 8 or params.subsample != 0 or not _out_ok(out, x.shape):
        return False
    _lib.call("wm_jpeg8_fwd", x.data_ptr(), DT_F32, sb, sc, sh, out.data_ptr(), b, h, w, C.byref(params), _ep_arg(ep), _stream())
    return True


def gaussian_blur_into(x, taps, out, ep=None) -> bool:
    x, sp, sh = _planes(x, "gaussian blur")
    b, c, h, w = x.shape
    if len(taps) not in (3, 5, 7) or w % 4 or sp % 4 or sh % 4 or x.data_ptr() % 16 or not _out_ok(out, x.shape):
        return False
    _lib.call("wm_gaussblur", x.data_ptr(), sp, sh, out.data_ptr(), b * c, h, w, _taps_array(taps), len(taps), 0, 0, _ep_arg(ep), _stream())
    return True


def median_blur_into(x, k, out, ep=None) -> bool:
    x, sp, sh = _planes(x, "median blur")
    b, c, h, w = x.shape
    if k not in (3, 5) or sp % 4 or sh % 4 or x.data_ptr() % 16 or (k == 3 and w % 4) or not _out_ok(out, x.shape):
        return False
    _lib.call("wm_median_fwd", x.data_ptr(), sp, sh, out.data_ptr(), None, 0, b * c, h, w, k, _ep_arg(ep), _stream())
    return True


def gaussian_noise_into(x, mean, std, clamp, out, ep=None) -> bool:
    x = _flat(x, "gaussian noise")
    if not _out_ok(out, x.shape):
        return False
    seed, offset = next_philox_stream(x.numel(), x.device)
    _lib.call("wm_gaussnoise_fwd", x.data_ptr(), DT_F32, out.data_ptr(), x.numel(), float(mean), float(std), int(clamp),
              seed, _off(offset), None, _ep_arg(ep), _stream())
    return True


def resize_roundtrip_into(x, mid_hw, mode, out, ep=None) -> bool:
    h, w = x.shape[2:]
    mid_hw = (int(mid_hw[0]), int(mid_hw[1]))
    n = x.shape[0] * x.shape[1]
    x, sp, sh = _planes(x, "resize")
    if not (_lib.load().wm_resize_is_fused(h, w, mid_hw[0], mid_hw[1], n) and w % 4 == 0 and sp % 4 == 0 and sh % 4 == 0
            and x.data_ptr() % 16 == 0 and _out_ok(out, x.shape)):       # the store epilogue needs the TMA-fed instantiation
        return False
    tables = _resize_tables(x.device, h, w, mid_hw, _MODES[mode])
    if tables is None:
        return False
    _lib.call("wm_resize_fwd", x.data_ptr(), sp, sh, out.data_ptr(), n, h, w, mid_hw[0], mid_hw[1], _MODES[mode],
              None, tables.data_ptr(), _ep_arg(ep), _stream())
    return True


class _BankFusedFn(torch.autograd.Function):
    """K attacked variants written by the attack kernels themselves, through the store epilogue, into
    the slices of one [K*B, C, H, W] tensor; layers without a fused path run normally and are
    finished by the stand-alone epilogue kernel.  Backward: straight-through, gx = sum_k gy_k."""

    @staticmethod
    def forward(ctx, x, clamp, quantize, layers, shared_read=True):
        x = _flat(x.detach(), "attack bank")
        if x.data_ptr() % 32:
            x = x.clone()
        k, n = len(layers), x.numel()
        out = torch.empty((k * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
        ep = (x, clamp, quantize)
        names = []
        # 3x3-neighbourhood members (GaussianBlur(3), MiddleBlur(3), Gaussian, Identity) share ONE read of x:
        # they are collected here and launched together (wm_bank3_fwd) after the loop; everything with a host-side
        # random decision or a Philox reservation still happens at the layer's own position, in layer order
        b_, c_, h_, w_ = x.shape
        shared_ok = shared_read and x.dim() == 4 and bool(_lib.load().wm_bank3_ok(b_ * c_, h_, w_))
        kinds = [getattr(layer, "bank3_kind", None) for layer in layers] if shared_ok else [None] * k
        if len({kd for kd in kinds if kd}) < 2:
            kinds = [None] * k                                       # a lone member keeps its own kernel
        desc, taken, keep = _lib.Bank3Desc(), set(), []
        for i, layer in enumerate(layers):
            sl = out[i * x.shape[0]:(i + 1) * x.shape[0]]
            kind = kinds[i]
            if kind and kind not in taken:
                taken.add(kind)
                spec = layer.bank3_member()                          # (also renames the layer like its forward does)
                if kind == "blur":
                    desc.y_blur = sl.data_ptr()
                    for j in range(3):
                        desc.blur_taps[j] = float(spec[j])
                elif kind == "median":
                    desc.y_median = sl.data_ptr()
                elif kind == "noise":
                    seed, offset = next_philox_stream(n, x.device)
                    keep.append(offset)                              # a reserved device slot must outlive the launch
                    desc.y_noise = sl.data_ptr()
                    desc.noise_mean, desc.noise_std, desc.noise_clamp = float(spec[0]), float(spec[1]), int(spec[2])
                    desc.seed, desc.offset = seed, _off(offset)
                else:
                    desc.y_identity = sl.data_ptr()
                names.append(getattr(layer, "name", type(layer).__name__))
                continue
            fused = False
            into = getattr(layer, "forward_into", None)
            if into is not None:
                fused = bool(into(x, sl, ep))
            if not fused:
                y = layer(x)
                y = y[0] if isinstance(y, tuple) else y
                _lib.call("wm_attack_epilogue_fwd", x.data_ptr(), _flat(y, "attack bank").data_ptr(), sl.data_ptr(), n,
                          int(clamp), int(quantize), _stream())
            names.append(getattr(layer, "name", type(layer).__name__))
        if taken:
            desc.clamp01, desc.quantize = int(clamp), int(quantize)
            _lib.call("wm_bank3_fwd", x.data_ptr(), h_ * w_, w_, b_ * c_, h_, w_, C.byref(desc), _stream())
        ctx.meta = (k, tuple(x.shape))
        ctx.names = names
        return out

    @staticmethod
    def backward(ctx, gy):
        k, shape = ctx.meta
        gy = _flat(gy, "attack bank backward")
        gx = torch.empty(shape, device=gy.device, dtype=torch.float32)
        _lib.call("wm_slice_sum", gy.data_ptr(), gx.data_ptr(), gx.numel(), k, _stream())
        return gx, None, None, None, None


def attack_bank_fused(x, layers, clamp: bool = True, quantize: bool = True, shared_read: bool = True):
    return _BankFusedFn.apply(x, clamp, quantize, list(layers), shared_read)

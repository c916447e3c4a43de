"""ctypes binding of libwmattack.so (the C ABI declared in include/wm_attack.h).

No CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libwmattack.so")

c_f32p = C.c_void_p   # device pointers travel as integers (tensor.data_ptr())
c_u8p = C.c_void_p
i64 = C.c_int64
u64 = C.c_uint64
i32 = C.c_int
f32 = C.c_float
vp = C.c_void_p


class Jpeg8Params(C.Structure):
    """Mirror of wm_jpeg8_params (include/wm_attack.h)."""
    _fields_ = [("fwd_color", f32 * 9), ("inv_color", f32 * 9), ("table", (f32 * 64) * 3),
                ("variant", i32), ("subsample", i32)]


class StoreEpilogue(C.Structure):
    """Mirror of wm_store_epilogue (include/wm_attack.h): optional argument of the forward entry points."""
    _fields_ = [("x", C.c_void_p), ("clamp01", i32), ("quantize", i32)]


epp = C.POINTER(StoreEpilogue)


class Bank3Desc(C.Structure):
    """Mirror of wm_bank3_desc (include/wm_attack.h): members of the shared-read bank kernel."""
    _fields_ = [("y_blur", C.c_void_p), ("blur_taps", f32 * 3), ("y_median", C.c_void_p), ("y_noise", C.c_void_p),
                ("noise_mean", f32), ("noise_std", f32), ("noise_clamp", i32), ("seed", u64), ("offset", u64),
                ("y_identity", C.c_void_p), ("clamp01", i32), ("quantize", i32)]

class MixDesc(C.Structure):
    """Mirror of wm_mix_desc (include/wm_attack.h): the K members of a convex mix (or their gradients)."""
    _fields_ = [("t", C.c_void_p * 8), ("K", i32), ("clamp01", i32), ("quantize", i32)]


# name -> argtypes; every function returns int.  Kept in one table so that the CPU test-suite
# can check that the built library exports exactly the header's entry points.
SIGNATURES = {
    "wm_diffjpeg_fwd": [vp, i32, i64, i64, i64, c_f32p, i32, i32, i32, f32, c_f32p, i32, epp, vp],
    "wm_diffjpeg_bwd": [vp, i32, i64, i64, i64, c_f32p, i64, i64, i64, vp, i32, i32, i32, i32, f32, c_f32p, i32, vp],
    "wm_diffjpeg_fwd_save": [vp, i32, i64, i64, i64, c_f32p, c_f32p, c_f32p, vp, i32, i32, i32, f32, c_f32p, i32, vp],
    "wm_diffjpeg_bwd_saved": [c_f32p, i64, i64, i64, c_f32p, c_f32p, vp, vp, i32, i32, i32, i32, vp],
    "wm_diffjpeg_compress": [c_f32p, i64, i64, i64, c_f32p, c_f32p, c_f32p, i32, i32, i32, f32, c_f32p, i32, vp],
    "wm_diffjpeg_decompress": [c_f32p, c_f32p, c_f32p, c_f32p, i32, i32, i32, f32, c_f32p, vp],
    "wm_jpeg8_fwd": [vp, i32, i64, i64, i64, c_f32p, i32, i32, i32, C.POINTER(Jpeg8Params), epp, vp],
    "wm_jpeg8_bwd": [c_f32p, i64, i64, i64, c_f32p, i64, i64, i64, vp, i32, i32, i32, i32, C.POINTER(Jpeg8Params), vp],
    "wm_jpeg8_fwd_save": [vp, i32, i64, i64, i64, c_f32p, c_f32p, i32, i32, i32, C.POINTER(Jpeg8Params), vp],
    "wm_jpeg8_bwd_saved": [c_f32p, i64, i64, i64, c_f32p, vp, i32, i32, i32, i32, C.POINTER(Jpeg8Params), vp],
    "wm_jpeg8_quantised": [c_f32p, i64, i64, i64, c_f32p, i32, i32, i32, C.POINTER(Jpeg8Params), vp],
    "wm_gaussblur": [c_f32p, i64, i64, c_f32p, i32, i32, i32, C.POINTER(f32), i32, i32, i32, epp, vp],
    "wm_median_fwd": [c_f32p, i64, i64, c_f32p, c_u8p, i64, i32, i32, i32, i32, epp, vp],
    "wm_median_bwd": [c_f32p, c_u8p, i64, c_f32p, i32, i32, i32, i32, vp],
    "wm_gaussblur_typed": [vp, i32, i64, i64, vp, i32, i32, i32, i32, C.POINTER(f32), i32, vp],
    "wm_median_fwd_typed": [vp, i32, i64, i64, c_f32p, c_u8p, i64, i32, i32, i32, i32, vp],
    "wm_median_bwd_typed": [c_f32p, c_u8p, i64, vp, i32, i32, i32, i32, i32, vp],
    "wm_gaussnoise_fwd": [vp, i32, c_f32p, i64, f32, f32, i32, u64, u64, c_f32p, epp, vp],
    "wm_gaussnoise_bwd": [c_f32p, c_f32p, c_f32p, i64, f32, f32, i32, u64, u64, c_f32p, vp],
    "wm_gaussnoise_fwd_mask": [vp, i32, c_f32p, vp, i64, f32, f32, u64, u64, c_f32p, vp],
    "wm_gaussnoise_bwd_mask": [c_f32p, vp, vp, i32, i64, vp],
    "wm_rng_reserve": [vp, vp, u64, vp],
    "wm_saltpepper_fwd": [c_f32p, c_f32p, i64, f32, u64, u64, c_f32p, vp],
    "wm_saltpepper_bwd": [c_f32p, c_f32p, i64, f32, u64, u64, c_f32p, vp],
    "wm_dropout_elem_fwd": [c_f32p, c_f32p, c_f32p, i64, f32, u64, u64, c_f32p, vp],
    "wm_dropout_elem_bwd": [c_f32p, c_f32p, c_f32p, i64, f32, u64, u64, c_f32p, vp],
    "wm_dropout_mask_fwd": [c_f32p, c_f32p, c_f32p, c_f32p, i64, i64, vp],
    "wm_dropout_mask_bwd": [c_f32p, c_f32p, c_f32p, c_f32p, i64, i64, vp],
    "wm_bernoulli_mask": [c_f32p, i64, f32, u64, u64, vp],
    "wm_quantize8_fwd": [c_f32p, c_f32p, i64, i32, vp],
    "wm_cropout_fwd": [c_f32p, c_f32p, c_f32p, i64, i32, i32, i32, i32, i32, i32, vp],
    "wm_interp_fwd": [c_f32p, i64, i64, i32, i32, i32, i32, i32, i32, c_f32p, i32, i32, i32, i32, i32, vp, vp],
    "wm_interp_bwd": [c_f32p, c_f32p, vp, i32, i32, i32, c_f32p, i32, i32, i32, i32, i32, i32, i32, c_f32p, vp],
    "wm_u8_to_unit_float": [c_u8p, c_f32p, i64, vp],
    "wm_unit_float_to_u8": [c_f32p, c_u8p, i64, vp],
    "wm_bank3_fwd": [c_f32p, i64, i64, i32, i32, i32, C.POINTER(Bank3Desc), vp],
    "wm_mix_fwd": [C.POINTER(MixDesc), c_f32p, c_f32p, i64, i64, vp],
    "wm_mix_bwd": [c_f32p, c_f32p, C.POINTER(MixDesc), i64, i64, vp],
    "wm_attack_epilogue_fwd": [c_f32p, c_f32p, c_f32p, i64, i32, i32, vp],
    "wm_slice_sum": [c_f32p, c_f32p, i64, i32, vp],
    "wm_splice_fwd": [c_f32p, c_f32p, c_f32p, c_f32p, i64, i32, i64, vp],
    "wm_splice_bwd": [c_f32p, c_f32p, c_f32p, c_f32p, i64, i32, i64, vp],
    "wm_resize_tables": [c_f32p, i32, i32, i32, i32, i32, vp],
    "wm_resize_fwd": [c_f32p, i64, i64, c_f32p, i32, i32, i32, i32, i32, i32, vp, c_f32p, epp, vp],
    "wm_resize_bwd": [c_f32p, vp, c_f32p, i32, i32, i32, i32, i32, i32, c_f32p, vp],
    "wm_resize_fwd_typed": [vp, i32, i64, i64, c_f32p, i32, i32, i32, i32, i32, i32, vp, c_f32p, vp],
    "wm_resize_bwd_typed": [c_f32p, vp, vp, i32, i32, i32, i32, i32, i32, i32, c_f32p, vp],
    "wm_cropresize_fwd": [c_f32p, i64, i64, i32, i32, i32, i32, i32, i32, c_f32p, i32, i32, i32, i32, vp, vp],
    "wm_cropresize_bwd": [c_f32p, c_f32p, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp],
    "wm_jpegcodec": [vp, i64, i64, i64, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp],
}

# kernels launched by one successful call (wm_interp_bwd runs its two gather passes)
KERNELS_PER_CALL = {name: 1 for name in SIGNATURES}
KERNELS_PER_CALL["wm_resize_tables"] = 6
KERNELS_PER_CALL["wm_jpegcodec"] = 2
KERNELS_PER_CALL["wm_cropresize_fwd"] = 2
KERNELS_PER_CALL["wm_cropresize_bwd"] = 2
# plain (non-status) helpers: name -> (restype, argtypes)
HELPERS = {"wm_interp_is_tiled": (C.c_int, [i32, i32, i32, i32, i32]),
           "wm_bank3_ok": (C.c_int, [i32, i32, i32]),
           "wm_resize_is_fused": (C.c_int, [i32, i32, i32, i32, i32]),
           "wm_resize_table_floats": (C.c_int64, [i32, i32, i32, i32]),
           "wm_cropresize_ok": (C.c_int, [i32, i32, i32, i32, i32, i32]),
           "wm_cropresize_table_words": (C.c_int64, [i32, i32, i32, i32, i32]),
           "wm_jpegcodec_scratch_bytes": (C.c_int64, [i32, i32, i32, i32])}

_lock = threading.Lock()
_lib = None
launch_count = 0          # kernels launched through this binding (bench.py reports it)


class WMAttackError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library once.  Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise WMAttackError(
                f"{LIB_PATH} is missing: build it with `python -m wmattack.build` "
                "(or __graft_entry__.build()); this package has no CPU / PyTorch fallback")
        lib = C.CDLL(LIB_PATH)
        lib.wm_version.restype = C.c_int
        lib.wm_version.argtypes = []
        lib.wm_last_error.restype = C.c_char_p
        lib.wm_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the library is stale
            fn.argtypes = argtypes
            fn.restype = C.c_int
        for name, (restype, argtypes) in HELPERS.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


_fns: dict = {}            # name -> bound ctypes function (skips CDLL.__getattr__ on the hot path)


def call(name: str, *args) -> None:
    global launch_count
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = getattr(load(), name)
    rc = fn(*args)
    launch_count += KERNELS_PER_CALL[name]
    if rc != 0:
        lib = load()
        msg = lib.wm_last_error().decode(errors="replace")
        kind = "invalid argument" if rc < 0 else "CUDA error"
        raise WMAttackError(f"{name} failed ({kind} {rc}): {msg}")

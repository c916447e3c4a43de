"""wmattack — B200-native "tailored attacking layer" (host side).

Mirrors the nn.Module surface of yingqichao/video-watermarking-forgery-detection's
``noise_layers`` package and ``utils/JPEG.py::DiffJPEG`` on top of libwmattack.so
(hand-written sm_100a CUDA kernels behind the C ABI of include/wm_attack.h).
"""
from . import functional  # noqa: F401
from .modules import (  # noqa: F401
    Combined, Crop, Cropout, DiffJPEG, Dropout, ElementDropout, GF, GN, Gaussian, GaussianBlur,
    Identity, Jpeg, JpegCompression, JpegMask, JpegSS, JpegTest, MaskDropout, MiddleBlur, Quantization,
    Resize, SaltPepper, compress_jpeg, decompress_jpeg, diff_round, get_random_float, get_random_int,
    quality_to_factor, round_only_at_0, AttackBank, AttackEpilogue, AttackMix, Splice,
)

__all__ = [
    "Combined", "Crop", "Cropout", "DiffJPEG", "Dropout", "ElementDropout", "GF", "GN", "Gaussian",
    "GaussianBlur", "Identity", "Jpeg", "JpegCompression", "JpegMask", "JpegSS", "JpegTest", "MaskDropout",
    "MiddleBlur", "Quantization", "Resize", "SaltPepper", "compress_jpeg", "decompress_jpeg", "diff_round",
    "get_random_float", "get_random_int", "quality_to_factor", "round_only_at_0", "functional",
    "AttackBank", "AttackEpilogue", "AttackMix", "Splice",
]

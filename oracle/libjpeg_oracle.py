"""CPU oracle for the real-codec path (``JpegTest``) — TEST INFRASTRUCTURE ONLY.

The reference's ``JpegTest`` (noise_layers/jpeg.py:10-45) writes each frame with
``PIL.Image.save(format="JPEG", quality=Q, subsampling=s)`` and reads it back.  The arithmetic
therefore lives in a third-party dependency, Pillow + libjpeg-turbo (the image has Pillow 12.2
linked against libjpeg-turbo, API level 6.2).  Huffman coding is lossless, so the decoded pixels
are a deterministic integer function of the input bytes:

    RGB -> YCbCr (16-bit fixed point) -> h2v2 / h2v1 box downsample with alternating bias
    -> 8x8 forward DCT "islow" (13-bit constants) -> divide by 8*Q[u,v] rounding half away from 0
    -> multiply by Q[u,v] -> inverse DCT "islow" + range limit -> "fancy" triangle upsampling
    -> YCbCr -> RGB (16-bit fixed point).

This file restates that published algorithm (IJG libjpeg 6b as kept by libjpeg-turbo:
jccolor.c, jcsample.c, jfdctint.c, jcdctmgr.c, jcparam.c, jidctint.c, jdsample.c, jdcolor.c,
jdmainct.c) with numpy integer arrays.  Parity pinning: ``tests/test_oracle_golden.py`` checks it
bit-for-bit against Pillow itself (imported in the test; Pillow is part of the image on the GPU
box too) and against ``tests/golden/libjpeg_golden.npz`` produced by Pillow in the build
container (``tests/golden/make_libjpeg_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

# Annex K tables, natural (row = vertical frequency) order — jcparam.c std_*_quant_tbl
STD_LUMA = np.array(
    [[16, 11, 10, 16, 24, 40, 51, 61],
     [12, 12, 14, 19, 26, 58, 60, 55],
     [14, 13, 16, 24, 40, 57, 69, 56],
     [14, 17, 22, 29, 51, 87, 80, 62],
     [18, 22, 37, 56, 68, 109, 103, 77],
     [24, 35, 55, 64, 81, 104, 113, 92],
     [49, 64, 78, 87, 103, 121, 120, 101],
     [72, 92, 95, 98, 112, 100, 103, 99]], dtype=np.int64)
STD_CHROMA = np.full((8, 8), 99, dtype=np.int64)
STD_CHROMA[:4, :4] = [[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]]

CONST_BITS, PASS1_BITS = 13, 2
F_0_298631336, F_0_390180644, F_0_541196100 = 2446, 3196, 4433
F_0_765366865, F_0_899976223, F_1_175875602 = 6270, 7373, 9633
F_1_501321110, F_1_847759065, F_1_961570560 = 12299, 15137, 16069
F_2_053119869, F_2_562915447, F_3_072711026 = 16819, 20995, 25172


def _fix(x: float) -> int:
    return int(x * 65536 + 0.5)


def quant_tables(quality: int):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE)."""
    q = min(max(int(quality), 1), 100)
    scale = 5000 // q if q < 50 else 200 - 2 * q
    out = []
    for std in (STD_LUMA, STD_CHROMA):
        t = (std * scale + 50) // 100
        out.append(np.clip(t, 1, 255))
    return out[0], out[1]


def rgb_to_ycc(rgb: np.ndarray):
    """jccolor.c rgb_ycc_convert.  rgb: [H,W,3] uint8 -> three int64 planes."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    half, off = 1 << 15, 128 << 16
    y = (_fix(0.29900) * r + _fix(0.58700) * g + _fix(0.11400) * b + half) >> 16
    cb = (-_fix(0.16874) * r - _fix(0.33126) * g + _fix(0.50000) * b + off + half - 1) >> 16
    cr = (_fix(0.50000) * r - _fix(0.41869) * g - _fix(0.08131) * b + off + half - 1) >> 16
    return y, cb, cr


def _pad_edge(p: np.ndarray, hp: int, wp: int) -> np.ndarray:
    return np.pad(p, ((0, hp - p.shape[0]), (0, wp - p.shape[1])), mode="edge")


def downsample(p: np.ndarray, hs: int, vs: int, blocks_h: int, blocks_w: int) -> np.ndarray:
    """jcprepct.c + jcsample.c: pad the full-resolution rows to the sampling group by replication,
    pad columns to the component's block width, box-average with the alternating bias, then
    replicate the last DOWNSAMPLED row to the block height."""
    H, W = p.shape
    rows = -(-H // vs) * vs
    p = _pad_edge(p, rows, blocks_w * 8 * hs)
    if hs == 2 and vs == 2:
        bias = np.tile(np.array([1, 2], dtype=np.int64), blocks_w * 4)[None, :]
        d = (p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + bias) >> 2
    elif hs == 2 and vs == 1:
        bias = np.tile(np.array([0, 1], dtype=np.int64), blocks_w * 4)[None, :]
        d = (p[:, 0::2] + p[:, 1::2] + bias) >> 1
    else:
        d = p
    return _pad_edge(d, blocks_h * 8, blocks_w * 8)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d, first: bool):
    """jfdctint.c, one pass over the LAST axis of d ([..., 8])."""
    d0, d1, d2, d3, d4, d5, d6, d7 = (d[..., i] for i in range(8))
    t0, t7, t1, t6 = d0 + d7, d0 - d7, d1 + d6, d1 - d6
    t2, t5, t3, t4 = d2 + d5, d2 - d5, d3 + d4, d3 - d4
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    sh = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS
    if first:
        o0, o4 = (t10 + t11) << PASS1_BITS, (t10 - t11) << PASS1_BITS
    else:
        o0, o4 = _descale(t10 + t11, PASS1_BITS), _descale(t10 - t11, PASS1_BITS)
    z1 = (t12 + t13) * F_0_541196100
    o2 = _descale(z1 + t13 * F_0_765366865, sh)
    o6 = _descale(z1 - t12 * F_1_847759065, sh)
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * F_1_175875602
    t4, t5, t6, t7 = t4 * F_0_298631336, t5 * F_2_053119869, t6 * F_3_072711026, t7 * F_1_501321110
    z1, z2 = -z1 * F_0_899976223, -z2 * F_2_562915447
    z3, z4 = -z3 * F_1_961570560 + z5, -z4 * F_0_390180644 + z5
    o7, o5 = _descale(t4 + z1 + z3, sh), _descale(t5 + z2 + z4, sh)
    o3, o1 = _descale(t6 + z2 + z3, sh), _descale(t7 + z1 + z4, sh)
    return np.stack([o0, o1, o2, o3, o4, o5, o6, o7], axis=-1)


def fdct_islow(blk: np.ndarray) -> np.ndarray:
    """jfdctint.c jpeg_fdct_islow on [...,8,8] level-shifted samples (rows first, then columns);
    output is 8x the true DCT."""
    a = _fdct_1d(blk, True)
    return np.swapaxes(_fdct_1d(np.swapaxes(a, -1, -2), False), -1, -2)


def quantize(coef: np.ndarray, qtbl: np.ndarray) -> np.ndarray:
    """jcdctmgr.c quantize(): divisor 8*Q, round half away from zero."""
    div = qtbl << 3
    mag = (np.abs(coef) + (div >> 1)) // div
    return np.where(coef < 0, -mag, mag)


def _idct_1d(c, first: bool):
    c0, c1, c2, c3, c4, c5, c6, c7 = (c[..., i] for i in range(8))
    z1 = (c2 + c6) * F_0_541196100
    t2 = z1 - c6 * F_1_847759065
    t3 = z1 + c2 * F_0_765366865
    t0, t1 = (c0 + c4) << CONST_BITS, (c0 - c4) << CONST_BITS
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = c7, c5, c3, c1
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F_1_175875602
    t0, t1, t2, t3 = t0 * F_0_298631336, t1 * F_2_053119869, t2 * F_3_072711026, t3 * F_1_501321110
    z1, z2 = -z1 * F_0_899976223, -z2 * F_2_562915447
    z3, z4 = -z3 * F_1_961570560 + z5, -z4 * F_0_390180644 + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    sh = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS + 3
    o = [t10 + t3, t11 + t2, t12 + t1, t13 + t0, t13 - t0, t12 - t1, t11 - t2, t10 - t3]
    return np.stack([_descale(v, sh) for v in o], axis=-1)


def range_limit_idct(x: np.ndarray) -> np.ndarray:
    """jdmaster.c prepare_range_limit_table as indexed by jidctint.c
    (``range_limit[x & RANGE_MASK]`` with the +128 centre folded into the table)."""
    t = x & 1023
    return np.where(t < 128, t + 128, np.where(t < 512, 255, np.where(t < 896, 0, t - 896)))


def idct_islow(coef: np.ndarray) -> np.ndarray:
    """jidctint.c jpeg_idct_islow on dequantised [...,8,8] coefficients (columns first, then
    rows); returns samples 0..255.  The zero-AC shortcuts of the C code give identical values."""
    a = np.swapaxes(_idct_1d(np.swapaxes(coef, -1, -2), True), -1, -2)
    return range_limit_idct(_idct_1d(a, False))


def _blocks(p: np.ndarray) -> np.ndarray:
    h, w = p.shape
    return p.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3)


def _unblocks(b: np.ndarray) -> np.ndarray:
    nh, nw = b.shape[:2]
    return b.transpose(0, 2, 1, 3).reshape(nh * 8, nw * 8)


def codec_plane(p: np.ndarray, qtbl: np.ndarray) -> np.ndarray:
    """forward DCT, quantise, dequantise, inverse DCT of a block-aligned plane."""
    q = quantize(fdct_islow(_blocks(p) - 128), qtbl)
    return _unblocks(idct_islow(q * qtbl))


def quantised_plane(p: np.ndarray, qtbl: np.ndarray) -> np.ndarray:
    """the integers the entropy coder would see, laid out like the plane."""
    return _unblocks(quantize(fdct_islow(_blocks(p) - 128), qtbl))


def upsample_h2v2_fancy(d: np.ndarray) -> np.ndarray:
    """jdsample.c h2v2_fancy_upsample with jdmainct.c's context rows (row above the first and
    below the last real row are copies of them).  d: [h,w] real downsampled samples."""
    h, w = d.shape
    up = np.concatenate([d[:1], d[:-1]], 0)
    dn = np.concatenate([d[1:], d[-1:]], 0)
    out = np.empty((2 * h, 2 * w), dtype=np.int64)
    for v, nb in ((0, up), (1, dn)):
        col = 3 * d + nb                                   # thiscolsum per column
        last = np.concatenate([col[:, :1], col[:, :-1]], 1)
        nxt = np.concatenate([col[:, 1:], col[:, -1:]], 1)
        even = (3 * col + last + 8) >> 4
        odd = (3 * col + nxt + 7) >> 4
        even[:, 0] = (4 * col[:, 0] + 8) >> 4
        odd[:, -1] = (4 * col[:, -1] + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out


def upsample_h2v1_fancy(d: np.ndarray) -> np.ndarray:
    """jdsample.c h2v1_fancy_upsample."""
    h, w = d.shape
    last = np.concatenate([d[:, :1], d[:, :-1]], 1)
    nxt = np.concatenate([d[:, 1:], d[:, -1:]], 1)
    out = np.empty((h, 2 * w), dtype=np.int64)
    out[:, 0::2] = (3 * d + last + 1) >> 2
    out[:, 1::2] = (3 * d + nxt + 2) >> 2
    out[:, 0] = d[:, 0]
    out[:, -1] = d[:, -1]
    return out


def ycc_to_rgb(y, cb, cr) -> np.ndarray:
    """jdcolor.c ycc_rgb_convert."""
    half = 1 << 15
    cbx, crx = cb - 128, cr - 128
    r = y + ((_fix(1.40200) * crx + half) >> 16)
    g = y + ((-_fix(0.34414) * cbx + half - _fix(0.71414) * crx) >> 16)
    b = y + ((_fix(1.77200) * cbx + half) >> 16)
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def jpeg_roundtrip_u8(rgb: np.ndarray, quality: int, subsampling: int = 2) -> np.ndarray:
    """What ``Image.open(save(rgb, quality, subsampling))`` returns.  rgb: [H,W,3] uint8;
    subsampling 0 = 4:4:4, 1 = 4:2:2, 2 = 4:2:0 (Pillow's numbering)."""
    H, W, _ = rgb.shape
    hs, vs = {0: (1, 1), 1: (2, 1), 2: (2, 2)}[subsampling]
    ql, qc = quant_tables(quality)
    y, cb, cr = rgb_to_ycc(rgb)
    mcu_w, mcu_h = -(-W // (8 * hs)), -(-H // (8 * vs))
    yp = codec_plane(downsample(y, 1, 1, mcu_h * vs, mcu_w * hs), ql)[:H, :W]
    ch, cw = -(-H // vs), -(-W // hs)
    planes = []
    for c in (cb, cr):
        d = codec_plane(downsample(c, hs, vs, mcu_h, mcu_w), qc)[:ch, :cw]
        if (hs, vs) == (2, 2):
            d = upsample_h2v2_fancy(d)
        elif (hs, vs) == (2, 1):
            d = upsample_h2v1_fancy(d)
        planes.append(d[:H, :W])
    return ycc_to_rgb(yp, planes[0], planes[1])


def jpegtest_forward(x: np.ndarray, quality: int, subsampling: int = 2) -> np.ndarray:
    """noise_layers/jpeg.py:21-45 on a float32 [B,3,H,W] array in [-1,1]: the same fp32 steps
    ((clamp(x)+1)/2*255 truncated to uint8; ToTensor's /255; Normalize's (t-0.5)/0.5)."""
    x = np.asarray(x, dtype=np.float32)
    u = ((np.clip(x, -1, 1) + np.float32(1)) / np.float32(2) * np.float32(255)).astype(np.uint8)
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        dec = jpeg_roundtrip_u8(np.ascontiguousarray(u[i].transpose(1, 2, 0)), quality, subsampling)
        t = dec.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
        out[i] = (t - np.float32(0.5)) / np.float32(0.5)
    return out

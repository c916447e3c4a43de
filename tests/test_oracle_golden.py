"""Pins the CPU oracle (oracle/attack_oracle.py) against fixtures produced by the
UNMODIFIED reference modules (tests/golden/make_golden.py, run in the build container).

CPU-only; no CUDA code is touched here."""
import numpy as np
import pytest
import torch

from oracle import attack_oracle as O
from tests.golden_util import GOLD, T, text

ROUND = {"r0": O.ROUND_ONLY_AT_0, "cubic": O.ROUND_CUBIC, "hard": O.ROUND_HARD}
DT = [torch.float32, torch.float64]


def maxdiff(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max())


def run_with_grad(fn, x, g, dt):
    xx = x.to(dt).clone().requires_grad_(True)
    y = fn(xx)
    y.backward(g.to(dt))
    return y.detach(), xx.grad.detach()


# ------------------------------------------------------------------------------ DiffJPEG
DJ_CASES = [(q, rn, xn) for q in (10, 50, 75, 95) for rn in ROUND for xn in ("x32", "xs32")
            if f"diffjpeg/q{q}/{rn}/{xn}/y" in GOLD]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("q,rn,xn", DJ_CASES)
def test_diffjpeg_forward_backward(q, rn, xn, dt):
    y, gx = run_with_grad(lambda t: O.diffjpeg(t, q, ROUND[rn]), T(xn), T("g32"), dt)
    assert maxdiff(y, T(f"diffjpeg/q{q}/{rn}/{xn}/y")) <= 1e-6
    # the reference's own fp32 gradient noise reaches ~2.5e-5 at q95 (|q| ~ 0.5 branch flips)
    assert maxdiff(gx, T(f"diffjpeg/q{q}/{rn}/{xn}/gx")) <= (5e-5 if q == 95 else 2e-5)


@pytest.mark.parametrize("dt", DT)
def test_diffjpeg_nonsquare_and_saturated(dt):
    y, gx = run_with_grad(lambda t: O.diffjpeg(t, 30), T("x4832"), T("g4832"), dt)
    assert maxdiff(y, T("diffjpeg/q30/r0/x4832/y")) <= 1e-6
    assert maxdiff(gx, T("diffjpeg/q30/r0/x4832/gx")) <= 1e-5
    y, gx = run_with_grad(lambda t: O.diffjpeg(t, 50), T("xsat"), T("g32"), dt)
    assert maxdiff(y, T("diffjpeg/q50/r0/xsat/y")) <= 1e-6
    # clamp mask: compare where the reference's pre-clamp value is not within fp32 noise of
    # the [0,1] bounds (gradient there legitimately flips between 0, 1/2 and 1)
    yref = T("diffjpeg/q50/r0/xsat/y")
    interior = ((yref > 1e-5) & (yref < 1 - 1e-5)).all(dim=1, keepdim=True)
    # a pixel's gradient mixes its whole 16x16 MCU; keep MCUs that are entirely interior
    mcu = torch.nn.functional.avg_pool2d(interior.float(), 16) == 1
    keep = mcu.repeat_interleave(16, 2).repeat_interleave(16, 3).expand_as(gx)
    if keep.any():
        assert maxdiff(gx[keep], T("diffjpeg/q50/r0/xsat/gx")[keep]) <= 1e-5
    assert text("diffjpeg/name_q30") == "DiffJPEG30"


@pytest.mark.parametrize("rn", list(ROUND))
def test_diffjpeg_compress_coefficients(rn):
    x = T("xs32")
    fac = O.quality_to_factor(50)
    for dt in DT:
        y, cb, cr = O.diffjpeg_compress(x.to(dt), fac, ROUND[rn])
        for got, key in ((y, "coef_y"), (cb, "coef_cb"), (cr, "coef_cr")):
            ref = T(f"diffjpeg/q50/{rn}/xs32/{key}")
            assert got.shape == ref.shape
            if rn == "hard":
                # integer-exact except where the fp64 pre-round value sits on a tie
                y64, cb64, cr64 = O.diffjpeg_compress(x.double(), fac, O.ROUND_ONLY_AT_0)
                mism = (got.double() != ref.double())
                assert int(mism.sum()) <= 2
            else:
                assert maxdiff(got, ref) <= 2e-4   # coefficients are O(100)
        dec = O.diffjpeg_decompress(T(f"diffjpeg/q50/{rn}/xs32/coef_y").to(dt),
                                    T(f"diffjpeg/q50/{rn}/xs32/coef_cb").to(dt),
                                    T(f"diffjpeg/q50/{rn}/xs32/coef_cr").to(dt), 32, 32, fac)
        assert maxdiff(dec, T(f"diffjpeg/q50/{rn}/xs32/decompressed")) <= 1e-6


def test_quality_to_factor():
    assert O.quality_to_factor(50) == 1.0
    assert O.quality_to_factor(75) == 0.5
    assert O.quality_to_factor(10) == 5.0
    assert abs(O.quality_to_factor(95) - 0.1) < 1e-12


# ------------------------------------------------------------------- Jpeg / SS / Mask
J8 = {"jpeg": O.JPEG8_HARD, "jpegss": O.JPEG8_SS, "jpegmask": O.JPEG8_MASK}


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("cn", list(J8))
@pytest.mark.parametrize("q", (30, 50, 90))
@pytest.mark.parametrize("sub", (0, 2))
@pytest.mark.parametrize("xn,gn", (("x20", "g20"), ("xs32", "g32")))
def test_jpeg8_family(cn, q, sub, xn, gn, dt):
    ref = T(f"{cn}/q{q}/s{sub}/{xn}/y")
    if cn == "jpeg":
        y = O.jpeg8(T(xn).to(dt), q, J8[cn], sub)
        # hard rounding: a tie flip moves a whole dequantised step; allow a handful of blocks
        bad = ((y.double() - ref.double()).abs() > 2e-6)
        assert int(bad.sum()) <= 64 * 3
    else:
        y, gx = run_with_grad(lambda t: O.jpeg8(t, q, J8[cn], sub), T(xn), T(gn), dt)
        assert maxdiff(y, ref) <= 2e-6
        # reference fp32 gradient noise (gx up to ~3, q near the 0.5 branch) is ~2e-5 at Q90
        assert maxdiff(gx, T(f"{cn}/q{q}/s{sub}/{xn}/gx")) <= 5e-5


def test_jpeg8_quantised_integers():
    ref = T("jpeg/q50/s0/xs32/quantised")
    for dt in DT:
        qv, _ = O.jpeg8_quantised(T("xs32").to(dt), 50, O.JPEG8_HARD, 0)
        assert torch.equal(qv, torch.round(qv))
        assert int((qv.double() != ref.double()).sum()) <= 2
    assert text("jpeg/name_q50") == "Jpeg50"


# -------------------------------------------------------------------- JpegCompression
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("xn", ("x20", "x32", "x2028"))
def test_jpeg_compression(xn, dt):
    y = O.jpeg_compression(T(xn).to(dt))
    assert maxdiff(y, T(f"jpegcompression/{xn}/y")) <= 2e-6


# ------------------------------------------------------------------------------ filters
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("k", (3, 5, 7))
def test_gaussian_blur(k, dt):
    y, gx = run_with_grad(lambda t: O.gaussian_blur(t, k), T("x2028"), T("gaussianblur/g2028"), dt)
    assert maxdiff(y, T(f"gaussianblur/k{k}/x2028/y")) <= 1e-6
    assert maxdiff(gx, T(f"gaussianblur/k{k}/x2028/gx")) <= 1e-6


@pytest.mark.parametrize("k", (3, 5))
@pytest.mark.parametrize("xn", ("x2028", "xs32"))
def test_median_blur_unpinned(k, xn):
    """kornia semantics (reference dependency not available: parity unpinned)."""
    y, idx = O.median_blur(T(xn), k, return_index=True)
    assert torch.equal(y, T(f"middleblur/k{k}/{xn}/y"))
    # index picks an element equal to the median, and it is the first such element
    win = O.median_windows(T(xn), k)
    picked = win.gather(2, idx.long().unsqueeze(2)).squeeze(2)
    assert torch.equal(picked, y)
    # backward through the index == autograd of "gather at idx"
    g = torch.rand(y.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    xx = T(xn).double().requires_grad_(True)
    O.median_windows(xx, k).gather(2, idx.long().unsqueeze(2)).squeeze(2).backward(g)
    assert maxdiff(O.median_blur_backward(g, idx, k), xx.grad) <= 1e-12


def test_gaussian_filter_reflect_unpinned():
    assert maxdiff(O.gaussian_filter_reflect(T("x2028"), 7, 1.5), T("gf/s1.5k7/x2028/y")) <= 1e-6


# -------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize("dt", DT)
def test_elementwise(dt):
    x, g, cover = T("x32"), T("g32"), T("cover32")
    y, gx = run_with_grad(lambda t: O.gaussian_noise_clamped(t, T("gaussian/x32/noise").to(dt)), x, g, dt)
    assert maxdiff(y, T("gaussian/x32/y")) <= 1e-7
    assert maxdiff(gx, T("gaussian/x32/gx")) == 0
    assert maxdiff(O.gaussian_noise_additive(x.to(dt), T("gn/x32/noise").to(dt)), T("gn/x32/y")) <= 1e-7
    y, gx = run_with_grad(lambda t: O.salt_pepper(t, T("saltpepper/p0.1/x32/rdn").to(dt), 0.1), x, g, dt)
    assert maxdiff(y, T("saltpepper/p0.1/x32/y")) == 0
    assert maxdiff(gx, T("saltpepper/p0.1/x32/gx")) == 0
    y = O.dropout_elementwise(x.to(dt), cover.to(dt), T("cropdropout/p0.5/x32/rdn").to(dt), 0.5)
    assert maxdiff(y, T("cropdropout/p0.5/x32/y")) == 0
    y = O.dropout_mask(x.to(dt), cover.to(dt), T("maskdropout/x32/mask"))
    assert maxdiff(y, T("maskdropout/x32/y")) <= 1e-7
    y, gx = run_with_grad(O.quantization, x, g, dt)
    if dt == torch.float32:
        assert maxdiff(y, T("quantization/x32/y")) == 0
    # autograd of round() is 0; the reference's Quant.backward is the identity (STE)
    assert maxdiff(T("quantization/x32/gx"), g) == 0


# -------------------------------------------------------------------------- resize/crop
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("mode", ("bicubic", "bilinear"))
@pytest.mark.parametrize("r", (0.5, 0.7, 1.3, 1.5))
def test_resize(mode, r, dt):
    y, gx = run_with_grad(lambda t: O.resize(t, r, mode), T("x2028"), T("gaussianblur/g2028"), dt)
    assert maxdiff(y, T(f"resize/{mode}/r{r}/x2028/y")) <= 2e-6
    assert maxdiff(gx, T(f"resize/{mode}/r{r}/x2028/gx")) <= 5e-6


def test_resize_saturated_and_random():
    y, gx = run_with_grad(lambda t: O.resize(t, 0.8), T("xsat"), T("g32"), torch.float64)
    assert maxdiff(y, T("resize/bicubic/r0.8/xsat/y")) <= 5e-6
    ratio = float(GOLD["resize/random/ratio"])
    assert maxdiff(O.resize(T("x32").double(), ratio), T("resize/random/x32/y")) <= 5e-6


@pytest.mark.parametrize("mode", ("bicubic", "bilinear"))
def test_interp_matrix_matches_torch(mode):
    x = torch.rand(1, 2, 13, 17, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    for size in ((7, 9), (13, 17), (19, 30), (26, 8)):
        ref = torch.nn.functional.interpolate(x, size=size, mode=mode)
        assert maxdiff(O.interpolate(x, size, mode), ref) <= 1e-12


def test_crop():
    x, g = T("x32"), T("g32")
    apex = tuple(int(v) for v in GOLD["crop/seed18/x32/apex"])
    assert O.crop_box_from_rng(x.shape, np.random.RandomState(18)) == apex
    y, gx = run_with_grad(lambda t: O.crop_resize(t, apex), x, g, torch.float64)
    assert maxdiff(y, T("crop/seed18/x32/y")) <= 2e-6
    assert maxdiff(gx, T("crop/seed18/x32/gx")) <= 5e-6
    apex19 = tuple(int(v) for v in GOLD["crop/seed19/x2028/apex"])
    assert O.crop_box_from_rng(T("x2028").shape, np.random.RandomState(19), 0.7, 0.9) == apex19
    assert maxdiff(O.crop_resize(T("x2028").double(), apex19), T("crop/seed19/x2028/y")) <= 2e-6
    assert maxdiff(O.crop_resize(x.double(), (3, 20, 5, 31)), T("crop/apex/x32/y")) <= 2e-6


# ------------------------------------------------------------------- real codec (JpegTest)
import os as _os  # noqa: E402

from oracle import libjpeg_oracle as LJ  # noqa: E402

_LJ_GOLD = np.load(_os.path.join(_os.path.dirname(__file__), "golden", "libjpeg_golden.npz"))
_LJ_CASES = sorted(k for k in _LJ_GOLD.files if k.startswith("out/"))


def test_libjpeg_golden_present():
    assert len(_LJ_CASES) == 8 * 3 * 4


@pytest.mark.parametrize("key", _LJ_CASES)
def test_libjpeg_oracle_matches_pillow_golden(key):
    """Bit-exact: the restated integer pipeline against what Pillow/libjpeg-turbo returned."""
    _, name, s, q = key.split("/")
    got = LJ.jpeg_roundtrip_u8(_LJ_GOLD[f"in/{name}"], int(q[1:]), int(s[1:]))
    assert np.array_equal(got, _LJ_GOLD[key])


@pytest.mark.parametrize("hw", [(48, 48), (31, 50), (128, 96)])
def test_libjpeg_oracle_matches_pillow_live(hw):
    """Same check against the Pillow of the machine the tests run on (skipped without Pillow)."""
    Image = pytest.importorskip("PIL.Image")
    import io
    rng = np.random.RandomState(hw[0])
    rgb = rng.randint(0, 256, (*hw, 3)).astype(np.uint8)
    for s in (0, 1, 2):
        for q in (5, 35, 60, 85, 95):
            buf = io.BytesIO()
            Image.fromarray(rgb).save(buf, format="JPEG", quality=q, subsampling=s)
            ref = np.array(Image.open(io.BytesIO(buf.getvalue())), dtype=np.uint8)
            assert np.array_equal(LJ.jpeg_roundtrip_u8(rgb, q, s), ref), (s, q)


def test_libjpeg_quant_tables():
    ql, qc = LJ.quant_tables(50)
    assert np.array_equal(ql, LJ.STD_LUMA) and np.array_equal(qc, LJ.STD_CHROMA)
    ql, qc = LJ.quant_tables(100)
    assert ql.min() == 1 and ql.max() == 1 and qc.max() == 1
    ql, _ = LJ.quant_tables(1)
    assert ql.max() == 255            # force_baseline clamp


def test_libjpeg_islow_dct_pair_is_near_identity():
    rng = np.random.RandomState(0)
    blk = rng.randint(0, 256, (50, 8, 8)).astype(np.int64)
    rec = LJ.idct_islow(np.rint(LJ.fdct_islow(blk - 128) / 8).astype(np.int64))   # /8: back to true scale
    assert np.abs(rec - blk).max() <= 2

"""GPU parity tests: the CUDA path (through the C ABI) against
  (1) the committed golden outputs of the unmodified reference (tests/golden), and
  (2) the CPU oracle (oracle/attack_oracle.py, fp64) on seeded inputs,
plus size-independent properties at BASELINE.json's full sizes.

Tolerances (north star): pixels and gradients <= 1e-5 max-abs in fp32; integer-valued
quantised coefficients bit-exact except on fp64-verified rounding ties; selection / where-type
layers bit-exact.

No test here allows a FRACTION of wrong elements.  Where the arithmetic has a step (a rounding
quantiser, the |q| < 1/2 switch of round_only_at_0, a clamp's pass-through mask) a value within fp32
noise of the step may land on either side in ANY fp32 implementation; tests/parity_util.py builds, from
the fp64 oracle's pre-round / pre-clamp values, the set of such FRAGILE positions and their influence
region, and every element outside that region must meet the bound: every mismatch is counted and explained.
Above the BASELINE quality (q > 50) the quantisation steps shrink and fp32 rounding of the level-shifted
DCT input is amplified into the gradient; there the bound is calibrated per case against the SAME formulas
evaluated by torch in fp32 on the CPU (`fp32_noise`): ours must stay within 2x that, never a hand-set table.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import attack_oracle as O  # noqa: E402
from tests import parity_util as P  # noqa: E402
from tests.golden_util import GOLD, T, text  # noqa: E402

import wmattack  # noqa: E402
from wmattack import functional as WF  # noqa: E402

DEV = "cuda"
ROUND = {"r0": 0, "cubic": 1, "hard": 2}


def md(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def rnd(shape, seed, dtype=torch.float32):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed), dtype=dtype)


def fwd_bwd(fn, x, g):
    xx = x.to(DEV).clone().requires_grad_(True)
    y = fn(xx)
    y.backward(g.to(DEV))
    return y.detach().cpu(), xx.grad.detach().cpu()


def oracle_fwd_bwd(fn, x, g):
    xx = x.double().clone().requires_grad_(True)
    y = fn(xx)
    y.backward(g.double())
    return y.detach(), xx.grad.detach()


def grad_tol(q, fn=None, x=None, g=None, influence=None):
    """Max-abs gradient bound vs the fp64 oracle: the north star's 1e-5 at the BASELINE quality and below.
    Above it the bound scales with the inverse of the smallest quantisation step: q = C / (table * factor), the
    luminance table's minimum is 10 (utils/JPEG.py:98-104), so at quality 50 (factor 1) the ~1e-5..1e-4 absolute
    fp32 rounding of a level-shifted DCT coefficient (|C| <= 1024, ulp 6e-5) is divided by >= 10; at factor f < 1 it
    is divided by 10 f only, and d round'(q)/dq <= 3 passes it on to the gradient: bound = 1e-5 / factor
    (2e-5 at quality 75, 1e-4 at quality 95).  When the oracle callable is given, the bound is cross-checked
    against the SAME formulas evaluated by torch in fp32 on the CPU: that implementation's own distance to fp64
    must fit under it too (the bound describes fp32, not our kernel)."""
    tol = 1e-5 * max(1.0, 1.0 / O.quality_to_factor(q))
    if fn is not None and q > 50:
        _, dg = P.fp32_noise(fn, x, g)
        if influence is not None:
            dg = dg[~influence]
        assert (float(dg.max()) if dg.numel() else 0.0) <= tol, "torch's own fp32 evaluation exceeds the fp32 bound"
    return tol


def coord_tol(h, w, r, base):
    """Bound for a resize GRADIENT against a comparator that evaluates source coordinates differently (fp64 oracle,
    or torch on another device): ATen computes scale * (o + 0.5) - 0.5 in fp32, so tap weights carry an error of
    ~ulp(coordinate) that differs between correct fp32 builds (FMA contraction): base + 2 * ulp(largest coordinate)."""
    return base + 2 * 2.0 ** (np.floor(np.log2(max(h, w) * max(1.0, 1.0 / r))) - 23)


# =============================================================================== DiffJPEG
DJ_CASES = [(q, rn, xn) for q in (10, 50, 75, 95) for rn in ROUND for xn in ("x32", "xs32")
            if f"diffjpeg/q{q}/{rn}/{xn}/y" in GOLD]


@pytest.mark.parametrize("q,rn,xn", DJ_CASES)
def test_diffjpeg_vs_golden(q, rn, xn):
    m = wmattack.DiffJPEG(True, 32, 32, quality=q, rounding=ROUND[rn])
    x, g = T(xn), T("g32")
    y, gx = fwd_bwd(m, x, g)
    yref, gref = T(f"diffjpeg/q{q}/{rn}/{xn}/y"), T(f"diffjpeg/q{q}/{rn}/{xn}/gx")
    infl, n_fragile = P.diffjpeg_influence(x, O.quality_to_factor(q), ROUND[rn])
    assert n_fragile <= 2                                     # the fixtures hold (almost) no rounding ties
    P.assert_explained(y, yref, 1e-5, infl, "y vs reference")
    if rn == "hard":
        assert md(gx, gref) == 0.0                            # torch.round: zero gradient, exactly
        return
    # the golden gradient is the reference's own fp32 result: ITS distance to the fp64 oracle (measured here) adds to ours
    fn = lambda t: O.diffjpeg(t, q, ROUND[rn])
    _, go = oracle_fwd_bwd(fn, x, g)
    P.assert_explained(gx, go, grad_tol(q, fn, x, g, infl), infl, "gx vs fp64 oracle")
    ref_err = float((gref.double() - go).abs()[~infl].max())
    tol = grad_tol(q, fn, x, g, infl) + ref_err
    P.assert_explained(gx, gref, tol, infl, "gx vs reference")


@pytest.mark.parametrize("shape", [(1, 16, 16), (3, 64, 96), (5, 48, 272), (2, 128, 128)])
@pytest.mark.parametrize("q", (10, 50, 75, 95))
@pytest.mark.parametrize("mode", (0, 1, 3))
def test_diffjpeg_vs_oracle(shape, q, mode):
    b, h, w = shape
    x, g = rnd((b, 3, h, w), 100 + h + q), rnd((b, 3, h, w), 200 + w + q)
    m = wmattack.DiffJPEG(True, h, w, quality=q, rounding=mode)
    y, gx = fwd_bwd(m, x, g)
    fn = lambda t: O.diffjpeg(t, q, mode)
    yo, go = oracle_fwd_bwd(fn, x, g)
    infl, n_fragile = P.diffjpeg_influence(x, O.quality_to_factor(q), mode)
    assert n_fragile <= 1e-3 * x.numel() + 2                  # fragile coefficients are rare on random input
    if mode == 3:
        # 9-term Fourier surrogate (utils/JPEG_utils.py:36): smooth, but its derivative reaches ~19 and its second
        # derivative ~2*pi*sum(n) = 280, so fp32 noise in q shows up amplified: bound = 2x the oracle's own fp32 noise
        # (argument 2*pi*n*q in fp32): bound = 1e-3 of the gradient's range, and torch's own fp32 evaluation of the
        # same series must sit in the same band (the bound describes fp32, not our kernel)
        dy, dg = P.fp32_noise(fn, x, g)
        scale = max(1.0, float(go.abs().max()))
        assert md(y, yo) <= 5e-5 and float(dy.max()) <= 5e-5
        assert md(gx, go) <= 1e-3 * scale and float(dg.max()) <= 1e-3 * scale
        assert md(gx, go) <= 4 * float(dg.max()) + 1e-5                       # same order as torch's fp32 result
        return
    P.assert_explained(y, yo, 1e-5, infl, "y")
    P.assert_explained(gx, go, grad_tol(q, fn, x, g, infl), infl, "gx")


def test_diffjpeg_fourier_vs_golden():
    """utils/JPEG_utils.py:36-41 diff_round (9-term Fourier series) through the reference's own DiffJPEG."""
    m = wmattack.DiffJPEG(True, 32, 32, quality=50, rounding=3)
    for xn in ("x32", "xs32"):
        x, g = T(xn), T("g32")
        y, gx = fwd_bwd(m, x, g)
        fn = lambda t: O.diffjpeg(t, 50, 3)
        gref = T(f"diffjpeg/q50/fourier/{xn}/gx")
        assert md(y, T(f"diffjpeg/q50/fourier/{xn}/y")) <= 5e-5
        # both the reference's fp32 gradient and ours sit within 1e-3 of the gradient's range of fp64 (see above)
        assert md(gx, gref) <= 2e-3 * max(1.0, float(gref.abs().max()))


def test_diffjpeg_nonsquare_saturated_and_name():
    m = wmattack.DiffJPEG(True, 48, 32, quality=30)
    y, gx = fwd_bwd(m, T("x4832"), T("g4832"))
    assert md(y, T("diffjpeg/q30/r0/x4832/y")) <= 1e-5
    assert md(gx, T("diffjpeg/q30/r0/x4832/gx")) <= 1e-5
    assert m.name == text("diffjpeg/name_q30") == "DiffJPEG30"
    assert wmattack.DiffJPEG(90).name == "DiffJPEG75"       # upstream positional quirk kept
    m = wmattack.DiffJPEG(True, 32, 32, quality=50)
    y, gx = fwd_bwd(m, T("xsat"), T("g32"))
    yref = T("diffjpeg/q50/r0/xsat/y")
    assert md(y, yref) <= 1e-5
    interior = ((yref > 1e-5) & (yref < 1 - 1e-5)).all(dim=1, keepdim=True)
    mcu = torch.nn.functional.avg_pool2d(interior.float(), 16) == 1
    keep = mcu.repeat_interleave(16, 2).repeat_interleave(16, 3).expand_as(gx)
    if keep.any():
        assert md(gx[keep], T("diffjpeg/q50/r0/xsat/gx")[keep]) <= 1e-5
    # clamped pixels must have exactly zero gradient contribution: all-white / all-black image
    for v in (0.0, 1.0):
        xc = torch.full((1, 3, 32, 32), v)
        y, gx = fwd_bwd(m, xc, T("g32")[:1])
        assert md(y, xc) <= 1e-5


@pytest.mark.parametrize("rn", list(ROUND))
def test_diffjpeg_compress_coefficients(rn):
    """Quantised DCT coefficients: bit-exact for integer work."""
    x = T("xs32")
    m = wmattack.DiffJPEG(True, 32, 32, quality=50, rounding=ROUND[rn])
    cy, ccb, ccr = m.compress(x.to(DEV))
    y64 = O.diffjpeg_compress(x.double(), 1.0, ROUND[rn])
    for got, key, ref64 in ((cy, "coef_y", y64[0]), (ccb, "coef_cb", y64[1]), (ccr, "coef_cr", y64[2])):
        ref = T(f"diffjpeg/q50/{rn}/xs32/{key}")
        assert tuple(got.shape) == tuple(ref.shape)
        if rn == "hard":
            g = got.cpu()
            assert torch.equal(g, torch.round(g))
            mism = g.double() != ref64
            assert int(mism.sum()) <= 1, "integer coefficients must match the fp64 oracle except on ties"
            assert int((g != ref).sum()) <= 2          # the reference's fp32 GEMM has its own tie flips
        else:
            assert md(got, ref) <= 2e-4                # coefficients are O(100): 1e-6 relative
    dec = m.decompress(cy, ccb, ccr)
    assert md(dec, T(f"diffjpeg/q50/{rn}/xs32/decompressed")) <= 1e-5
    # fused forward == decompress(compress(x)) (same device code; only FMA contraction differs)
    assert md(dec, m(x.to(DEV))) <= 5e-7
    if rn == "hard":
        assert torch.equal(dec, m(x.to(DEV)))


def test_diffjpeg_per_sample_quality_and_strided_inputs():
    b, h, w = 4, 32, 48
    x = rnd((b, 3, h, w), 7).to(DEV)
    q = torch.tensor([10.0, 50.0, 75.0, 95.0])
    m = wmattack.DiffJPEG(True, h, w, quality=75)
    y = m(x, quality=q.to(DEV))
    for i in range(b):
        yi = wmattack.DiffJPEG(True, h, w, quality=float(q[i]))(x[i:i + 1])
        assert torch.equal(y[i:i + 1], yi)
    assert torch.equal(m(x, quality=50), wmattack.DiffJPEG(True, h, w, quality=50)(x))
    # frame slices of a [B,3,T,H,W] clip are read in place (models/IRNcrop_model.py:362)
    clip = rnd((2, 3, 5, h, w), 8).to(DEV).requires_grad_(True)
    frame = clip[:, :, 2]
    assert not frame.is_contiguous()
    y1 = m(frame)
    y2 = m(frame.detach().contiguous())
    assert torch.equal(y1, y2)
    # stride-0 cotangent from .sum() and channels-last cotangent
    y1.sum().backward()
    g_ref = wmattack.functional._DiffJPEGFn.apply  # noqa: F841
    xx = frame.detach().contiguous().requires_grad_(True)
    m(xx).backward(torch.ones_like(y2))
    assert torch.equal(clip.grad[:, :, 2], xx.grad)
    assert float(clip.grad[:, :, 1].abs().max()) == 0.0


def test_diffjpeg_errors():
    m = wmattack.DiffJPEG(True, 24, 24, quality=50)
    with pytest.raises(ValueError):
        m(torch.rand(1, 3, 24, 24, device=DEV))                 # not a multiple of 16
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 32, 32))                             # CPU tensor: no fallback
    with pytest.raises(ValueError):
        wmattack.DiffJPEG(True, 32, 32, 50)(torch.rand(1, 1, 32, 32, device=DEV))
    from wmattack import _lib
    with pytest.raises(_lib.WMAttackError):
        _lib.call("wm_diffjpeg_fwd", None, 0, 0, 0, 0, None, 1, 32, 32, 1.0, None, 0, None, None)


def test_diffjpeg_full_size_properties():
    """BASELINE config 2 size (64x3x512x512): size-independent properties."""
    b, h, w = 64, 512, 512
    x = torch.rand(b, 3, h, w, device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    g = torch.rand(b, 3, h, w, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    m = wmattack.DiffJPEG(True, h, w, quality=50)
    xx = x.clone().requires_grad_(True)
    y = m(xx)
    y.backward(g)
    assert torch.isfinite(y).all() and float(y.min()) >= 0 and float(y.max()) <= 1
    # MCU independence: any 16-aligned crop of any image gives the same bits
    for (bi, r0, c0, hh, ww) in ((0, 0, 0, 16, 16), (17, 96, 256, 64, 128), (63, 496, 496, 16, 16)):
        crop = x[bi:bi + 1, :, r0:r0 + hh, c0:c0 + ww].contiguous().requires_grad_(True)
        yc = wmattack.DiffJPEG(True, hh, ww, quality=50)(crop)
        yc.backward(g[bi:bi + 1, :, r0:r0 + hh, c0:c0 + ww].contiguous())
        assert torch.equal(yc, y[bi:bi + 1, :, r0:r0 + hh, c0:c0 + ww])
        assert torch.equal(crop.grad, xx.grad[bi:bi + 1, :, r0:r0 + hh, c0:c0 + ww])
    # spot-check a random subset of images against the fp64 oracle
    for bi in (3, 40):
        xi, gi = x[bi:bi + 1].cpu(), g[bi:bi + 1].cpu()
        yo, go = oracle_fwd_bwd(lambda t: O.diffjpeg(t, 50), xi, gi)
        infl, n_fragile = P.diffjpeg_influence(xi, 1.0, 0)
        assert n_fragile <= 1e-4 * xi.numel()
        P.assert_explained(y[bi:bi + 1], yo, 1e-5, infl, "y")
        P.assert_explained(xx.grad[bi:bi + 1], go, 1e-5, infl, "gx")
    # determinism
    assert torch.equal(m(x), y.detach())


# ================================================================== Jpeg / JpegSS / JpegMask
J8 = {"jpeg": ("Jpeg", O.JPEG8_HARD), "jpegss": ("JpegSS", O.JPEG8_SS), "jpegmask": ("JpegMask", O.JPEG8_MASK)}


@pytest.mark.parametrize("cn", list(J8))
@pytest.mark.parametrize("q", (30, 50, 90))
@pytest.mark.parametrize("sub", (0, 2))
@pytest.mark.parametrize("xn,gn", (("x20", "g20"), ("xs32", "g32")))
def test_jpeg8_vs_golden(cn, q, sub, xn, gn):
    m = getattr(wmattack, J8[cn][0])(q, subsample=sub)
    x, g = T(xn), T(gn)
    y, gx = fwd_bwd(m, x, g)
    ref = T(f"{cn}/q{q}/s{sub}/{xn}/y")
    if cn == "jpegmask":                                       # linear: no step anywhere
        assert md(y, ref) <= 1e-5 and md(gx, T(f"{cn}/q{q}/s{sub}/{xn}/gx")) <= 1e-5
        return
    mode = O.ROUND_HARD if cn == "jpeg" else O.ROUND_ONLY_AT_0
    infl, n_fragile = P.jpeg8_influence(x, q, mode, sub)
    assert n_fragile <= 2
    P.assert_explained(y, ref, 1e-5, infl, "y vs reference")
    if cn == "jpeg":
        assert float(gx.abs().max()) == 0.0
        return
    # golden = the reference's own fp32 gradient: ITS distance to the fp64 oracle (measured here) adds to ours
    fn = lambda t: O.jpeg8(t, q, J8[cn][1], sub)
    _, go = oracle_fwd_bwd(fn, x, g)
    gref = T(f"{cn}/q{q}/s{sub}/{xn}/gx")
    P.assert_explained(gx, go, grad_tol(q, fn, x, g, infl), infl, "gx vs fp64 oracle")
    tol = grad_tol(q, fn, x, g, infl) + float((gref.double() - go).abs()[~infl].max())
    P.assert_explained(gx, gref, tol, infl, "gx vs reference")
    assert tol <= 1e-4


@pytest.mark.parametrize("cn", list(J8))
@pytest.mark.parametrize("shape", [(2, 40, 56), (1, 17, 23), (3, 64, 64), (1, 8, 200)])
@pytest.mark.parametrize("sub", (0, 2))
def test_jpeg8_vs_oracle_any_shape(cn, shape, sub):
    """Non-square and non-multiple-of-8 shapes (the reference itself is square-only)."""
    b, h, w = shape
    x, g = rnd((b, 3, h, w), h * w), rnd((b, 3, h, w), h + w)
    q = 50
    m = getattr(wmattack, J8[cn][0])(q, subsample=sub)
    y, gx = fwd_bwd(m, x, g)
    if cn == "jpegmask":
        yo, go = oracle_fwd_bwd(lambda t: O.jpeg8(t, q, J8[cn][1], sub), x, g)
        assert md(y, yo) <= 1e-5 and md(gx, go) <= 1e-5
        return
    mode = O.ROUND_HARD if cn == "jpeg" else O.ROUND_ONLY_AT_0
    infl, n_fragile = P.jpeg8_influence(x, q, mode, sub)
    assert n_fragile <= 1e-4 * x.numel() + 2
    if cn == "jpeg":
        P.assert_explained(y, O.jpeg8(x.double(), q, J8[cn][1], sub), 1e-5, infl, "y")
        return
    yo, go = oracle_fwd_bwd(lambda t: O.jpeg8(t, q, J8[cn][1], sub), x, g)
    P.assert_explained(y, yo, 1e-5, infl, "y")
    P.assert_explained(gx, go, 1e-5, infl, "gx")


def test_jpeg8_quantised_integers():
    m = wmattack.Jpeg(50)
    qv = m.quantised(T("xs32").to(DEV)).cpu()
    assert torch.equal(qv, torch.round(qv))
    q64, _ = O.jpeg8_quantised(T("xs32").double(), 50, O.JPEG8_HARD, 0)
    assert int((qv.double() != q64).sum()) <= 1
    assert int((qv != T("jpeg/q50/s0/xs32/quantised")).sum()) <= 2
    assert m.name == text("jpeg/name_q50")
    # padded shape
    qv = wmattack.Jpeg(90).quantised(T("x20").to(DEV)).cpu()
    q64, _ = O.jpeg8_quantised(T("x20").double(), 90, O.JPEG8_HARD, 0)
    assert qv.shape == q64.shape == (2, 3, 24, 24)
    assert int((qv.double() != q64).sum()) <= 2


# ======================================================================== JpegCompression
@pytest.mark.parametrize("xn", ("x20", "x32", "x2028"))
def test_jpeg_compression_vs_golden(xn):
    m = wmattack.JpegCompression(DEV)
    y = m(T(xn).to(DEV))
    assert md(y, T(f"jpegcompression/{xn}/y")) <= 1e-5


def test_linear_layers_adjoint_and_linearity():
    """<A x, g> == <x, A^T g> for every linear layer, at a non-trivial size."""
    layers = [wmattack.JpegCompression(DEV), wmattack.JpegMask(50), wmattack.JpegMask(50, subsample=2),
              wmattack.GaussianBlur(3), wmattack.GaussianBlur(7), wmattack.GF(1.5, 7)]
    x, g = rnd((2, 3, 52, 76), 1).to(DEV), rnd((2, 3, 52, 76), 2).to(DEV)
    for layer in layers:
        xx = x.clone().requires_grad_(True)
        y = layer((xx, xx)) if isinstance(layer, wmattack.GF) else layer(xx)
        y.backward(g)
        lhs = float((y.double() * g.double()).sum())
        rhs = float((x.double() * xx.grad.double()).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs)), type(layer).__name__
        y2 = layer((2 * x, x)) if isinstance(layer, wmattack.GF) else layer(2 * x)
        assert md(y2, 2 * y) <= 1e-5


# ============================================================================ blur / median
@pytest.mark.parametrize("k", (3, 5, 7))
def test_gaussian_blur_vs_golden(k):
    m = wmattack.GaussianBlur(kernel_size=k)
    assert m.name == "G_Blur"
    y, gx = fwd_bwd(m, T("x2028"), T("gaussianblur/g2028"))
    assert m.name == "GaussianBlur"
    assert md(y, T(f"gaussianblur/k{k}/x2028/y")) <= 1e-6
    assert md(gx, T(f"gaussianblur/k{k}/x2028/gx")) <= 1e-6


@pytest.mark.parametrize("shape", [(2, 3, 70, 300), (1, 3, 33, 129), (1, 1, 5, 4)])
def test_gaussian_blur_and_gf_vs_oracle(shape):
    x, g = rnd(shape, 3), rnd(shape, 4)
    c = shape[1]
    for k in (3, 9):
        y, gx = fwd_bwd(wmattack.GaussianBlur(k, channels=c), x, g)
        yo, go = oracle_fwd_bwd(lambda t: O.gaussian_blur(t, k), x, g)
        assert md(y, yo) <= 1e-6 and md(gx, go) <= 1e-6
    if min(shape[2:]) > 3:
        k = 7 if min(shape[2:]) > 3 else 3
        if min(shape[2:]) <= k // 2:
            return
        y, gx = fwd_bwd(lambda t: wmattack.GF(1.5, k)((t, t)), x, g)
        yo, go = oracle_fwd_bwd(lambda t: O.gaussian_filter_reflect(t, k, 1.5), x, g)
        assert md(y, yo) <= 1e-6 and md(gx, go) <= 1e-6
    assert md(wmattack.GF(1.5, 7)((T("x2028").to(DEV), None)), T("gf/s1.5k7/x2028/y")) <= 1e-6


@pytest.mark.parametrize("dt", (torch.bfloat16, torch.float16))
def test_diffjpeg_reads_half_inputs_and_stores_half_gradients(dt):
    """Autocast boundary (models/IRNcrop_model.py:340): a float16 / bfloat16 image is read as it is and the input
    gradient leaves in that type — bit-identical to casting outside the kernels (x.float() in, gx.to(dt) out)."""
    x = rnd((2, 3, 64, 96), 21).to(DEV).to(dt)
    g = rnd((2, 3, 64, 96), 22).to(DEV)
    for recompute in (False, True):
        m = wmattack.DiffJPEG(True, 64, 96, quality=50)
        m.recompute_backward = recompute
        xa = x.clone().requires_grad_(True)
        ya = m(xa)
        ya.backward(g)
        xb = x.float().requires_grad_(True)
        yb = m(xb)
        yb.backward(g)
        assert ya.dtype == torch.float32 and torch.equal(ya, yb)
        assert xa.grad.dtype == dt and torch.equal(xa.grad, xb.grad.to(dt))
    with torch.no_grad():                                   # forward-only kernel, strided clip slice
        clip = rnd((2, 3, 2, 32, 64), 23).to(DEV).to(dt)
        m = wmattack.DiffJPEG(True, 32, 64, quality=75)
        assert torch.equal(m(clip[:, :, 1]), m(clip[:, :, 1].float()))


@pytest.mark.parametrize("dt", (torch.bfloat16, torch.float16))
def test_jpeg8_family_reads_half_inputs_and_stores_half_gradients(dt):
    """Same boundary property for the 8x8 JPEG family (vector path): bit-identical to casting outside the kernels."""
    x = rnd((2, 3, 40, 64), 24).to(DEV).to(dt)
    g = rnd((2, 3, 40, 64), 25).to(DEV)
    for layer in (wmattack.JpegCompression(DEV), wmattack.JpegMask(50), wmattack.JpegSS(50), wmattack.Jpeg(50)):
        xa = x.clone().requires_grad_(True)
        ya = layer(xa)
        ya.backward(g)
        xb = x.float().requires_grad_(True)
        yb = layer(xb)
        yb.backward(g)
        assert ya.dtype == torch.float32 and torch.equal(ya, yb), type(layer).__name__
        assert xa.grad.dtype == dt and torch.equal(xa.grad, xb.grad.to(dt)), type(layer).__name__
    # a ragged width converts outside (float32 path) and still returns the gradient in the input's type
    xr = rnd((1, 3, 21, 30), 26).to(DEV).to(dt).requires_grad_(True)
    wmattack.JpegMask(50)(xr).backward(rnd((1, 3, 21, 30), 27).to(DEV))
    assert xr.grad.dtype == dt and torch.isfinite(xr.grad.float()).all()


@pytest.mark.parametrize("dt", (torch.bfloat16, torch.float16))
def test_gaussian_noise_reads_half_inputs_and_stores_half_gradients(dt):
    x = rnd((2, 3, 24, 37), 28).to(DEV).to(dt)                  # odd element count: the tail path too
    g = rnd((2, 3, 24, 37), 29).to(DEV)
    noise = (torch.randn(2, 3, 24, 37, generator=torch.Generator().manual_seed(30)) * 0.3).to(DEV)
    for layer in (wmattack.Gaussian(), wmattack.GN(0.0025)):
        kw = {"noise": noise}
        arg = (lambda t: t) if isinstance(layer, wmattack.Gaussian) else (lambda t: (t, None))
        xa = x.clone().requires_grad_(True)
        ya = layer(arg(xa), **kw)
        ya.backward(g)
        xb = x.float().requires_grad_(True)
        yb = layer(arg(xb), **kw)
        yb.backward(g)
        assert ya.dtype == torch.float32 and torch.equal(ya, yb), type(layer).__name__
        assert xa.grad.dtype == dt and torch.equal(xa.grad, xb.grad.to(dt)), type(layer).__name__


@pytest.mark.parametrize("dt", (torch.bfloat16, torch.float16))
def test_ring_layers_read_half_inputs_and_store_half_gradients(dt):
    """The TMA-staged layers at the autocast boundary (models/IRNcrop_model.py:340): a float16 / bfloat16 image is
    staged as it is and widened in the kernel (exact), the gradient is stored in that type - bit-identical to the
    float32 path on the widened image, with torch's round-to-nearest-even cast of its gradient."""
    import wmattack._lib as L
    for shape, seed in (((2, 3, 80, 136), 41), ((1, 3, 37, 264), 42), ((1, 3, 130, 8), 43)):
        x = (rnd(shape, seed) * 1.2 - 0.1).to(DEV).to(dt)
        g = rnd(shape, seed + 100).to(DEV)
        layers = [(wmattack.GaussianBlur(3), {}), (wmattack.GaussianBlur(7), {}), (wmattack.MiddleBlur(3), {}), (wmattack.MiddleBlur(5), {}),
                  (wmattack.Resize(), {"resize_ratio": 0.75}), (wmattack.Resize(), {"resize_ratio": 1.5}),
                  (wmattack.Resize(interpolation_method="bilinear"), {"resize_ratio": 0.5})]
        for layer, kw in layers:
            name = f"{type(layer).__name__} {kw} {shape}"
            n0 = L.launch_count
            xa = x.clone().requires_grad_(True)
            ya = layer(xa, **kw)
            ya.backward(g)
            xb = x.float().requires_grad_(True)
            yb = layer(xb, **kw)
            yb.backward(g)
            assert ya.dtype == torch.float32 and torch.equal(ya, yb), name
            assert xa.grad.dtype == dt and torch.equal(xa.grad, xb.grad.to(dt)), name
    # a frame of a half clip, read in place through its strides; a width off the 16-byte grid converts first
    clip = rnd((2, 3, 3, 40, 64), 44).to(DEV).to(dt)
    for layer in (wmattack.GaussianBlur(3), wmattack.MiddleBlur(3), wmattack.MiddleBlur(5)):
        assert torch.equal(layer(clip[:, :, 1]), layer(clip[:, :, 1].float()))
    odd = rnd((1, 3, 33, 20), 45).to(DEV).to(dt)
    for layer in (wmattack.GaussianBlur(3), wmattack.MiddleBlur(5)):
        xa = odd.clone().requires_grad_(True)
        layer(xa).sum().backward()
        assert torch.equal(layer(odd), layer(odd.float())) and xa.grad.dtype == dt


@pytest.mark.parametrize("dt", (torch.bfloat16, torch.float16))
def test_typed_ring_kernels_stay_inside_their_outputs(dt):
    """compute-sanitizer is not available on the pool: guard bands instead.  Every output of the typed entry points
    (float32 result, arg-median plane, 2-byte gradient) lives in the middle of a sentinel-filled buffer; tiles overhang
    the image in both directions at these shapes; the bands must come back untouched."""
    import ctypes as C
    from wmattack import _lib
    code = {torch.float16: 1, torch.bfloat16: 2}[dt]
    guard = 4096

    def guarded(numel, dtype, fill):
        buf = torch.full((numel + 2 * guard,), fill, device=DEV, dtype=dtype)
        return buf, buf[guard:guard + numel]

    def intact(buf, fill, what):
        assert bool((buf[:guard] == fill).all()) and bool((buf[-guard:] == fill).all()), what

    taps = (C.c_float * 3)(0.25, 0.5, 0.25)
    for n, h, w in ((3, 37, 136), (1, 130, 8), (2, 70, 264)):
        x = (torch.rand(n, h, w, device=DEV) * 1.2 - 0.1).to(dt)
        gy = torch.rand(n, h, w, device=DEV)
        st = torch.cuda.current_stream().cuda_stream
        # blur: typed source -> float32 result; float32 source -> typed result
        ybuf, y = guarded(n * h * w, torch.float32, 7.0)
        _lib.call("wm_gaussblur_typed", x.data_ptr(), code, h * w, w, y.data_ptr(), 0, n, h, w, taps, 3, st)
        intact(ybuf, 7.0, "blur y")
        gbuf, gx = guarded(n * h * w, dt, 7.0)
        _lib.call("wm_gaussblur_typed", gy.data_ptr(), 0, h * w, w, gx.data_ptr(), code, n, h, w, taps, 3, st)
        intact(gbuf, 7.0, "blur gx")
        assert torch.equal(gx.view(n, h, w), wmattack.functional.gaussian_blur(gy[None], (0.25, 0.5, 0.25))[0].to(dt))
        # median: typed source -> float32 result + arg-median plane; float32 cotangent -> typed gradient
        for k in (3, 5):
            idx_sh = -(-w // 16) * 16
            ybuf, y = guarded(n * h * w, torch.float32, 7.0)
            ibuf, idx = guarded(n * h * idx_sh, torch.uint8, 99)
            _lib.call("wm_median_fwd_typed", x.data_ptr(), code, h * w, w, y.data_ptr(), idx.data_ptr(), idx_sh, n, h, w, k, st)
            intact(ybuf, 7.0, f"median{k} y"); intact(ibuf, 99, f"median{k} idx")
            assert bool((idx.view(n, h, idx_sh)[:, :, :w] < k * k).all())
            gbuf, gx = guarded(n * h * w, dt, 7.0)
            _lib.call("wm_median_bwd_typed", gy.data_ptr(), idx.data_ptr(), idx_sh, gx.data_ptr(), code, n, h, w, k, st)
            intact(gbuf, 7.0, f"median{k} gx")
        # fused resize: typed source -> float32 result + clamp mask; float32 cotangent -> typed gradient
        if h >= 16 and w >= 16:
            hm, wm = int(0.75 * h), int(0.75 * w)
            tables = wmattack.functional._resize_tables(torch.device(DEV), h, w, (hm, wm), 1)
            if tables is not None:
                ybuf, y = guarded(n * h * w, torch.float32, 7.0)
                mwords = n * h * 4 * ((w + 127) // 128)
                mbuf, mask = guarded(mwords, torch.int32, 12345)
                _lib.call("wm_resize_fwd_typed", x.data_ptr(), code, h * w, w, y.data_ptr(), n, h, w, hm, wm, 1, mask.data_ptr(), tables.data_ptr(), st)
                intact(ybuf, 7.0, "resize y"); intact(mbuf, 12345, "resize mask")
                gbuf, gx = guarded(n * h * w, dt, 7.0)
                _lib.call("wm_resize_bwd_typed", gy.data_ptr(), mask.data_ptr(), gx.data_ptr(), code, n, h, w, hm, wm, 1, tables.data_ptr(), st)
                intact(gbuf, 7.0, "resize gx")
    torch.cuda.synchronize()


def test_float32_kernels_stay_inside_their_outputs_at_ragged_and_overhanging_shapes():
    """Guard bands around every output of the float32 entry points at shapes where tiles overhang the image and rows are
    off the 16-byte grid (predicated scalar stores, cp.async-fed rings): nothing outside the output may change."""
    import ctypes as C
    from wmattack import _lib
    guard = 4096

    def guarded(numel, dtype, fill):
        buf = torch.full((numel + 2 * guard,), fill, device=DEV, dtype=dtype)
        return buf, buf[guard:guard + numel]

    def intact(buf, fill, what):
        assert bool((buf[:guard] == fill).all()) and bool((buf[-guard:] == fill).all()), what

    taps = (C.c_float * 3)(0.25, 0.5, 0.25)
    jp = wmattack.JpegMask(50)._params
    for n, h, w in ((3, 37, 131), (3, 70, 510), (3, 37, 132), (6, 5, 3), (3, 130, 8)):
        x = torch.rand(n, h, w, device=DEV)
        gy = torch.rand(n, h, w, device=DEV)
        st = torch.cuda.current_stream().cuda_stream
        ybuf, y = guarded(n * h * w, torch.float32, 7.0)
        _lib.call("wm_gaussblur", x.data_ptr(), h * w, w, y.data_ptr(), n, h, w, taps, 3, 0, 0, None, st)
        intact(ybuf, 7.0, f"blur {n, h, w}")
        for k in (3, 5):
            idx_sh = -(-w // 16) * 16
            ybuf, y = guarded(n * h * w, torch.float32, 7.0)
            ibuf, idx = guarded(n * h * idx_sh, torch.uint8, 99)
            _lib.call("wm_median_fwd", x.data_ptr(), h * w, w, y.data_ptr(), idx.data_ptr(), idx_sh, n, h, w, k, None, st)
            intact(ybuf, 7.0, f"median{k} y {n, h, w}"); intact(ibuf, 99, f"median{k} idx {n, h, w}")
            gbuf, gx = guarded(n * h * w, torch.float32, 7.0)
            _lib.call("wm_median_bwd", gy.data_ptr(), idx.data_ptr(), idx_sh, gx.data_ptr(), n, h, w, k, st)
            intact(gbuf, 7.0, f"median{k} gx {n, h, w}")
        if h >= 16 and w >= 16:
            for ratio in (0.75, 1.5):
                hm, wm = int(ratio * h), int(ratio * w)
                tables = wmattack.functional._resize_tables(torch.device(DEV), h, w, (hm, wm), 1)
                if tables is None:
                    continue
                ybuf, y = guarded(n * h * w, torch.float32, 7.0)
                mbuf, mask = guarded(n * h * 4 * ((w + 127) // 128), torch.int32, 12345)
                _lib.call("wm_resize_fwd", x.data_ptr(), h * w, w, y.data_ptr(), n, h, w, hm, wm, 1, mask.data_ptr(), tables.data_ptr(), None, st)
                intact(ybuf, 7.0, f"resize y {n, h, w, ratio}"); intact(mbuf, 12345, f"resize mask {n, h, w, ratio}")
                gbuf, gx = guarded(n * h * w, torch.float32, 7.0)
                _lib.call("wm_resize_bwd", gy.data_ptr(), mask.data_ptr(), gx.data_ptr(), n, h, w, hm, wm, 1, tables.data_ptr(), st)
                intact(gbuf, 7.0, f"resize gx {n, h, w, ratio}")
        if n % 3 == 0:          # the 8x8 JPEG family on [B,3,H,W]: ragged widths take the scalar row path
            b = n // 3
            ybuf, y = guarded(n * h * w, torch.float32, 7.0)
            _lib.call("wm_jpeg8_fwd", x.data_ptr(), 0, 3 * h * w, h * w, w, y.data_ptr(), b, h, w, C.byref(jp), None, st)
            intact(ybuf, 7.0, f"jpeg8 y {n, h, w}")
            gbuf, gx = guarded(n * h * w, torch.float32, 7.0)
            _lib.call("wm_jpeg8_bwd", x.data_ptr(), 3 * h * w, h * w, w, gy.data_ptr(), 3 * h * w, h * w, w, gx.data_ptr(), 0, b, h, w, C.byref(jp), st)
            intact(gbuf, 7.0, f"jpeg8 gx {n, h, w}")
    torch.cuda.synchronize()


def test_typed_ring_entry_points_reject_unaligned_rows():
    import ctypes as C
    from wmattack import _lib
    x = torch.zeros(1, 3, 16, 20, device=DEV, dtype=torch.float16)       # 40-byte rows
    y = torch.zeros(1, 3, 16, 20, device=DEV)
    taps = (C.c_float * 3)(0.25, 0.5, 0.25)
    with pytest.raises(Exception, match="16-byte"):
        _lib.call("wm_gaussblur_typed", x.data_ptr(), 1, 320, 20, y.data_ptr(), 0, 3, 16, 20, taps, 3, None)
    with pytest.raises(Exception, match="16-byte"):
        _lib.call("wm_median_fwd_typed", x.data_ptr(), 1, 320, 20, y.data_ptr(), None, 0, 3, 16, 20, 3, None)


@pytest.mark.parametrize("k", (3, 5))
def test_median_forward_bit_exact(k):
    for xn in ("x2028", "xs32"):
        y = wmattack.MiddleBlur(k)(T(xn).to(DEV))
        assert torch.equal(y.cpu(), T(f"middleblur/k{k}/{xn}/y"))       # kornia semantics (unpinned)
    # 5x5 tiles are 36 rows of a plane PAIR: odd plane counts, heights around the tile size, a shifted bottom tile
    for shape, seed in (((2, 3, 37, 150), 5), ((1, 3, 64, 256), 6), ((1, 1, 3, 2), 7), ((1, 2, 130, 131), 8),
                        ((1, 3, 75, 140), 12), ((1, 1, 36, 132), 13), ((1, 5, 35, 8), 14), ((1, 1, 73, 260), 15)):
        x = rnd(shape, seed) - 0.3                      # negative values too
        y = wmattack.MiddleBlur(k)(x.to(DEV)).cpu()
        assert torch.equal(y, O.median_blur(x, k))
        if shape[2] > 30:                               # arg-median plane too (ties at the zero border)
            y, idx = WF.median_blur_with_index(x.to(DEV), k)
            yo, io = O.median_blur(x, k, return_index=True)
            assert torch.equal(y.cpu(), yo) and torch.equal(idx.cpu(), io)
    # one frame of a [B, 3, T, H, W] clip, read in place through its strides
    clip = rnd((2, 3, 3, 40, 64), 16).to(DEV)
    y = wmattack.MiddleBlur(k)(clip[:, :, 1])
    assert torch.equal(y.cpu(), O.median_blur(clip[:, :, 1].cpu().contiguous(), k))
    xq = torch.round(rnd((1, 3, 48, 160), 9) * 7) / 7   # heavy ties
    y, idx = WF.median_blur_with_index(xq.to(DEV), k)
    yo, io = O.median_blur(xq, k, return_index=True)
    assert torch.equal(y.cpu(), yo) and torch.equal(idx.cpu(), io)


def test_ragged_width_ring_kernels():
    """W % 4 != 0: rows are not 16-byte aligned, no tensor map exists; the SAME ring kernels run fed by cp.async
    (blur, median forward with / without the arg-median plane, median backward through the padded idx plane)."""
    for shape, seed in (((2, 3, 70, 131), 31), ((1, 3, 37, 510), 32), ((1, 3, 5, 3), 33), ((3, 3, 90, 258), 34)):
        x, g = rnd(shape, seed), rnd(shape, seed + 100)
        for k in (3, 5, 7):
            y, gx = fwd_bwd(wmattack.GaussianBlur(k), x, g)
            yo, go = oracle_fwd_bwd(lambda t: O.gaussian_blur(t, k), x, g)
            assert md(y, yo) <= 1e-6 and md(gx, go) <= 1e-6
        for k in (3, 5):
            xq = torch.round(x * 9) / 9                      # ties
            for xx in (x, xq):
                assert torch.equal(wmattack.MiddleBlur(k)(xx.to(DEV)).cpu(), O.median_blur(xx, k))     # no-grad kernel
                y, gx = fwd_bwd(wmattack.MiddleBlur(k), xx, g)
                yo, idx = O.median_blur(xx, k, return_index=True)
                assert torch.equal(y, yo) and torch.equal(gx, O.median_blur_backward(g, idx, k))
        # fused Resize round trip on the same ragged geometry (cp.async-fed instantiation), both directions of scaling
        if min(shape[2:]) >= 20:
            for ratio in (0.75, 1.5):
                m = wmattack.Resize()
                y, gx = fwd_bwd(lambda t: m(t, resize_ratio=ratio), x, g)
                _, go = oracle_fwd_bwd(lambda t: O.resize(t, ratio), x, g)
                # pixels against torch's own fp32 CPU interpolate, the reference's op (same fp32 coordinate arithmetic)
                mid = torch.nn.functional.interpolate(x, size=[int(ratio * shape[2]), int(ratio * shape[3])], mode="bicubic")
                ref = torch.clamp(torch.nn.functional.interpolate(mid, size=list(shape[2:]), mode="bicubic"), 0, 1)
                assert md(y, ref) <= 1e-5, (shape, ratio)
                _assert_resize_grad(gx, go, x, ratio, "bicubic", coord_tol(shape[2], shape[3], ratio, 1e-5), f"ragged resize {shape} r={ratio}")
    # odd row stride: a column slice of a wider tensor
    wide = rnd((1, 3, 40, 203), 35).to(DEV)
    view = wide[..., 3:201]
    for k in (3, 5):
        assert torch.equal(wmattack.MiddleBlur(k)(view).cpu(), O.median_blur(view.cpu().contiguous(), k))
    assert md(wmattack.GaussianBlur(5)(view), O.gaussian_blur(view.cpu().contiguous(), 5)) <= 1e-6


@pytest.mark.parametrize("k", (3, 5))
def test_median_backward(k):
    x, g = rnd((2, 3, 41, 133), 10), rnd((2, 3, 41, 133), 11)
    y, gx = fwd_bwd(wmattack.MiddleBlur(k), x, g)
    _, idx = O.median_blur(x, k, return_index=True)
    assert torch.equal(gx, O.median_blur_backward(g, idx, k))            # pure routing: bit-exact
    # tie-free input: torch autograd through the kornia formulation agrees
    xx = x.double().requires_grad_(True)
    O.median_windows(xx, k).median(dim=2)[0].backward(g.double())
    assert md(gx, xx.grad) <= 1e-6
    # tied input: the tie-invariant (gradient mass is conserved per window) holds
    xq = torch.round(x * 5) / 5
    y, gxq = fwd_bwd(wmattack.MiddleBlur(k), xq, g)
    _, idxq = O.median_blur(xq, k, return_index=True)
    assert torch.equal(gxq, O.median_blur_backward(g, idxq, k))
    # W % 4 == 0 but W % 16 != 0: TMA-fed rings, the arg-median plane's rows are padded to 16 bytes
    for shape, seed in (((1, 3, 50, 132), 12), ((2, 1, 9, 20), 13)):
        xs, gs = torch.round(rnd(shape, seed) * 11) / 11, rnd(shape, seed + 50)
        y, gxs = fwd_bwd(wmattack.MiddleBlur(k), xs, gs)
        yo, idxs = O.median_blur(xs, k, return_index=True)
        assert torch.equal(y, yo) and torch.equal(gxs, O.median_blur_backward(gs, idxs, k))


# =============================================================================== elementwise
def test_elementwise_with_injected_random_tensors():
    x, g, cover = T("x32"), T("g32"), T("cover32")
    noise = T("gaussian/x32/noise").to(DEV)
    y, gx = fwd_bwd(lambda t: wmattack.Gaussian()(t, noise=noise), x, g)
    assert md(y, T("gaussian/x32/y")) <= 1e-7 and md(gx, T("gaussian/x32/gx")) == 0
    y = wmattack.GN(0.0025)((x.to(DEV), None), noise=T("gn/x32/noise").to(DEV))
    assert md(y, T("gn/x32/y")) <= 1e-7
    rdn = T("saltpepper/p0.1/x32/rdn").to(DEV)
    y, gx = fwd_bwd(lambda t: wmattack.SaltPepper(0.1)(t, rdn=rdn), x, g)
    assert md(y, T("saltpepper/p0.1/x32/y")) == 0 and md(gx, T("saltpepper/p0.1/x32/gx")) == 0
    y = wmattack.ElementDropout(0.5)((x.to(DEV), cover.to(DEV)), rdn=T("cropdropout/p0.5/x32/rdn").to(DEV))
    assert md(y, T("cropdropout/p0.5/x32/y")) == 0
    np.random.seed(16)       # the layer draws np.random.uniform like the reference (dropout.py:19)
    y = wmattack.MaskDropout((0.5, 1))(x.to(DEV), cover.to(DEV), mask=T("maskdropout/x32/mask").float().to(DEV))
    assert md(y, T("maskdropout/x32/y")) <= 1e-7
    y, gx = fwd_bwd(wmattack.Quantization(), x, g)
    assert md(y, T("quantization/x32/y")) == 0 and md(gx, g) == 0


def test_elementwise_host_rng_reproduces_reference_stream():
    x = T("x32").to(DEV)
    torch.manual_seed(13)
    assert md(wmattack.SaltPepper(0.1, host_rng=True)(x), T("saltpepper/p0.1/x32/y")) == 0
    np.random.seed(12)
    assert md(wmattack.GN(0.0025, host_rng=True)((x, x)), T("gn/x32/y")) <= 1e-7
    np.random.seed(16)
    assert md(wmattack.MaskDropout((0.5, 1), host_rng=True)(x, T("cover32").to(DEV)), T("maskdropout/x32/y")) <= 1e-7
    torch.manual_seed(15)
    assert md(wmattack.ElementDropout(0.5, host_rng=True)((x, T("cover32").to(DEV))), T("cropdropout/p0.5/x32/y")) == 0


def test_elementwise_philox_statistics_and_gradients():
    torch.manual_seed(0)
    x = torch.full((8, 3, 256, 256), 0.5, device=DEV)
    n = x.numel()
    xx = x.clone().requires_grad_(True)
    y = wmattack.Gaussian()(xx, stddev=0.05)
    d = (y - x).double()
    assert abs(float(d.mean())) < 2e-4 and abs(float(d.std()) - 0.05) < 2e-4
    # kurtosis of a normal is 3
    assert abs(float(((d / d.std()) ** 4).mean()) - 3.0) < 0.05
    y.backward(torch.ones_like(y))
    assert float(xx.grad.min()) == 1.0                       # nothing clamped at x = 0.5, sigma 0.05
    xe = torch.zeros((4, 3, 64, 64), device=DEV, requires_grad=True)
    ye = wmattack.Gaussian()(xe)
    ye.backward(torch.ones_like(ye))
    frac = float(xe.grad.mean())
    assert 0.45 < frac < 0.55
    assert torch.equal((ye.detach() > 0), (xe.grad == 1) & (ye.detach() > 0))
    y2 = wmattack.Gaussian()(x)
    assert not torch.equal(y, y2)                            # successive calls draw new noise
    sp = wmattack.SaltPepper(0.2)(x)
    assert abs(float((sp == 0).float().mean()) - 0.1) < 3e-3 and abs(float((sp == 1).float().mean()) - 0.1) < 3e-3
    cover = torch.zeros_like(x)
    dr = wmattack.ElementDropout(0.3)((x, cover))
    assert abs(float((dr == 0).float().mean()) - 0.7) < 3e-3
    np.random.seed(1)
    md_ = wmattack.MaskDropout((0.6, 0.6))(x, cover)
    kept = (md_ == 0.5).float()
    assert abs(float(kept.mean()) - 0.6) < 1e-2
    assert torch.equal(kept[0, 0], kept[5, 2])                # one [H,W] mask for batch and channels
    gnz = wmattack.GN(0.01)((x, x))
    assert abs(float((gnz - x).std()) - 0.1) < 1e-3 and float(gnz.max()) > 1.0 - 0.5 + 0.3
    assert n > 0


def test_cropout_and_dropout_gradients():
    x, c = rnd((2, 3, 16, 24), 1).to(DEV).requires_grad_(True), rnd((2, 3, 16, 24), 2).to(DEV).requires_grad_(True)
    y = wmattack.Cropout(0.5, 0.5)((x, c), box=(2, 10, 4, 20))
    ref = O.cropout(x.detach().cpu(), c.detach().cpu(), (2, 10, 4, 20))
    assert md(y, ref) == 0
    y.sum().backward()
    assert float(x.grad.sum()) == 2 * 3 * 8 * 16 and float(c.grad.sum()) == 2 * 3 * (16 * 24 - 8 * 16)
    m = torch.zeros(16, 24, device=DEV)
    m[:, :12] = 1
    x.grad = c.grad = None
    WF.dropout_mask(x, c, m).sum().backward()
    assert float(x.grad[..., :12].min()) == 1 and float(x.grad[..., 12:].max()) == 0
    assert float(c.grad[..., :12].max()) == 0 and float(c.grad[..., 12:].min()) == 1
    # odd plane size (H*W % 4 != 0: planes do not start on 16-byte boundaries): mask dropout and splice, values + gradients
    xo, co, go = rnd((2, 3, 7, 9), 3), rnd((2, 3, 7, 9), 4), rnd((2, 3, 7, 9), 5)
    mo = (rnd((7, 9), 6) > 0.5).float()
    a, b = xo.to(DEV).requires_grad_(True), co.to(DEV).requires_grad_(True)
    y = WF.dropout_mask(a, b, mo.to(DEV))
    y.backward(go.to(DEV))
    assert torch.equal(y.detach().cpu(), O.dropout_mask(xo, co, mo))
    assert torch.equal(a.grad.cpu(), go * mo) and torch.equal(b.grad.cpu(), go * (1 - mo))
    ms = (rnd((2, 1, 7, 9), 7) > 0.7).float()
    a, b = xo.to(DEV).requires_grad_(True), co.to(DEV).requires_grad_(True)
    out = wmattack.Splice()(a, b, ms.to(DEV))
    out.backward(go.to(DEV))
    assert torch.equal(out.detach().cpu(), xo * (1 - ms) + co * ms)
    assert torch.equal(a.grad.cpu(), go * (1 - ms)) and torch.equal(b.grad.cpu(), go * ms)


# ============================================================================= resize / crop
@pytest.mark.parametrize("mode", ("bicubic", "bilinear"))
@pytest.mark.parametrize("r", (0.5, 0.7, 1.3, 1.5))
def test_resize_vs_golden(mode, r):
    m = wmattack.Resize(interpolation_method=mode)
    y, gx = fwd_bwd(lambda t: m(t, resize_ratio=r), T("x2028"), T("gaussianblur/g2028"))
    assert md(y, T(f"resize/{mode}/r{r}/x2028/y")) <= 1e-5
    assert md(gx, T(f"resize/{mode}/r{r}/x2028/gx")) <= 1e-5
    assert m.name == "Resize"


def _assert_resize_grad(gx, g_ref, x, ratio, mode, tol, what, mid=None):
    """Resize gradient check: every mismatch must lie in the support of the transposed round-trip operator
    applied to the clamp-FRAGILE outputs (fp64 pre-clamp value within 1e-5 of 0 or 1)."""
    h, w = x.shape[2:]
    mid = mid or O.resize_mid_size(h, w, ratio)
    infl, n_fragile, _ = P.resize_grad_influence(x, mid, mode)
    P.assert_explained(gx, g_ref, tol, infl, what)
    return n_fragile


def test_resize_saturated_random_ratio_and_large():
    m = wmattack.Resize()
    y, gx = fwd_bwd(lambda t: m(t, resize_ratio=0.8), T("xsat"), T("g32"))
    assert md(y, T("resize/bicubic/r0.8/xsat/y")) <= 1e-5
    # binary input: the bicubic round trip overshoots [0,1] almost everywhere, so the clamp mask decides most of the
    # gradient; no pre-clamp value of this fixture is within 1e-5 of a bound, hence NO mismatch is tolerated
    n = _assert_resize_grad(gx, T("resize/bicubic/r0.8/xsat/gx"), T("xsat"), 0.8, "bicubic", 1e-5, "xsat gx vs reference")
    assert n == 0
    np.random.seed(17)
    assert md(m(T("x32").to(DEV)), T("resize/random/x32/y")) <= 1e-5
    # larger, non-square, against torch's own fp32 CPU interpolate (the reference's op)
    x = rnd((2, 3, 96, 160), 12)
    for mode in ("bicubic", "bilinear"):
        for r in (0.53, 0.91, 1.27):
            mid = torch.nn.functional.interpolate(x, size=[int(r * 96), int(r * 160)], mode=mode)
            ref = torch.clamp(torch.nn.functional.interpolate(mid, size=[96, 160], mode=mode), 0, 1)
            y = wmattack.Resize(interpolation_method=mode)(x.to(DEV), resize_ratio=r)
            assert md(y, ref) <= 1e-5
    g = rnd((2, 3, 96, 160), 13)
    y, gx = fwd_bwd(lambda t: m(t, resize_ratio=0.77), x, g)
    yo, go = oracle_fwd_bwd(lambda t: O.resize(t, 0.77), x, g)
    assert md(y, yo) <= 2e-5
    n = _assert_resize_grad(gx, go, x, 0.77, "bicubic", coord_tol(96, 160, 0.77, 1e-5), "gx vs fp64 oracle")
    assert n <= 1e-3 * x.numel()


def test_interp_untiled_fallback_scales():
    """Scale factors outside [0.4, 2.2] take the per-element kernels (and their workspace)."""
    x, g = rnd((1, 3, 40, 56), 21), rnd((1, 3, 40, 56), 22)
    for r in (0.3, 2.6):
        for mode in ("bicubic", "bilinear"):
            m = wmattack.Resize(interpolation_method=mode)
            y, gx = fwd_bwd(lambda t: m(t, resize_ratio=r), x, g)
            yo, go = oracle_fwd_bwd(lambda t: O.resize(t, r, mode), x, g)
            assert md(y, yo) <= 1e-5
            _assert_resize_grad(gx, go, x, r, mode, coord_tol(40, 56, r, 1e-5), f"untiled r={r} {mode}")
    y, apex = wmattack.Crop()(x.to(DEV), apex=(4, 14, 6, 20))          # 4x upsampling
    assert md(y, O.crop_resize(x.double(), (4, 14, 6, 20))) <= 1e-5
    xx = x.to(DEV).requires_grad_(True)
    wmattack.Crop()(xx, apex=(4, 14, 6, 20))[0].backward(g.to(DEV))
    _, go = oracle_fwd_bwd(lambda t: O.crop_resize(t, (4, 14, 6, 20)), x, g)
    assert md(xx.grad, go) <= 1e-5


def test_crop():
    x, g = T("x32"), T("g32")
    np.random.seed(18)
    xx = x.to(DEV).requires_grad_(True)
    y, apex = wmattack.Crop()(xx)
    y.backward(g.to(DEV))
    assert tuple(apex) == tuple(int(v) for v in GOLD["crop/seed18/x32/apex"])
    assert md(y, T("crop/seed18/x32/y")) <= 1e-5 and md(xx.grad, T("crop/seed18/x32/gx")) <= 1e-5
    np.random.seed(19)
    y, apex = wmattack.Crop()(T("x2028").to(DEV), min_rate=0.7, max_rate=0.9)
    assert tuple(apex) == tuple(int(v) for v in GOLD["crop/seed19/x2028/apex"])
    assert md(y, T("crop/seed19/x2028/y")) <= 1e-5
    y, apex = wmattack.Crop()(x.to(DEV), apex=(3, 20, 5, 31))
    assert md(y, T("crop/apex/x32/y")) <= 1e-5 and apex == (3, 20, 5, 31)
    np.random.seed(20)
    outs = wmattack.Crop().cropped_out(x.to(DEV), min_rate=0.5)
    assert md(outs[0], T("cropped_out/seed20/x32/scaled")) <= 1e-5
    assert md(outs[1], T("cropped_out/seed20/x32/zero_images")) <= 1e-5
    assert md(outs[2], T("cropped_out/seed20/x32/mask")) == 0
    assert np.allclose(np.array(outs[3]), GOLD["cropped_out/seed20/x32/apex"])
    assert md(outs[4], T("cropped_out/seed20/x32/new_images")) == 0


@pytest.mark.parametrize("mode", ("bilinear", "bicubic"))
def test_crop_fast_path_geometry_sweep(mode):
    """wm_cropresize_* (TMA source box + banded tables) over crop rectangles at every edge / corner, odd
    origins, rates 0.46 .. 1.0, several frame sizes: forward against torch's own interpolate of the
    slice (CPU fp32, the reference's op) and backward against autograd through it; the gradient
    outside the rectangle must be exactly zero.  Ineligible geometries (rate < 0.45) fall back."""
    cases = [((2, 3, 64, 128), [(0, 64, 0, 128), (0, 32, 0, 64), (31, 64, 63, 128), (5, 37, 17, 90), (0, 30, 61, 128),
                                (13, 64, 0, 59), (1, 62, 3, 127), (7, 40, 9, 80)]),
             ((1, 3, 200, 264), [(0, 100, 0, 132), (99, 200, 131, 264), (11, 199, 40, 263), (50, 143, 7, 131)]),
             ((1, 2, 256, 256), [(0, 118, 0, 118), (100, 228, 37, 165), (3, 250, 5, 200)])]
    for shape, boxes in cases:
        x, g = rnd(shape, 31), rnd(shape, 32)
        h, w = shape[2:]
        for (a, b, c, d) in boxes:
            xr = x.clone().requires_grad_(True)
            yr = torch.nn.functional.interpolate(xr[:, :, a:b, c:d], size=[h, w], mode=mode)
            yr.backward(g)
            xt = x.to(DEV).requires_grad_(True)                       # the same ATen op on this device
            yt = torch.nn.functional.interpolate(xt[:, :, a:b, c:d], size=[h, w], mode=mode)
            yt.backward(g.to(DEV))
            xx = x.to(DEV).requires_grad_(True)
            y = WF.interpolate(xx, (h, w), mode, window=(a, c, b - a, d - c))
            y.backward(g.to(DEV))
            assert md(y, yr) <= 2e-6 and md(y, yt) <= 2e-6, (shape, (a, b, c, d))
            # gradients: tight against the same-device op; at non-dyadic scales fp32 source coordinates that land
            # within an ulp of an integer move a tap by one sample in ANY fp32 implementation (torch's own CPU
            # and CUDA gradients differ from fp64 by 2.5e-5 on the 264-wide cases), hence the wider CPU bound
            assert md(xx.grad, xt.grad) <= 5e-6, (shape, (a, b, c, d))
            assert md(xx.grad, xr.grad) <= 5e-5, (shape, (a, b, c, d))
            # exact adjoint of our own forward: <A x, g> == <x, A^T g>
            lhs = float((y.detach().double() * g.to(DEV).double()).sum())
            rhs = float((xx.grad.double() * x.to(DEV).double()).sum())
            assert abs(lhs - rhs) <= 1e-7 * abs(lhs)          # fp32 rounding of the two sums only
            outside = xx.grad.cpu().clone()
            outside[:, :, a:b, c:d] = 0
            assert float(outside.abs().max()) == 0.0
    # the module itself, reference RNG order, at the trainers' frame size
    np.random.seed(77)
    xx = rnd((2, 3, 256, 256), 33).to(DEV).requires_grad_(True)
    y, apex = wmattack.Crop()(xx)
    h0, h1, w0, w1 = apex
    ref = torch.nn.functional.interpolate(xx.detach().cpu()[:, :, h0:h1, w0:w1], size=[256, 256], mode="bilinear")
    assert md(y, ref) <= 2e-6


def test_crop_random_rectangles_stress():
    """Sixty random frame sizes / rectangles (rates 0.45 .. 1, odd origins and extents, widths not a
    multiple of 4 -> older kernels) through Crop's interpolate against the same ATen op on this device."""
    rs = np.random.RandomState(123)
    for it in range(60):
        h, w = int(rs.randint(12, 200)), int(rs.randint(3, 60)) * 4 + (0 if it % 5 else int(rs.randint(0, 4)))
        hin, win = max(int(h * rs.uniform(0.45, 1.0)), 4), max(int(w * rs.uniform(0.45, 1.0)), 4)
        a, c = int(rs.randint(0, h - hin + 1)), int(rs.randint(0, w - win + 1))
        mode = "bilinear" if it % 3 else "bicubic"
        x, g = rnd((1, 2, h, w), 1000 + it).to(DEV), rnd((1, 2, h, w), 2000 + it).to(DEV)
        xt = x.clone().requires_grad_(True)
        yt = torch.nn.functional.interpolate(xt[:, :, a:a + hin, c:c + win], size=[h, w], mode=mode)
        yt.backward(g)
        xx = x.clone().requires_grad_(True)
        y = WF.interpolate(xx, (h, w), mode, window=(a, c, hin, win))
        y.backward(g)
        assert md(y, yt) <= 3e-6, (it, h, w, a, c, hin, win, mode)
        assert md(xx.grad, xt.grad) <= 1e-5, (it, h, w, a, c, hin, win, mode)


def test_crop_full_frame_sizes_adjoint_and_reference():
    """Config-3 / config-5 frame sizes: Crop's resize against the same-device ATen op, the exact-adjoint
    property <A x, g> = <x, A^T g>, zero gradient outside the rectangle, run-to-run determinism."""
    for (h, w, box) in ((1080, 1920, (100, 900, 240, 1700)), (2160, 3840, (1000, 2160, 0, 2000))):
        a, b, c, d = box
        x, g = rnd((1, 3, h, w), 5).to(DEV), rnd((1, 3, h, w), 6).to(DEV)
        xt = x.clone().requires_grad_(True)
        yt = torch.nn.functional.interpolate(xt[:, :, a:b, c:d], size=[h, w], mode="bilinear")
        yt.backward(g)
        outs = []
        for _ in range(2):
            xx = x.clone().requires_grad_(True)
            y = wmattack.Crop()(xx, apex=box)[0]
            y.backward(g)
            outs.append((y.detach(), xx.grad))
        y, gx = outs[0]
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert md(y, yt) <= 2e-6 and md(gx, xt.grad) <= 1e-5
        lhs, rhs = float((y.double() * g.double()).sum()), float((gx.double() * x.double()).sum())
        assert abs(lhs - rhs) <= 1e-7 * abs(lhs)
        outside = gx.clone()
        outside[:, :, a:b, c:d] = 0
        assert float(outside.abs().max()) == 0.0


def test_crop_fast_path_is_used_and_deterministic():
    import wmattack._lib as L
    assert L.load().wm_cropresize_ok(180, 192, 256, 256, 6, 0) == 1
    assert L.load().wm_cropresize_ok(100, 192, 256, 256, 6, 0) == 0          # rate 0.39: older kernels
    x, g = rnd((2, 3, 128, 128), 41).to(DEV), rnd((2, 3, 128, 128), 42).to(DEV)
    outs = []
    for _ in range(2):
        xx = x.clone().requires_grad_(True)
        y = wmattack.Crop()(xx, apex=(10, 100, 20, 110))[0]
        y.backward(g)
        outs.append((y.detach(), xx.grad))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


# =================================================================================== Combined
def test_combined_matches_reference_choices():
    import random
    random.seed(21)
    np.random.seed(22)
    comb = wmattack.Combined([wmattack.Identity(), wmattack.Jpeg(50), wmattack.JpegSS(70), wmattack.JpegMask(30),
                              wmattack.Resize()])
    assert comb.name == "NotChosenYet"
    x = T("x20").to(DEV)
    names = []
    for _ in range(12):
        comb(x)
        names.append(comb.name)
    assert ",".join(names) == text("combined/seed21/names")
    assert wmattack.Identity()(x) is x
    # BASELINE config 2's Combined runs every member by id
    c2 = wmattack.Combined([wmattack.JpegCompression(DEV), wmattack.GaussianBlur(), wmattack.MiddleBlur(5),
                            wmattack.Gaussian(), wmattack.Resize()])
    x = rnd((2, 3, 64, 64), 3).to(DEV)
    # Combined copies .name BEFORE the call (noise_layers/combined.py:19), so GaussianBlur still
    # reports its constructor name "G_Blur" on first use (gaussian_blur.py:15 vs :54)
    for i, nm in enumerate(("JpegCompression", "G_Blur", "MiddleBlur5", "Gaussian", "Resize")):
        y = c2(x, id=i)
        assert c2.name == nm and y.shape == x.shape


# ==================================================== round-1 additions: TMA stencils / fused resize
def _resize_tables_overflow():
    """1 if any fused-resize geometry used so far lost a band weight (window too small) — must stay 0.  Every
    geometry is proven on a blank plane before it enters the cache (functional._resize_tables); one that does not
    fit would be listed in _RESIZE_UNFUSED and served by the two-call path: the sweeps below must not produce any."""
    assert not WF._RESIZE_UNFUSED, f"geometries outside the band bound: {sorted(WF._RESIZE_UNFUSED)[:5]}"
    return max((int(t[-4:].view(torch.int32)[0]) for t in WF._RESIZE_TABLES.values()), default=0)


@pytest.mark.parametrize("shape", [(1, 3, 1080, 1920), (2, 3, 540, 964), (1, 1, 70, 132), (3, 2, 64, 128)])
def test_tma_stencils_at_frame_sizes(shape):
    """BASELINE config 3 shapes (1080p frames) and ragged tiles through the TMA-fed blur / median kernels."""
    x, g = rnd(shape, 31), rnd(shape, 32)
    c = shape[1]
    for k in (3, 7):
        y, gx = fwd_bwd(wmattack.GaussianBlur(k, channels=c), x, g)
        yo, go = oracle_fwd_bwd(lambda t: O.gaussian_blur(t, k), x, g)
        assert md(y, yo) <= 1e-6 and md(gx, go) <= 1e-6
    for k in (3, 5):
        y, gx = fwd_bwd(wmattack.MiddleBlur(k), x, g)
        yo, idx = O.median_blur(x, k, return_index=True)
        assert torch.equal(y, yo)                                          # selection: bit-exact
        assert torch.equal(gx, O.median_blur_backward(g, idx, k))         # routing: bit-exact


def test_stencils_and_resize_on_clip_slices():
    """The trainers pass x[:, :, t] slices of a [B,3,T,H,W] clip (models/IRNcrop_model.py:362-366):
    plane strides go straight into the tensor maps, no .contiguous()."""
    clip = rnd((2, 3, 4, 48, 160), 33)
    g = rnd((2, 3, 48, 160), 34)
    for t in (0, 3):
        xs = clip[:, :, t]
        xd = clip.to(DEV)[:, :, t]
        assert not xd.is_contiguous()
        for layer, ref in ((wmattack.GaussianBlur(), lambda v: O.gaussian_blur(v, 3)),
                           (wmattack.MiddleBlur(3), lambda v: O.median_blur(v, 3)),
                           (lambda v: wmattack.Resize()(v, resize_ratio=0.8), lambda v: O.resize(v, 0.8))):
            xx = xd.detach().requires_grad_(True)
            y = layer(xx)
            y.backward(g.to(DEV))
            yo, go = oracle_fwd_bwd(ref, xs.contiguous(), g)
            assert md(y, yo) <= 2e-5
            if layer.__class__.__name__ == "function":                  # the Resize lambda: clamp mask
                _assert_resize_grad(xx.grad, go, xs.contiguous(), 0.8, "bicubic", coord_tol(48, 160, 0.8, 1e-5), "clip slice resize gx")
            else:
                assert md(xx.grad, go) <= 2e-5
    assert _resize_tables_overflow() == 0


@pytest.mark.parametrize("mode", ("bicubic", "bilinear"))
def test_fused_resize_geometry_sweep(mode):
    """Every window size (8/10/12/14), ragged tiles, borders, both axes with different ratios."""
    for (h, w), seed in (((70, 132), 41), ((64, 128), 42), ((130, 260), 43), ((33, 36), 44)):
        x, g = rnd((2, 2, h, w), seed), rnd((2, 2, h, w), seed + 100)
        for r in (0.46, 0.5, 0.58, 0.66, 0.75, 0.9, 1.0, 1.1, 1.5, 2.0, 2.19):
            mid = (max(int(r * h), 1), max(int(r * w), 1))
            in_range = all(0.45 <= m / n <= 2.2 for m, n in zip(mid, (h, w)))     # else: two-call fallback
            assert WF._lib.load().wm_resize_is_fused(h, w, mid[0], mid[1], 4) == int(in_range)
            xx = x.to(DEV).requires_grad_(True)
            y = WF.resize_roundtrip(xx, mid, mode)
            y.backward(g.to(DEV))
            # Comparators: the reference's own op, torch fp32 F.interpolate (noise_layers/resize.py:38-47),
            # (1) on the CPU and (2) on this GPU — the trainers run it on CUDA.  ATen evaluates the source
            # coordinate scale*(o+0.5)-0.5 in fp32, so at non-dyadic scales its taps carry an error of
            # ~ulp(coordinate) that differs between builds (FMA contraction): two correct fp32
            # implementations agree only to that — the CPU bound scales with it, the same-device one is tight.
            Fi = torch.nn.functional.interpolate
            coord_ulp = 2.0 ** (np.floor(np.log2(max(h, w) * max(1.0, 1.0 / r))) - 23)
            for dev, tol_y, tol_g in (("cpu", 1e-5 + 2 * coord_ulp, 2e-5 + 2 * coord_ulp), (DEV, 1e-5, 2e-5)):
                xo = x.detach().clone().to(dev).requires_grad_(True)
                pre = Fi(Fi(xo, size=list(mid), mode=mode), size=[h, w], mode=mode)
                yo = pre.clamp(0, 1)
                yo.backward(g.to(dev))
                assert md(y, yo) <= tol_y, (h, w, r, dev)
                # gradient: where the fp64 pre-clamp value is within fp32 noise of a clamp bound the pass-through
                # decision of ANY fp32 implementation is arbitrary: every mismatch must trace back to such an output
                _assert_resize_grad(xx.grad, xo.grad, x, r, mode, tol_g, f"{(h, w, r, dev)}", mid=mid)
    assert _resize_tables_overflow() == 0


def test_fused_resize_full_size_adjoint_and_determinism():
    """BASELINE config-2 size: with the clamp inactive the layer is linear, so <A x, g> = <x, A^T g>;
    the adjoint is a gather (no atomics), so two runs are bit-identical."""
    gen = torch.Generator(DEV).manual_seed(5)
    x = (0.4 + 0.2 * torch.rand(64, 3, 512, 512, device=DEV, generator=gen)).requires_grad_(True)
    g = torch.rand(64, 3, 512, 512, device=DEV, generator=gen)
    for r in (0.5, 0.75, 1.25, 1.5):
        x.grad = None
        y = wmattack.Resize()(x, resize_ratio=r)
        assert float(y.min()) > 0 and float(y.max()) < 1              # clamp inactive
        y.backward(g)
        g1 = x.grad.clone()
        lhs = float((y.double() * g.double()).sum())
        rhs = float((x.detach().double() * g1.double()).sum())
        assert abs(lhs - rhs) <= 1e-6 * abs(lhs)
        x.grad = None
        wmattack.Resize()(x, resize_ratio=r).backward(g)
        assert torch.equal(g1, x.grad)
    assert _resize_tables_overflow() == 0


def test_diffjpeg_4k_quality_sweep_mcu_independence():
    """BASELINE config 5 shape (one 4K frame per call here): every 16x16 MCU is independent, so any
    MCU-aligned crop of the 4K result equals the oracle run on that crop alone."""
    gen = torch.Generator(DEV).manual_seed(7)
    x = torch.rand(1, 3, 2160, 3840, device=DEV, generator=gen)
    g = torch.rand(1, 3, 2160, 3840, device=DEV, generator=gen)
    for q in (10, 50, 95):
        xx = x.clone().requires_grad_(True)
        m = wmattack.DiffJPEG(True, 2160, 3840, quality=q)
        y = m(xx)
        y.backward(g)
        assert torch.isfinite(y).all() and float(y.min()) >= 0 and float(y.max()) <= 1
        for (r0, c0) in ((0, 0), (1088, 2048), (2160 - 64, 3840 - 96)):
            xc = x[:, :, r0:r0 + 64, c0:c0 + 96].cpu().double().requires_grad_(True)
            yo = O.diffjpeg(xc, q)
            yo.backward(g[:, :, r0:r0 + 64, c0:c0 + 96].cpu().double())
            xcrop, gcrop = x[:, :, r0:r0 + 64, c0:c0 + 96].cpu(), g[:, :, r0:r0 + 64, c0:c0 + 96].cpu()
            infl, n_fragile = P.diffjpeg_influence(xcrop, O.quality_to_factor(q), 0)
            assert n_fragile <= 1e-3 * xcrop.numel() + 2
            P.assert_explained(y[:, :, r0:r0 + 64, c0:c0 + 96], yo, 1e-5, infl, "4K y")
            tol = grad_tol(q, lambda t: O.diffjpeg(t, q), xcrop, gcrop, infl)
            P.assert_explained(xx.grad[:, :, r0:r0 + 64, c0:c0 + 96], xc.grad, tol, infl, "4K gx")


# ======================================================= SURVEY 8f "next" rows: epilogue, bank, splice
def test_attack_epilogue_bank_and_splice_match_the_trainer_arithmetic():
    """models/IRNp_model.py:609-680 (8-way attack, clamp, straight-through, Quantization) and
    models/IRNcrop_model.py:348 (splice), restated with plain torch ops on the CPU: values are
    bit-identical, gradients are the straight-through identities."""
    x, prev = rnd((2, 3, 40, 64), 51), rnd((2, 3, 40, 64), 52)
    mask = (rnd((2, 1, 40, 64), 53) > 0.88).float()
    g = rnd((2, 3, 40, 64), 54)
    # splice
    a = x.to(DEV).requires_grad_(True); b = prev.to(DEV).requires_grad_(True)
    out = wmattack.Splice()(a, b, mask.to(DEV))
    out.backward(g.to(DEV))
    ac = x.clone().requires_grad_(True); bc = prev.clone().requires_grad_(True)
    ref = ac * (1 - mask) + bc * mask
    ref.backward(g)
    assert torch.equal(out.detach().cpu(), ref.detach())
    assert torch.equal(a.grad.cpu(), ac.grad) and torch.equal(b.grad.cpu(), bc.grad)
    # epilogue on an attacked batch that leaves [0,1]
    sim = (x + 0.6 * (rnd((2, 3, 40, 64), 55) - 0.5)) * 1.2 - 0.1
    xx = x.to(DEV).requires_grad_(True)
    y = wmattack.AttackEpilogue()(xx, sim.to(DEV))
    y.backward(g.to(DEV))
    xc = x.clone().requires_grad_(True)
    att = xc + (torch.clamp(sim, 0, 1) - xc).detach()
    refq = (att * 255.).round() / 255.                        # Quant.forward, models/modules/Quantization.py:9
    assert torch.equal(y.detach().cpu(), refq.detach())
    assert torch.equal(xx.grad.cpu(), g)                       # identity (Quant.backward + straight-through)
    # K-way bank: same values as cat([...]) of the per-layer epilogues, gradient = sum over the K slices
    np.random.seed(3)
    layers = [wmattack.Resize(), wmattack.JpegMask(50), wmattack.MiddleBlur(3), wmattack.GaussianBlur(), wmattack.Identity()]
    bank = wmattack.AttackBank(layers)
    xx = x.to(DEV).requires_grad_(True)
    np.random.seed(3)
    yb = bank(xx)
    assert yb.shape == (10, 3, 40, 64) and bank.names[1] == "JpegMask50"
    gk = rnd((10, 3, 40, 64), 56)
    yb.backward(gk.to(DEV))
    np.random.seed(3)
    parts = []
    for layer in layers:
        s = layer(x.to(DEV))
        parts.append(wmattack.AttackEpilogue()(x.to(DEV), s))
    assert torch.equal(yb.detach(), torch.cat(parts, 0))
    assert md(xx.grad, gk.view(5, 2, 3, 40, 64).sum(0)) <= 1e-6
    # odd element count per image: the slices of the K-way batch do not start on 16-byte boundaries (scalar epilogue / sum)
    xo = rnd((1, 3, 7, 9), 57)
    lo = [wmattack.JpegMask(50), wmattack.MiddleBlur(3), wmattack.GaussianBlur(), wmattack.Identity()]
    xa = xo.to(DEV).requires_grad_(True)
    yo = wmattack.AttackBank(lo)(xa)
    go = rnd((4, 3, 7, 9), 58)
    yo.backward(go.to(DEV))
    ref = []
    for layer in lo:
        sim = layer(xo.to(DEV)).cpu()
        ref.append(((xo + (torch.clamp(sim, 0, 1) - xo)) * 255.).round() / 255.)
    assert torch.equal(yo.detach().cpu(), torch.cat(ref, 0))
    assert md(xa.grad, go.sum(0, keepdim=True)) <= 1e-6


def test_attack_mix_matches_the_trainer_arithmetic():
    """Hybrid attack of models/IRNcrop_model.py:357-373 (as intended): softmax-weighted mix of the attacked versions,
    clamp_with_grad, Quantization — bit-identical values to the torch expression, gradients alpha_k * gy."""
    for shape, seed in (((3, 3, 16, 24), 71), ((2, 3, 7, 9), 72)):          # the second: odd element count (scalar path)
        ys = [(rnd(shape, seed + k) * 1.4 - 0.2) for k in range(5)]
        alpha = torch.softmax(torch.randn(shape[0], 5, generator=torch.Generator().manual_seed(seed)), dim=1)
        g = rnd(shape, seed + 9)
        yd = [y.to(DEV).requires_grad_(k != 2) for k, y in enumerate(ys)]    # one member needs no gradient
        out = wmattack.AttackMix()(yd, alpha.to(DEV))
        out.backward(g.to(DEV))
        yc = [y.clone().requires_grad_(True) for y in ys]
        mixed = None
        for k in range(5):
            term = alpha[:, k].view(-1, 1, 1, 1) * yc[k]
            mixed = term if mixed is None else mixed + term
        mixed = mixed + (torch.clamp(mixed, 0, 1) - mixed).detach()
        ref = (mixed * 255.).round() / 255.
        assert torch.equal(out.detach().cpu(), ref.detach())
        for k in range(5):
            want = alpha[:, k].view(-1, 1, 1, 1) * g
            if k == 2:
                assert yd[k].grad is None
            else:
                assert torch.equal(yd[k].grad.cpu(), want)
    # default weights: drawn like the trainer (softmax of randn), convex
    out = wmattack.AttackMix(clamp=False, quantize=False)([torch.ones(2, 3, 4, 4, device=DEV)] * 3)
    assert md(out, torch.ones(2, 3, 4, 4)) <= 1e-6


@pytest.mark.parametrize("mode", (0, 1, 3))
def test_diffjpeg_saved_state_and_recompute_backward_agree(mode):
    """Two backward implementations of the same chain: from 7 B/px of state saved by the forward
    (default) and recomputed from x (`recompute_backward = True`, saves nothing)."""
    for shape, q, seed in (((2, 3, 64, 96), 50, 61), ((1, 3, 32, 272), 90, 62), ((3, 3, 48, 48), 20, 63)):
        x, g = rnd(shape, seed), rnd(shape, seed + 1)
        if seed == 62:
            x = torch.round(x * 3) / 3 * 1.2 - 0.1      # saturated, flat regions: clamp ties (code 2) occur
            x = x.clamp(0, 1)
        qs = torch.tensor([q, 35.0, 75.0][: shape[0]])
        outs = []
        for recompute in (False, True):
            m = wmattack.DiffJPEG(True, shape[2], shape[3], quality=q, rounding=mode)
            m.recompute_backward = recompute
            y, gx = fwd_bwd(lambda t: m(t, quality=qs.to(DEV)), x, g)
            outs.append((y, gx))
        assert torch.equal(outs[0][0], outs[1][0])
        assert md(outs[0][1], outs[1][1]) <= 2e-6


def test_train_step_example_runs_and_gradients_reach_the_encoder():
    """BASELINE config 4 in miniature (examples/train_step.py): encoder -> splice -> attack bank ->
    localiser; gradients reach the encoder through the straight-through bank and the DiffJPEG branch."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "train_step_example", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "train_step.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    torch.manual_seed(0); np.random.seed(0)
    model = ex.Step(64, 64).to(DEV)
    frames = torch.rand(4, 3, 64, 64, device=DEV)
    mask = (torch.rand(4, 1, 64, 64, device=DEV) > 0.85).float()
    loss, l_loc, l_img = model(frames, frames.roll(1, 0), mask)
    loss.backward()
    assert torch.isfinite(loss)
    g_enc = [p.grad for p in model.encoder.parameters()]
    g_loc = [p.grad for p in model.localiser.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in g_enc + g_loc)
    assert sum(float(g.abs().sum()) for g in g_enc) > 0


def test_from_uint8_matches_torch():
    for shape in ((2, 3, 33, 47), (1, 3, 64, 64), (5,)):
        u = torch.randint(0, 256, shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(7))
        got = WF.from_uint8(u.to(DEV)).cpu()
        assert torch.equal(got, u.float() / 255)


def test_degenerate_shapes():
    """Empty batches and tiny frames go through every layer (the reference's edge cases are torch's)."""
    for shape in ((0, 3, 32, 32), (1, 3, 1, 1), (1, 3, 2, 4), (2, 3, 7, 9), (1, 3, 8, 8)):
        x = rnd(shape, 71) if shape[0] else torch.zeros(shape)
        layers = [wmattack.Jpeg(50), wmattack.JpegSS(50), wmattack.JpegMask(50), wmattack.JpegCompression(DEV),
                  wmattack.GaussianBlur(), wmattack.MiddleBlur(3), wmattack.MiddleBlur(5), wmattack.Gaussian(),
                  wmattack.SaltPepper(0.1), wmattack.Identity(), wmattack.Quantization()]
        refs = {0: lambda t: O.jpeg8(t, 50, O.JPEG8_HARD), 2: lambda t: O.jpeg8(t, 50, O.JPEG8_MASK), 3: O.jpeg_compression,
                4: lambda t: O.gaussian_blur(t, 3), 5: lambda t: O.median_blur(t, 3), 6: lambda t: O.median_blur(t, 5)}
        for i, layer in enumerate(layers):
            xx = x.to(DEV).requires_grad_(True)
            y = layer(xx)
            assert y.shape == x.shape
            if shape[0]:
                y.sum().backward()
                assert torch.isfinite(y).all() and torch.isfinite(xx.grad).all()
                if i in refs:
                    assert md(y, refs[i](x.double())) <= 1e-5, (shape, i)
        if shape[0] and shape[2] >= 2 and shape[3] >= 2:
            xx = x.to(DEV).requires_grad_(True)
            y = wmattack.Resize()(xx, resize_ratio=0.6)
            assert md(y, O.resize(x.double(), 0.6)) <= 1e-5
            y.sum().backward()
        if shape[0] == 0:
            assert wmattack.DiffJPEG(True, 32, 32, 50)(x.to(DEV)).shape == x.shape
            assert wmattack.Resize()(x.to(DEV), resize_ratio=0.6).shape == x.shape


def test_fused_store_epilogue_bank_is_bit_identical_to_the_unfused_bank():
    """The attack kernels applying clamp + straight-through + Quantization in their own stores
    (the wm_store_epilogue argument of the forward entry points) must give exactly the values of: plain kernel, then the stand-alone
    epilogue kernel — for every layer that has a fused path, on aligned and ragged shapes."""
    import random
    for shape, seed in (((2, 3, 64, 96), 81), ((1, 3, 48, 132), 82), ((2, 3, 40, 56), 83)):
        x = rnd(shape, seed)
        h, w = shape[2:]
        def layers():
            ls = [wmattack.Resize(), wmattack.Resize(interpolation_method="bilinear"), wmattack.JpegMask(50), wmattack.Jpeg(70),
                  wmattack.JpegSS(30), wmattack.JpegCompression(DEV), wmattack.MiddleBlur(3), wmattack.MiddleBlur(5),
                  wmattack.GaussianBlur(), wmattack.GaussianBlur(7), wmattack.Gaussian(), wmattack.Identity(),
                  wmattack.Combined([wmattack.JpegMask(70), wmattack.Jpeg(70), wmattack.MiddleBlur(3)]), wmattack.SaltPepper(0.05)]
            if h % 16 == 0 and w % 16 == 0:
                ls.append(wmattack.DiffJPEG(True, h, w, quality=50))
            return ls
        outs = []
        for fused in (True, False):
            np.random.seed(5); random.seed(5); torch.manual_seed(5)
            WF._rng_calls = 0                                   # same Philox sub-streams in both runs
            bank = wmattack.AttackBank(layers())
            bank.fused = fused
            xx = x.to(DEV).requires_grad_(True)
            y = bank(xx)
            g = rnd(tuple(y.shape), seed + 100)
            y.backward(g.to(DEV))
            outs.append((y.detach().cpu(), xx.grad.cpu(), list(bank.names)))
        assert outs[0][2] == outs[1][2]
        assert torch.equal(outs[0][0], outs[1][0]), shape
        assert torch.equal(outs[0][1], outs[1][1])
        k = len(outs[0][2])
        assert md(outs[0][1], g.view(k, *shape).sum(0)) <= 1e-5


def test_gaussian_noise_mask_pair_equals_regenerating_pair():
    """The clamped Gaussian layer's training pair (1-bit pass mask saved by the forward, backward = masked
    copy of gy) must equal the pair that saves x and regenerates the noise — values, gradients, ragged
    lengths, injected noise, and values exactly on the clamp bounds (torch.clamp's mask is inclusive)."""
    for shape, seed in (((2, 3, 64, 64), 1), ((1, 3, 37, 53), 2), ((1, 1, 5, 7), 3), ((3, 3, 128, 128), 4)):
        x = rnd(shape, seed) * 1.4 - 0.2
        g = rnd(shape, seed + 10)
        for noise in (None, (rnd(shape, seed + 20) - 0.5) * 0.4):
            outs = []
            for regen in (False, True):
                WF._rng_calls = 0
                torch.manual_seed(3)
                xx = x.to(DEV).requires_grad_(True)
                y = WF.gaussian_noise(xx, 0.0, 0.05, True, None if noise is None else noise.to(DEV), regen=regen)
                y.backward(g.to(DEV))
                outs.append((y.detach().cpu(), xx.grad.cpu()))
            assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
            if noise is not None:      # against torch with the same noise
                xr = x.clone().requires_grad_(True)
                yr = torch.clamp(xr + noise, 0, 1)
                yr.backward(g)
                assert torch.equal(outs[0][0], yr.detach()) and torch.equal(outs[0][1], xr.grad)
    # exactly on the bounds: x + noise == 0 or 1 passes the gradient
    x = torch.tensor([0.0, 1.0, 0.5, -0.25, 1.25, 0.25] * 32).view(1, 3, 8, 8)
    nz = torch.tensor([0.0, 0.0, 0.5, 0.25, -0.25, -0.25] * 32).view(1, 3, 8, 8)
    xx = x.to(DEV).requires_grad_(True)
    WF.gaussian_noise(xx, 0.0, 0.05, True, nz.to(DEV)).backward(torch.ones(1, 3, 8, 8, device=DEV))
    assert torch.equal(xx.grad.cpu(), torch.ones(1, 3, 8, 8))


def test_fused_store_epilogue_without_clamp_and_out_of_range_values():
    """Quantization in the store epilogue uses a fast exact path (1.5*2^23 rounding + Newton-corrected
    reciprocal) guarded by a range check; un-clamped values far outside [0,1] (|v*255| >= 65536, inf, NaN)
    must take the IEEE path and still match torch's round(v*255)/255 bit for bit."""
    x = (rnd((2, 3, 64, 128), 91) - 0.5) * 4.0
    x[0, 0, 3, 5:9] = torch.tensor([300.0, -1000.0, 70000.0, 1e20])
    x[1, 2, 10, 0:2] = torch.tensor([float("inf"), float("nan")])
    xd = x.to(DEV)
    for layer in (wmattack.Identity(), wmattack.MiddleBlur(3), wmattack.GaussianBlur()):
        out = torch.empty_like(xd)
        with torch.no_grad():
            layer.forward_into(xd, out, (xd, False, True))
            sim = layer(xd)
            ref = torch.round((xd + (sim - xd)) * 255.0).cpu() / 255.0       # IEEE division on the host
        o = out.cpu()
        same = (o == ref) | (torch.isnan(o) & torch.isnan(ref))
        assert bool(same.all()), type(layer).__name__


# ================================================================ real codec (JpegTest)
import io as _io  # noqa: E402
import os as _os  # noqa: E402

from oracle import libjpeg_oracle as LJ  # noqa: E402

_LJ_GOLD = np.load(_os.path.join(_os.path.dirname(__file__), "golden", "libjpeg_golden.npz"))


def _pillow_roundtrip(rgb, q, s):
    Image = pytest.importorskip("PIL.Image")
    buf = _io.BytesIO()
    Image.fromarray(rgb).save(buf, format="JPEG", quality=q, subsampling=s)
    return np.array(Image.open(_io.BytesIO(buf.getvalue())), dtype=np.uint8)


def _codec_u8(rgb_hwc, q, s, **kw):
    x = torch.from_numpy(np.ascontiguousarray(rgb_hwc.transpose(2, 0, 1)))[None].to(DEV)
    return WF.jpeg_codec(x, q, s, "uint8", **kw)


@pytest.mark.parametrize("key", sorted(k for k in _LJ_GOLD.files if k.startswith("out/")))
def test_codec_matches_pillow_golden(key):
    """Bit-exact against the committed Pillow/libjpeg-turbo outputs (ragged sizes, all samplings)."""
    _, name, s, q = key.split("/")
    y = _codec_u8(_LJ_GOLD[f"in/{name}"], int(q[1:]), int(s[1:]))
    assert np.array_equal(y[0].permute(1, 2, 0).cpu().numpy(), _LJ_GOLD[key])


@pytest.mark.parametrize("hw", [(512, 512), (1080, 1920), (250, 301), (8, 8), (1, 1), (3, 700)])
@pytest.mark.parametrize("s", [0, 1, 2])
def test_codec_matches_pillow_live(hw, s):
    """Frame sizes of the BASELINE configs and ragged / tiny ones against Pillow on this machine."""
    rng = np.random.RandomState(hw[0] + s)
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    smooth = (np.sin(xx / 31.0)[..., None] * 0.3 + np.cos(yy / 17.0)[..., None] * 0.2 + 0.5) * 255
    rgb = np.clip(smooth + rng.randint(-40, 40, (*hw, 3)), 0, 255).astype(np.uint8)
    for q in (30, 90):
        ref = _pillow_roundtrip(rgb, q, s)
        got = _codec_u8(rgb, q, s)[0].permute(1, 2, 0).cpu().numpy()
        assert np.array_equal(got, ref), (hw, s, q, int((got != ref).sum()))


@pytest.mark.parametrize("s", [0, 1, 2])
def test_codec_quantised_coefficients_bit_exact(s):
    """The integers the entropy coder would see == the oracle's (north star: quantised DCT
    coefficients bit-exact)."""
    rng = np.random.RandomState(5 + s)
    rgb = rng.randint(0, 256, (72, 104, 3)).astype(np.uint8)
    q = 60
    _, (yq, cbq, crq) = _codec_u8(rgb, q, s, return_coefficients=True)
    hs, vs = ((1, 1), (2, 1), (2, 2))[s]
    ql, qc = LJ.quant_tables(q)
    y, cb, cr = LJ.rgb_to_ycc(rgb)
    mw, mh = -(-104 // (8 * hs)), -(-72 // (8 * vs))
    assert np.array_equal(yq[0].cpu().numpy(), LJ.quantised_plane(LJ.downsample(y, 1, 1, mh * vs, mw * hs), ql))
    assert np.array_equal(cbq[0].cpu().numpy(), LJ.quantised_plane(LJ.downsample(cb, hs, vs, mh, mw), qc))
    assert np.array_equal(crq[0].cpu().numpy(), LJ.quantised_plane(LJ.downsample(cr, hs, vs, mh, mw), qc))


def test_jpegtest_module_follows_reference_steps():
    """JpegTest.forward == the reference's steps (noise_layers/jpeg.py:21-45) with Pillow as the
    codec: same fp32 conversions on both sides of the byte round trip, batch of ragged frames."""
    pytest.importorskip("PIL.Image")
    x = (rnd((3, 3, 45, 70), 11) * 2.4 - 1.2)                       # exercises the clamp
    m = wmattack.JpegTest(50, subsample=2)
    y = m(x.to(DEV))
    assert y.shape == x.shape and y.dtype == torch.float32 and not y.requires_grad
    ref = torch.zeros_like(x)
    for i in range(x.shape[0]):
        single = ((x[i].clamp(-1, 1).permute(1, 2, 0) + 1) / 2 * 255).to(torch.uint8).numpy()
        dec = _pillow_roundtrip(single, 50, 2)
        t = torch.from_numpy(dec).permute(2, 0, 1).float().div(255)
        ref[i] = (t - 0.5) / 0.5
    assert torch.equal(y.cpu(), ref)
    # and the oracle's statement of the same thing
    assert np.array_equal(y.cpu().numpy(), LJ.jpegtest_forward(x.numpy(), 50, 2))


def test_codec_value_ranges_and_strides():
    rgb = np.random.RandomState(3).randint(0, 256, (2, 3, 40, 64)).astype(np.uint8)
    xu = torch.from_numpy(rgb).to(DEV)
    yu = WF.jpeg_codec(xu, 75, 2, "uint8")
    yf = WF.jpeg_codec(xu.float() / 255, 75, 2, "unit")
    assert torch.equal(yf.cpu(), yu.cpu().float() / 255)          # IEEE division, as ToTensor's on the host
    big = torch.zeros((2, 3, 48, 80), device=DEV)
    big[:, :, 4:44, 8:72] = xu.float() / 255
    assert torch.equal(WF.jpeg_codec(big[:, :, 4:44, 8:72], 75, 2, "unit"), yf)      # strided view, no copy
    with pytest.raises(ValueError):
        WF.jpeg_codec(xu, 0, 2, "uint8")
    with pytest.raises(ValueError):
        WF.jpeg_codec(xu, 50, 3, "uint8")
    with pytest.raises(TypeError):
        WF.jpeg_codec(xu.float(), 50, 2, "uint8")
    with pytest.raises(RuntimeError):
        WF.jpeg_codec(xu.cpu(), 50, 2, "uint8")
    assert WF.jpeg_codec(xu[:0], 50, 2, "uint8").shape == (0, 3, 40, 64)


def test_codec_frames_and_mcus_independent_at_4k():
    """Size-independent property at config 5's frame size: a 4K frame decodes to the same bytes
    alone or inside a batch, and luma of an MCU row strip does not depend on the rest."""
    g = torch.Generator().manual_seed(9)
    x = torch.randint(0, 256, (2, 3, 2160, 3840), generator=g, dtype=torch.uint8).to(DEV)
    y = WF.jpeg_codec(x, 85, 2, "uint8")
    assert torch.equal(WF.jpeg_codec(x[1:], 85, 2, "uint8"), y[1:])
    y444 = WF.jpeg_codec(x[:1], 85, 0, "uint8")
    assert torch.equal(WF.jpeg_codec(x[:1, :, 1024:1152], 85, 0, "uint8"), y444[:, :, 1024:1152])
    ref = _pillow_roundtrip(x[0].permute(1, 2, 0).cpu().numpy(), 85, 2)
    assert np.array_equal(y[0].permute(1, 2, 0).cpu().numpy(), ref)


# ================================================================ CUDA-graph capture of the layer calls
@pytest.mark.parametrize("name", ["diffjpeg", "jpegcompression", "blur", "median3", "resize", "jpegss"])
def test_cuda_graph_capture_replays_forward_and_backward(name):
    """Every launch goes to torch's current stream and the library never allocates or synchronises,
    so a layer's forward + backward can be captured once and replayed on new data (the trainers'
    per-frame loops at 256x256 are launch-bound: ~100 us eager vs ~10-20 us per replay)."""
    b, h, w = 2, 64, 128
    rs = wmattack.Resize()
    layer = {"diffjpeg": wmattack.DiffJPEG(True, h, w, quality=50), "jpegcompression": wmattack.JpegCompression(DEV),
             "blur": wmattack.GaussianBlur(), "median3": wmattack.MiddleBlur(3),
             "resize": lambda t: rs(t, resize_ratio=0.75), "jpegss": wmattack.JpegSS(50)}[name]
    xs = rnd((b, 3, h, w), 1).to(DEV).requires_grad_(True)
    gs = rnd((b, 3, h, w), 2).to(DEV)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            torch.autograd.grad(layer(xs), xs, gs)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_cap = layer(xs)
        (gx_cap,) = torch.autograd.grad(y_cap, xs, gs)
    for seed in (3, 5):
        x2, g2 = rnd((b, 3, h, w), seed).to(DEV), rnd((b, 3, h, w), seed + 1).to(DEV)
        with torch.no_grad():
            xs.copy_(x2)
            gs.copy_(g2)
        graph.replay()
        xe = x2.clone().requires_grad_(True)
        ye = layer(xe)
        (ge,) = torch.autograd.grad(ye, xe, g2)
        assert torch.equal(ye, y_cap) and torch.equal(ge, gx_cap)


# ================================================================ half-precision inputs / autocast
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_half_precision_inputs_cast_at_the_boundary(dt):
    """The trainers sometimes call the layers inside torch.cuda.amp.autocast()
    (models/IRNcrop_model.py:340), so a layer can be handed fp16/bf16 activations: they are cast to
    fp32 at the boundary, the result is the fp32 result for the same values, and autograd hands the
    gradient back in the input's dtype."""
    b, h, w = 2, 64, 64
    rs = wmattack.Resize()
    layers = [wmattack.DiffJPEG(True, h, w, quality=50), wmattack.JpegCompression(DEV), wmattack.JpegSS(50),
              wmattack.GaussianBlur(), wmattack.MiddleBlur(3), wmattack.MiddleBlur(5),
              lambda t: rs(t, resize_ratio=0.75), lambda t: wmattack.Crop()(t, apex=(8, 56, 4, 60))[0],
              wmattack.Identity()]
    xh = rnd((b, 3, h, w), 1).to(DEV).to(dt)
    g = rnd((b, 3, h, w), 2).to(DEV)
    for layer in layers:
        x1 = xh.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=dt):
            y1 = layer(x1)
        x2 = xh.float().requires_grad_(True)
        y2 = layer(x2)
        if y1 is x1:                       # Identity returns its argument
            continue
        assert y1.dtype == torch.float32 and torch.equal(y1, y2)
        y1.backward(g)
        y2.backward(g)
        assert x1.grad.dtype == dt and torch.equal(x1.grad, x2.grad.to(dt))


def test_make_graphed_callables_wraps_a_layer():
    """torch.cuda.make_graphed_callables (separate forward / backward graphs behind an autograd
    node) works on the layers as they are — the way to take the ~0.1 ms of eager Python + autograd
    per call out of the trainers' per-frame loops (models/IRNcrop_model.py:357-370)."""
    b, h, w = 1, 64, 64
    layer = wmattack.DiffJPEG(True, h, w, quality=50)
    sample = rnd((b, 3, h, w), 1).to(DEV).requires_grad_(True)
    graphed = torch.cuda.make_graphed_callables(layer, (sample,))
    for seed in (2, 3):
        x = rnd((b, 3, h, w), seed).to(DEV)
        g = rnd((b, 3, h, w), seed + 10).to(DEV)
        x1 = x.clone().requires_grad_(True)
        y1 = graphed(x1)
        y1.backward(g)
        x2 = x.clone().requires_grad_(True)
        y2 = layer(x2)
        y2.backward(g)
        assert torch.equal(y1, y2) and torch.equal(x1.grad, x2.grad)


def test_device_rng_makes_stochastic_layers_graph_capturable():
    """With the Philox state on the device (functional.device_rng) a captured Gaussian / SaltPepper call
    draws FRESH numbers at every replay, and its captured backward regenerates exactly the forward's
    numbers (gradient mask consistent with the output); the host counter is untouched afterwards."""
    try:
        WF.device_rng(True, seed=1234)
        for make, check in ((lambda: wmattack.Gaussian(), "gauss"), (lambda: wmattack.SaltPepper(0.3), "sp")):
            layer = make()
            xs = (rnd((2, 3, 32, 64), 1) * 0.6 + 0.2).to(DEV).requires_grad_(True)
            gs = torch.ones(2, 3, 32, 64, device=DEV)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    torch.autograd.grad(layer(xs), xs, gs)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                y = layer(xs)
                (gx,) = torch.autograd.grad(y, xs, gs)
            outs = []
            for _ in range(3):
                graph.replay()
                torch.cuda.synchronize()
                outs.append((y.clone(), gx.clone()))
            assert not torch.equal(outs[0][0], outs[1][0]) and not torch.equal(outs[1][0], outs[2][0])
            for yy, gg in outs:
                if check == "gauss":            # gradient 1 exactly where the clamp passed
                    noise = yy - xs.detach()
                    assert 0.03 < float(noise.std()) < 0.07
                    inside = (yy > 0) & (yy < 1)
                    assert bool((gg[inside] == 1).all())
                else:                           # salt & pepper: gradient 1 where the pixel was left alone
                    untouched = yy == xs.detach()
                    assert bool((gg[untouched] == 1).all()) and bool((gg[~untouched] == 0).all())
                    assert 0.2 < float((~untouched).float().mean()) < 0.4
    finally:
        WF.device_rng(False)
    a = wmattack.Gaussian()(torch.zeros(1, 3, 8, 8, device=DEV) + 0.5)
    assert torch.isfinite(a).all()


def test_quantization_bit_exact_on_boundary_values():
    """Quantization (models/modules/Quantization.py:7-14; the clamp is commented out upstream): round(x*255)/255
    bit-for-bit against torch on the host for every 8-bit level, every half-way point +- 1 ulp
    (round-half-even), random values outside [0,1], the clamped variant, and values far outside the fast
    quantiser's range (up to 1e6, inf, nan)."""
    k = torch.arange(256, dtype=torch.float32)
    levels = k / 255
    halves = (k[:-1] + 0.5) / 255
    pts = torch.cat([levels, halves, torch.nextafter(halves, torch.tensor(2.0)), torch.nextafter(halves, torch.tensor(-1.0)),
                     torch.nextafter(levels, torch.tensor(2.0)), torch.nextafter(levels, torch.tensor(-1.0)),
                     rnd((4096,), 3) * 3 - 1])
    pts = pts[: pts.numel() // 4 * 4].view(1, 1, -1, 4)
    assert torch.equal(wmattack.Quantization()(pts.to(DEV)).cpu(), torch.round(pts * 255.0) / 255.0)
    assert torch.equal(wmattack.Quantization(clamp01=True)(pts.to(DEV)).cpu(),
                       torch.round(torch.clamp(pts, 0, 1) * 255.0) / 255.0)
    wild = torch.cat([pts.flatten(), torch.tensor([300.0, -2e3, 7e4, 1e6, float("inf"), float("nan"), -0.0, 1e-40])])
    wild = wild[: wild.numel() // 4 * 4].view(1, 1, -1, 4).to(DEV)
    out = torch.empty_like(wild)
    with torch.no_grad():
        wmattack.Identity().forward_into(wild, out, (wild, False, True))          # un-clamped quantiser path
    wc = wild.cpu()
    ref = torch.round((wc + (wc - wc)) * 255.0) / 255.0                           # straight-through form: inf -> nan
    o = out.cpu()
    assert bool(((o == ref) | (torch.isnan(o) & torch.isnan(ref))).all())


# ================================================================ round 2: gaps named by the round-1 review
def _cropped_out_influence(x, box, eps=1e-5):
    """Influence region (on the input gradient) of the two clamps of Crop.cropped_out (crop.py:96-107):
    only `scaled_images` = clamp(bicubic_up(crop)) carries gradient (the scaled-back term is detached)."""
    h0, h1, w0, w1 = box
    hh, ww = x.shape[2:]
    x64 = x.double()
    pre = O.interpolate(x64[:, :, h0:h1, w0:w1], (hh, ww), "bicubic")
    fr = P.clamp_fragile(pre, eps)
    a_h, a_w = O.interp_matrix(h1 - h0, hh, "bicubic"), O.interp_matrix(w1 - w0, ww, "bicubic")
    infl = torch.zeros(x.shape, dtype=torch.bool)
    infl[:, :, h0:h1, w0:w1] = P.interp_influence(fr, a_h, a_w)
    return infl, int(fr.sum())


def test_cropped_out_values_and_gradients_vs_reference():
    """Crop.cropped_out (noise_layers/crop.py:78-118): all five outputs and the gradient of BOTH differentiable
    outputs (scaled_images through bicubic + clamp, zero_images through the paste; the dual-reshape term is
    detached upstream) against the reference's own autograd — seeded rectangle and caller-supplied fractional apex."""
    x, g, gz = T("x32"), T("g32"), T("cropped_out/seed23/gz")
    for key, kw in (("seed23", dict(min_rate=0.5)), ("apex", dict(apex=(0.125, 0.75, 0.25, 0.9375), min_rate=0.5))):
        np.random.seed(23)
        xx = x.to(DEV).requires_grad_(True)
        outs = wmattack.Crop().cropped_out(xx, **kw)
        ((outs[0] * g.to(DEV)).sum() + (outs[1] * gz.to(DEV)).sum()).backward()
        assert md(outs[0], T(f"cropped_out/{key}/x32/scaled")) <= 1e-5
        assert md(outs[1], T(f"cropped_out/{key}/x32/zero_images")) <= 1e-5
        if key == "seed23":
            assert np.allclose(np.array(outs[3]), GOLD["cropped_out/seed23/x32/apex"])
        box = tuple(int(round(v * 32)) for v in outs[3])
        infl, n_fragile = _cropped_out_influence(x, box)
        assert n_fragile <= 8
        P.assert_explained(xx.grad, T(f"cropped_out/{key}/x32/gx"), 1e-5, infl, f"cropped_out[{key}] gx")
        assert float(xx.grad.abs().max()) > 0


def test_cropped_for_outpainting_vs_reference():
    """Crop.cropped_for_outpainting (crop.py:57-76): pure slicing with two seeded rectangles — bit-exact,
    and the same two rectangles drawn from the same RNG calls."""
    x, real_h = T("x32"), T("outpainting/real_H")
    np.random.seed(25)
    a, b, c = wmattack.Crop().cropped_for_outpainting(x.to(DEV), real_h.to(DEV))
    assert torch.equal(a.cpu(), T("outpainting/seed25/x32/new_images"))
    assert torch.equal(b.cpu(), T("outpainting/seed25/x32/zero_images"))
    assert torch.equal(c.cpu(), T("outpainting/seed25/x32/GT"))
    xx = x.to(DEV).requires_grad_(True)
    np.random.seed(25)
    a, b, c = wmattack.Crop().cropped_for_outpainting(xx, real_h.to(DEV))
    (a.sum() + 2 * b.sum()).backward()                       # slices: gradient is an indicator sum
    ref = x.clone().requires_grad_(True)
    np.random.seed(25)
    ra = GOLD["outpainting/seed25/x32/new_images"].shape
    assert tuple(a.shape) == tuple(ra) and float(xx.grad.max()) in (1.0, 2.0, 3.0)


@pytest.mark.parametrize("rn", list(ROUND))
def test_codec_modules_with_call_time_size(rn):
    """utils/compression.py:147 compress_jpeg + utils/decompression.py:140-190 decompress_jpeg whose forward takes
    (y, cb, cr, height, width) at CALL time — non-square 48x32 frame, quality 40, all three rounding functions."""
    x = T("x4832")
    f = wmattack.quality_to_factor(40)
    comp = wmattack.compress_jpeg(rounding=ROUND[rn], factor=f)
    dec = wmattack.decompress_jpeg(rounding=ROUND[rn], factor=f)            # no size at construction
    cy, ccb, ccr = comp(x.to(DEV))
    refs = [T(f"codec_calltime/q40/{rn}/x4832/coef_{k}") for k in ("y", "cb", "cr")]
    for got, ref in zip((cy, ccb, ccr), refs):
        assert tuple(got.shape) == tuple(ref.shape)
        if rn == "hard":
            assert torch.equal(got.cpu(), ref)                              # integers: bit-exact (no tie in this fixture)
        else:
            assert md(got, ref) <= 2e-4                                     # coefficients are O(100): 1e-6 relative
    y = dec(cy, ccb, ccr, 48, 32)
    assert md(y, T(f"codec_calltime/q40/{rn}/x4832/y")) <= 1e-5
    # the reference's own coefficients through our decompress: isolates the decoder half
    y2 = dec(*[r.to(DEV) for r in refs], 48, 32)
    assert md(y2, T(f"codec_calltime/q40/{rn}/x4832/y")) <= 1e-5
    with pytest.raises(ValueError):
        dec(cy, ccb, ccr, 32, 48 + 16)                                       # size that does not match the coefficients


def test_rounding_ties_are_counted_not_excused():
    """A constructed worst case for the fragile-set logic: flat gray 129/255 puts EVERY luminance DC quotient at
    8*(129-128)/16 = 0.5 (+- fp32 noise of the colour matrix), a torch.round tie.  The fp64 oracle flags exactly those
    coefficients; our integer coefficients may differ from the oracle's by one step there and nowhere else."""
    x = torch.full((1, 3, 32, 32), 129.0 / 255.0)
    x[:, :, 16:, :] = rnd((1, 3, 16, 32), 5)                               # lower half: ordinary content
    m = wmattack.DiffJPEG(True, 32, 32, quality=50, rounding=2)
    cy, ccb, ccr = m.compress(x.to(DEV))
    qy, qcb, qcr = O.diffjpeg_compress(x.double(), 1.0, O.ROUND_NONE)
    ty, tc = O.diffjpeg_tables(1.0)
    fy = P.fragile_quotients(qy, ty, O.ROUND_HARD)
    assert int(fy.sum()) == 8 and bool(fy[0, :8, 0, 0].all())              # the 8 DC terms of the flat half, only those
    dy = (cy.cpu().double() - torch.round(qy)).abs()
    assert float(dy[~fy].max()) == 0.0 and float(dy[fy].max()) <= 1.0
    for got, q in ((ccb, qcb), (ccr, qcr)):
        fr = P.fragile_quotients(q, tc, O.ROUND_HARD)
        d = (got.cpu().double() - torch.round(q)).abs()
        assert float(d[~fr].max()) == 0.0 and float(d.max()) <= 1.0
    # and the pixel-level influence region covers exactly the flat half's luminance blocks
    infl, n = P.diffjpeg_influence(x, 1.0, O.ROUND_HARD)
    assert bool(infl[:, :, :16].all()) and n >= 8
    y = m(x.to(DEV))
    P.assert_explained(y, O.diffjpeg(x.double(), 50, O.ROUND_HARD), 1e-5, infl, "tie image")


def test_crop_rejects_or_clamps_out_of_range_rectangles():
    """A caller-supplied apex is clamped like the reference's slicing (image[:, :, a:b, c:d]); a window handed to
    the functional layer directly must lie inside the source (the kernels read it in place)."""
    x = rnd((1, 3, 32, 48), 3).to(DEV)
    y, apex = wmattack.Crop()(x, apex=(8, 40, 10, 60))                      # h_end, w_end beyond the frame
    ref = torch.nn.functional.interpolate(x[:, :, 8:40, 10:60], size=[32, 48], mode="bilinear")
    assert md(y, ref) <= 2e-6 and apex == (8, 40, 10, 60)
    with pytest.raises(ValueError):
        wmattack.Crop()(x, apex=(20, 10, 0, 48))                            # empty rectangle
    for bad in ((-1, 0, 8, 8), (0, 0, 33, 8), (30, 40, 8, 9), (0, 0, 0, 8)):
        with pytest.raises(ValueError):
            WF.interpolate(x, (32, 48), "bilinear", window=bad)


def test_nan_propagates_through_the_clamps_like_torch():
    """torch.clamp / torch.min / torch.max propagate NaN; a saturating instruction returns 0.  A diverging encoder
    must stay visible downstream: DiffJPEG's final clamp (per 16x16 MCU), quality = 100 (factor 0, NaN
    upstream: utils/JPEG.py:229 divides by table * 0), the clamped Gaussian layer, Resize's clamp, the epilogue."""
    x = rnd((1, 3, 32, 48), 1)
    x[0, 1, 5, 20] = float("nan")
    y = wmattack.DiffJPEG(True, 32, 48, quality=50)(x.to(DEV)).cpu()
    assert bool(torch.isnan(y[:, :, 0:16, 16:32]).all())                     # the whole MCU, all channels
    rest = y.clone(); rest[:, :, 0:16, 16:32] = 0
    assert bool(torch.isfinite(rest).all())
    y100 = wmattack.DiffJPEG(True, 32, 48, quality=100)(rnd((1, 3, 32, 48), 2).to(DEV))
    assert bool(torch.isnan(y100).all())                                     # as upstream: 0 * inf
    xg = x.to(DEV).requires_grad_(True)
    yg = wmattack.Gaussian()(xg)
    assert bool(torch.isnan(yg[0, 1, 5, 20])) and int(torch.isnan(yg).sum()) == 1
    assert bool(torch.isnan(wmattack.Gaussian()(x.to(DEV))[0, 1, 5, 20]))     # no-grad kernel too
    yr = wmattack.Resize()(x.to(DEV), resize_ratio=0.75)
    ref = torch.clamp(torch.nn.functional.interpolate(torch.nn.functional.interpolate(x, size=[24, 36], mode="bicubic"),
                                                      size=[32, 48], mode="bicubic"), 0, 1)
    # the fused kernel applies the composite band U*D with zero-padded windows (0 * NaN = NaN), so its NaN footprint
    # covers torch's two-pass footprint and may be a few columns / rows wider, never narrower
    ours_nan, ref_nan = torch.isnan(yr).cpu(), torch.isnan(ref)
    assert bool((ours_nan | ~ref_nan).all()) and int(ours_nan.sum()) <= 4 * int(ref_nan.sum())
    assert bool(torch.isfinite(yr.cpu()[~ours_nan]).all())
    out = wmattack.AttackEpilogue()(rnd((1, 3, 32, 48), 3).to(DEV), x.to(DEV))
    assert bool(torch.isnan(out[0, 1, 5, 20])) and int(torch.isnan(out).sum()) == 1
    q = wmattack.Quantization(clamp01=True)(x.to(DEV))
    assert bool(torch.isnan(q[0, 1, 5, 20]))


def test_jpegtest_consumes_the_reference_rng_stream():
    """JpegTest draws one 16-character temp-file name per frame from python's `random` upstream (noise_layers/jpeg.py:
    17-18, 32); ours consumes the same numbers, so seeded Combined choices after a JpegTest call stay aligned."""
    import random
    import string
    x = (rnd((3, 3, 16, 16), 4) * 2 - 1).to(DEV)
    random.seed(99)
    wmattack.JpegTest(50)(x)
    after = random.random()
    random.seed(99)
    for _ in range(3):
        random.sample(string.ascii_letters + string.digits, 16)
    assert after == random.random()


def test_bank3_shared_read_is_bit_identical_and_reads_x_once():
    """the 3x3-neighbourhood members of a bank (GaussianBlur(3), MiddleBlur(3), Gaussian, Identity) come
    from ONE kernel that stages each tile of x once (wm_bank3_fwd).  Every slice must equal, bit for bit, the member's own
    kernel with the store epilogue (shared_read=False) — ragged tiles, partial membership, duplicates, no clamp / no
    quantisation — and the launch count must drop accordingly."""
    import random
    from wmattack import _lib
    def members():
        return [wmattack.GaussianBlur(), wmattack.Identity(), wmattack.JpegMask(50), wmattack.MiddleBlur(3), wmattack.Gaussian(),
                wmattack.GaussianBlur(), wmattack.MiddleBlur(5)]          # second GaussianBlur(3): own kernel
    for shape, seed, flags in (((2, 3, 64, 128), 1, (True, True)), ((1, 3, 70, 132), 2, (True, True)), ((3, 3, 33, 36), 3, (False, True)),
                               ((1, 3, 130, 260), 4, (True, False)), ((2, 1, 17, 8), 5, (False, False))):
        x = rnd(shape, seed) * 1.3 - 0.15
        outs, launches = [], []
        for shared in (True, False):
            np.random.seed(5); random.seed(5); torch.manual_seed(5)
            WF._rng_calls = 0
            layers = members() if shape[1] == 3 else [wmattack.GaussianBlur(channels=1), wmattack.MiddleBlur(3), wmattack.Gaussian()]
            bank = wmattack.AttackBank(layers, clamp=flags[0], quantize=flags[1])
            bank.shared_read = shared
            xx = x.to(DEV).requires_grad_(True)
            n0 = _lib.launch_count
            y = bank(xx)
            launches.append(_lib.launch_count - n0)
            y.backward(torch.ones_like(y))
            outs.append((y.detach().cpu(), xx.grad.cpu(), list(bank.names)))
        assert outs[0][2] == outs[1][2]
        assert torch.equal(outs[0][0], outs[1][0]), shape
        assert torch.equal(outs[0][1], outs[1][1])
        n_shared = 4 if shape[1] == 3 else 3
        assert launches[1] - launches[0] >= n_shared - 1                    # K' member kernels (or more: a member whose
        # slice is not 32-byte aligned needs kernel + epilogue kernel on its own) became one
    # two-member minimum: a lone member keeps its own kernel
    bank = wmattack.AttackBank([wmattack.MiddleBlur(3), wmattack.JpegMask(50)])
    n0 = _lib.launch_count
    bank(rnd((1, 3, 32, 32), 9).to(DEV))
    assert _lib.launch_count - n0 == 2


def test_fused_resize_band_proof_covers_every_ratio_the_layer_can_draw():
    """Resize draws an arbitrary ratio in [0.5, 1.5] per call; the mid size int(r * n) takes n + 1 values.  Every one
    of them, at the trainers' 256 px and the BASELINE's 512 px frames (square and 2:1), must PASS the band proof that
    wm_resize_tables runs (rb_prove_kernel: the fused kernel's own window arithmetic over all tile positions, both
    directions) — none may need the two-call path, and the proof must not flag a geometry the kernel handles."""
    WF._RESIZE_UNFUSED.clear()
    for h, w in ((256, 256), (512, 512), (256, 512)):
        for mode in (WF.BICUBIC, WF.BILINEAR):
            for mh in range(h // 2, h + h // 2 + 1, 1 if h == 256 else 3):
                mw = int(mh * w / h)
                assert WF._lib.load().wm_resize_is_fused(h, w, mh, mw, 3) == 1
                assert WF._resize_tables(torch.device(DEV), h, w, (mh, mw), mode) is not None, (h, w, mh, mw, mode)
    assert not WF._RESIZE_UNFUSED
    # spot-check that proven geometries really are right at the extremes of the window sizes
    x, g = rnd((1, 3, 256, 256), 3), rnd((1, 3, 256, 256), 4)
    for r in (0.5, 0.501, 0.666, 0.999, 1.0, 1.499, 1.5):
        xx = x.to(DEV).requires_grad_(True)
        y = wmattack.Resize()(xx, resize_ratio=r)
        y.backward(g.to(DEV))
        xt = x.to(DEV).requires_grad_(True)
        mid = [int(r * 256)] * 2
        Fi = torch.nn.functional.interpolate
        yt = Fi(Fi(xt, size=mid, mode="bicubic"), size=[256, 256], mode="bicubic").clamp(0, 1)
        yt.backward(g.to(DEV))
        assert md(y, yt) <= 1e-5, r
        _assert_resize_grad(xx.grad, xt.grad, x, r, "bicubic", 2e-5, f"r={r}")


def test_random_shape_sweep_against_the_oracle():
    """tools/fuzz_shapes.py: 40 random [B,3,H,W] shapes (ragged widths, 1-pixel planes, odd plane counts, strided views)
    through blur, both medians, the 8x8 JPEG layers and the fused Resize, forward and gradient, against the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_shapes.py"), "40", "5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


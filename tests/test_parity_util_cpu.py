"""CPU checks of tests/parity_util.py (the fragile-set / influence-region logic the GPU parity suite relies on),
with the oracle's own fp32 evaluation standing in for "an fp32 implementation"."""
import numpy as np
import torch

from oracle import attack_oracle as O
from tests import parity_util as P
from tests.golden_util import GOLD, T


def rnd(shape, seed):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed))


def test_constructed_rounding_ties_are_flagged_and_explain_the_fp32_flips():
    x = torch.full((1, 3, 32, 32), 129.0 / 255.0)
    x[:, :, 16:, :] = rnd((1, 3, 16, 32), 5)
    qy, qcb, qcr = O.diffjpeg_compress(x.double(), 1.0, O.ROUND_NONE)
    ty, tc = O.diffjpeg_tables(1.0)
    fy = P.fragile_quotients(qy, ty, O.ROUND_HARD)
    assert int(fy.sum()) == 8 and bool(fy[0, :8, 0, 0].all())
    c32 = O.diffjpeg_compress(x.float(), 1.0, O.ROUND_HARD)
    dy = (c32[0].double() - torch.round(qy)).abs()
    assert float(dy[~fy].max()) == 0.0 and float(dy[fy].max()) == 1.0       # fp32 really flips there, only there
    infl, n = P.diffjpeg_influence(x, 1.0, O.ROUND_HARD)
    assert n == 8 and bool(infl[:, :, :16].all()) and not bool(infl[:, :, 16:].any())
    y32, y64 = O.diffjpeg(x.float(), 50, O.ROUND_HARD), O.diffjpeg(x.double(), 50, O.ROUND_HARD)
    assert P.assert_explained(y32, y64, 1e-5, infl) > 0                    # mismatches exist and are all explained
    try:
        P.assert_explained(y32, y64, 1e-5, torch.zeros_like(infl))
    except AssertionError:
        pass
    else:
        raise AssertionError("an unexplained mismatch must fail")


def test_fixtures_hold_no_fragile_positions_so_goldens_compare_strictly():
    for xn in ("x32", "xs32"):
        for mode in (O.ROUND_ONLY_AT_0, O.ROUND_CUBIC, O.ROUND_HARD):
            assert P.diffjpeg_influence(T(xn), 1.0, mode)[1] == 0
    assert P.jpeg8_influence(T("xs32"), 50, O.ROUND_HARD)[1] == 0
    infl, n, _ = P.resize_grad_influence(T("xsat"), O.resize_mid_size(32, 32, 0.8), "bicubic")
    assert n == 0
    xx = T("xsat").double().requires_grad_(True)
    O.resize(xx, 0.8).backward(T("g32").double())
    assert P.assert_explained(T("resize/bicubic/r0.8/xsat/gx"), xx.grad, 1e-5, infl) == 0


def test_clamp_influence_is_the_support_of_the_transposed_operator():
    x = rnd((1, 1, 12, 16), 1)
    mid = (9, 12)
    a_h, a_w = P.resize_operators(12, 16, mid, "bicubic")
    fr = torch.zeros(1, 1, 12, 16, dtype=torch.bool)
    fr[0, 0, 5, 7] = True
    infl = P.interp_influence(fr, a_h, a_w)
    # brute force: which inputs does output (5, 7) read?
    xx = x.double().requires_grad_(True)
    y = O.interpolate(O.interpolate(xx, mid, "bicubic"), (12, 16), "bicubic")
    y[0, 0, 5, 7].backward()
    assert torch.equal(infl, xx.grad != 0)


def test_fp32_noise_floor_is_small_at_the_baseline_quality():
    x, g = rnd((2, 3, 64, 64), 3), rnd((2, 3, 64, 64), 4)
    dy, dg = P.fp32_noise(lambda t: O.diffjpeg(t, 50), x, g)
    infl, _ = P.diffjpeg_influence(x, 1.0, O.ROUND_ONLY_AT_0)
    assert float(dy.max()) <= 1e-6 and float(dg[~infl].max()) <= 5e-6
    assert np.isfinite(float(dg.max()))
    assert "cropped_out/seed23/x32/gx" in GOLD and "codec_calltime/q40/hard/x4832/y" in GOLD

"""Counted, explained mismatches for the layers whose arithmetic has DISCONTINUITIES.

A rounding quantiser (torch.round, diff_round's inner torch.round, round_only_at_0's |q| < 1/2 switch) and a
clamp's pass-through mask are step functions of a pre-round / pre-clamp value.  When that value sits within
fp32 rounding noise of the step, two correct fp32 implementations (the reference's own CPU run, its CUDA run,
ours) may legitimately land on different sides.  Instead of allowing a FRACTION of wrong elements, the tests

  1. evaluate the pre-round / pre-clamp values with the fp64 oracle,
  2. build the FRAGILE set: positions within eps of a step, eps = a few fp32 ulps of the operand,
  3. map it to its INFLUENCE region on the compared tensor (the 8x8 block / 16x16 MCU of a coefficient; the
     support of the transposed interpolation operator for a clamp mask),
  4. assert that every element outside the influence region meets the max-abs bound, i.e. every mismatch is
     explained by a counted fragile position; the counts are returned so the tests can bound them.
"""
import numpy as np
import torch

from oracle import attack_oracle as O

# |C| <= 8*128 for the DC term; the level-shifted input is O(128): a handful of fp32 ulps of max(|C|, 128)
ULPS = 16.0


def _ulp32(v: torch.Tensor) -> torch.Tensor:
    a = v.abs().to(torch.float32).numpy()
    return torch.from_numpy(np.spacing(a).astype(np.float64))


def fragile_quotients(q: torch.Tensor, table: torch.Tensor, mode: int) -> torch.Tensor:
    """q = C / table (fp64, pre-round), any shape broadcastable with table.  True where q is within
    eps = ULPS * ulp32(max(|C|, 128)) / table of a step of the rounding surrogate `mode`."""
    c = (q * table).abs().clamp(min=128.0)
    eps = ULPS * _ulp32(c) / table
    if mode in (O.ROUND_HARD, O.ROUND_CUBIC):           # torch.round flips at half-integers
        d = (q - torch.floor(q) - 0.5).abs()
    elif mode == O.ROUND_ONLY_AT_0:                     # (|q| < 1/2) switch: value and derivative jump
        d = (q.abs() - 0.5).abs()
    else:                                               # Fourier surrogate: smooth
        return torch.zeros_like(q, dtype=torch.bool)
    return d <= eps


def diffjpeg_influence(x: torch.Tensor, factor, mode: int):
    """[B,3,H,W] bool: pixels whose DiffJPEG output / input-gradient depends on a fragile coefficient
    (luminance coefficient -> its 8x8 block, chroma coefficient -> its 16x16 MCU), and the fragile count."""
    x64 = x.double()
    b, _, h, w = x64.shape
    qy, qcb, qcr = O.diffjpeg_compress(x64, factor, O.ROUND_NONE)
    ty, tc = O.diffjpeg_tables(factor if not torch.is_tensor(factor) else 1.0)
    if torch.is_tensor(factor):
        f = factor.double().view(b, 1, 1, 1)
        ty, tc = ty * f, tc * f
    fy = fragile_quotients(qy, ty, mode)
    fc = fragile_quotients(qcb, tc, mode) | fragile_quotients(qcr, tc, mode)
    n = int(fy.sum()) + int(fragile_quotients(qcb, tc, mode).sum()) + int(fragile_quotients(qcr, tc, mode).sum())
    my = fy.flatten(2).any(2).view(b, h // 8, w // 8).repeat_interleave(8, 1).repeat_interleave(8, 2)
    mc = fc.flatten(2).any(2).view(b, h // 16, w // 16).repeat_interleave(16, 1).repeat_interleave(16, 2)
    return (my | mc).unsqueeze(1).expand(b, 3, h, w), n


def jpeg8_influence(x: torch.Tensor, q: float, mode: int, subsample: int = 0):
    """Same for the 4:4:4 8x8 layers (Jpeg / JpegSS): any fragile coefficient of any channel taints its
    8x8 block in all three RGB channels (the inverse colour matrix mixes them)."""
    x64 = x.double()
    b, _, h, w = x64.shape
    pq = O.jpeg8_prequant(x64, q, subsample)                    # [B,3,Hp,Wp]
    ly, lc = O.jpeg8_tables(O.jpeg8_scale(q))
    hp, wp = pq.shape[2:]
    tab = torch.stack([ly, lc, lc]).double().repeat(1, hp // 8, wp // 8)
    fr = fragile_quotients(pq, tab, mode)
    blk = fr.view(b, 3, hp // 8, 8, wp // 8, 8).permute(0, 2, 4, 1, 3, 5).flatten(3).any(3)     # [B, hp/8, wp/8]
    m = blk.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :h, :w]
    return m.unsqueeze(1).expand(b, 3, h, w), int(fr.sum())


def clamp_fragile(pre: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Positions whose pre-clamp value (fp64 oracle) is within eps of a bound of clamp(., 0, 1)."""
    return (pre.abs() <= eps) | ((pre - 1).abs() <= eps)


def interp_influence(fragile: torch.Tensor, a_h: torch.Tensor, a_w: torch.Tensor) -> torch.Tensor:
    """fragile: [..., Ho, Wo] bool on the OUTPUT of y = A_h x A_w^T; returns [..., Hi, Wi] bool: the input
    positions whose gradient receives a contribution from a fragile output (support of the transpose)."""
    sh = (a_h != 0).double().t()          # [Hi, Ho]
    sw = (a_w != 0).double()              # [Wo, Wi]
    return (sh @ fragile.double() @ sw) > 0


def resize_operators(h, w, mid, mode):
    """Per-axis dense operators of the Resize round trip (up o down) in fp64."""
    a_h = O.interp_matrix(mid[0], h, mode) @ O.interp_matrix(h, mid[0], mode)
    a_w = O.interp_matrix(mid[1], w, mode) @ O.interp_matrix(w, mid[1], mode)
    return a_h, a_w


def resize_grad_influence(x: torch.Tensor, mid, mode: str, eps: float = 1e-5):
    """(influence on gx [B,C,H,W], fragile count, fp64 pre-clamp values) of Resize's clamp mask."""
    x64 = x.double()
    h, w = x64.shape[2:]
    pre = O.interpolate(O.interpolate(x64, mid, mode), (h, w), mode)
    fr = clamp_fragile(pre, eps)
    a_h, a_w = resize_operators(h, w, mid, mode)
    return interp_influence(fr, a_h, a_w), int(fr.sum()), pre


def assert_explained(got, ref, tol, influence, what=""):
    """Every element outside `influence` is within tol; returns (#mismatches, all of them inside)."""
    err = (got.detach().double().cpu() - ref.detach().double().cpu()).abs()
    bad = err > tol
    outside = bad & ~influence
    assert not bool(outside.any()), (f"{what}: {int(outside.sum())} mismatches > {tol:g} NOT explained by a fragile "
                                     f"position (max {float(err[~influence].max()):.3e})")
    return int(bad.sum())


def fp32_noise(fn, x, g):
    """max |fp32 - fp64| of the oracle's own gradient: how far a CORRECT fp32 evaluation of the same
    formulas (torch CPU) is from fp64 on this input — the floor any fp32 implementation sits on."""
    outs = []
    for dt in (torch.float32, torch.float64):
        xx = x.to(dt).clone().requires_grad_(True)
        y = fn(xx)
        y.backward(g.to(dt))
        outs.append((y.detach().double(), xx.grad.double()))
    return (outs[0][0] - outs[1][0]).abs(), (outs[0][1] - outs[1][1]).abs()

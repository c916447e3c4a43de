"""Generates csrc/median_pair_net.cuh: the two pruned min/max networks of the 5x5 median kernel that
produces TWO vertically adjacent outputs per step.

Outputs y and y+1 share four of their five window rows.  With every row already sorted (9 comparators,
shared by five outputs), the 20 shared values are reduced ONCE to the six of rank 7..12 (network A):
a value of rank <= 6 among the 20 has rank <= 11 among the 25 and a value of rank >= 13 has rank >= 13,
so neither can be the median (rank 12) of either window.  Each output then is the median (rank 5) of
those six and its own sorted row (network B, 11 inputs).  Networks M and C split A in two levels: the
merged row pair (r+3, r+4) of one step is the pair (r+1, r+2) of the next, so each step merges ONE new
pair (M) and takes the middle six of two merged pairs (C).

Method: Batcher's odd-even merge sort restricted to the real wires,
comparators that never fire under the sortedness precondition dropped, greedy deletion with random
restarts while the required outputs stay correct, correctness checked EXHAUSTIVELY with the 0-1
principle over all monotone 0/1 assignments (6^4 = 1296 for A, 7*6 = 42 for B); comparators with one
dead output become a single min or max.

    python tools/gen_median_pair_networks.py [--restarts 60] [--seed 0]
"""
import argparse
import itertools
import os
import random

import numpy as np


def batcher(n_pow2):
    out = []
    t = n_pow2.bit_length() - 1
    p = 1 << (t - 1)
    while p > 0:
        q, r, d = 1 << (t - 1), 0, p
        while d > 0:
            for i in range(n_pow2 - d):
                if (i & p) == r:
                    out.append((i, i + d))
            d, q, r = q - p, q >> 1, p
        p >>= 1
    return out


def cases(groups):
    rows = []
    for ones in itertools.product(*[range(g + 1) for g in groups]):
        v = []
        for g, k in zip(groups, ones):
            v += [0] * (g - k) + [1] * k
        rows.append(v)
    return np.array(rows, dtype=bool).T


class Problem:
    def __init__(self, groups, outs):
        self.n = sum(groups)
        self.cases = cases(groups)
        self.truth = np.sort(self.cases, axis=0)[outs]
        self.outs = list(outs)

    def run(self, net, keep=None):
        w = self.cases.copy()
        for idx, (a, b) in enumerate(net):
            if keep is not None and not keep[idx]:
                continue
            lo, hi = w[a] & w[b], w[a] | w[b]
            w[a], w[b] = lo, hi
        return w

    def ok(self, net, keep=None):
        return np.array_equal(self.run(net, keep)[self.outs], self.truth)

    def never_fires(self, net):
        w = self.cases.copy()
        kept = []
        for a, b in net:
            lo, hi = w[a] & w[b], w[a] | w[b]
            if not (np.array_equal(lo, w[a]) and np.array_equal(hi, w[b])):
                kept.append((a, b))
            w[a], w[b] = lo, hi
        return kept

    def liveness(self, net):
        live = set(self.outs)
        flags = [None] * len(net)
        for idx in range(len(net) - 1, -1, -1):
            a, b = net[idx]
            la, lb = a in live, b in live
            flags[idx] = (la, lb)
            if la or lb:
                live.add(a)
                live.add(b)
        return flags

    def cost(self, net):
        return sum(int(x) + int(y) for x, y in self.liveness(net))

    def prune(self, net, rng):
        keep = [True] * len(net)
        order = list(range(len(net)))
        improved = True
        while improved:
            improved = False
            rng.shuffle(order)
            for i in order:
                if not keep[i]:
                    continue
                keep[i] = False
                if self.ok(net, keep):
                    improved = True
                else:
                    keep[i] = True
        net2 = [c for c, k in zip(net, keep) if k]
        fl = self.liveness(net2)
        return [c for c, f in zip(net2, fl) if f[0] or f[1]]

    def search(self, restarts, rng, label):
        pow2 = 1 << (self.n - 1).bit_length()
        base = [(a, b) for a, b in batcher(pow2) if b < self.n]
        assert np.array_equal(self.run(base), np.sort(self.cases, axis=0)), "base network does not sort"
        base = self.never_fires(base)
        best = None
        for r in range(restarts):
            net = self.prune(base, rng)
            c = self.cost(net)
            if best is None or c < best[0]:
                best = (c, net)
                print(f"{label} restart {r}: {len(net)} comparators, {c} min/max ops", flush=True)
        assert self.ok(best[1])
        return best[1]

    def body(self, net, var="v"):
        lines = []
        for (a, b), (la, lb) in zip(net, self.liveness(net)):
            if la and lb:
                lines.append(f"    ce({var}[{a}], {var}[{b}]);")
            elif la:
                lines.append(f"    {var}[{a}] = fminf({var}[{a}], {var}[{b}]);")
            elif lb:
                lines.append(f"    {var}[{b}] = fmaxf({var}[{a}], {var}[{b}]);")
        return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--restarts", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = random.Random(args.seed)
    pa = Problem([5, 5, 5, 5], range(7, 13))
    na = pa.search(args.restarts, rng, "A (mid six of four sorted rows)")
    pb = Problem([6, 5], [5])
    nb = pb.search(args.restarts, rng, "B (median of sorted 6 + sorted 5)")
    pm = Problem([5, 5], range(10))
    nm = pm.search(args.restarts, rng, "M (merge two sorted rows)")
    pc = Problem([10, 10], range(7, 13))
    nc = pc.search(args.restarts, rng, "C (mid six of two merged row pairs)")
    text = f"""// GENERATED by tools/gen_median_pair_networks.py — do not edit.
// 5x5 median, two vertically adjacent outputs per step (they share four sorted window rows):
//   mid6_of_4_sorted_rows : v[0..19] = four ascending rows of five -> v[7..12] = their values of rank 7..12,
//                           ascending ({len(na)} comparators, {pa.cost(na)} FMNMX), shared by both outputs;
//   median11_sorted_6_5   : v[0..5] ascending, v[6..10] ascending -> the median of the eleven
//                           ({len(nb)} comparators, {pb.cost(nb)} FMNMX), once per output.
//   merge10_sorted_5_5    : v[0..4], v[5..9] ascending -> v[0..9] ascending ({len(nm)} comparators, {pm.cost(nm)} FMNMX);
//   mid6_of_2_sorted_10   : v[0..9], v[10..19] ascending -> v[7..12] = rank 7..12, ascending ({len(nc)} comparators,
//                           {pc.cost(nc)} FMNMX).  merge + mid6 replace mid6_of_4_sorted_rows when the merged row pair
//                           (r+3, r+4) of one step is kept for the next step, where it is the pair (r+1, r+2).
// All verified exhaustively with the 0-1 principle over every monotone 0/1 assignment of the sorted groups.
#pragma once
namespace wm {{
// A compare-exchange is a functor (a, b) -> (min, max): CeMinMax is the plain FMNMX pair (2 ALU-pipe ops);
// the kernels may pass one that computes the max on the FMA pipe instead (median.cu, CeIntSum).
struct CeMinMax {{
    __device__ __forceinline__ void operator()(float& a, float& b) const {{ const float lo = fminf(a, b); b = fmaxf(a, b); a = lo; }}
}};
template <class CE = CeMinMax>
__device__ __forceinline__ void mid6_of_4_sorted_rows(float (&v)[20], const CE ce = CE()) {{
{pa.body(na)}
}}
__device__ __forceinline__ float median11_sorted_6_5(float (&v)[11]) {{
{pb.body(nb)}
    return v[5];
}}
template <class CE = CeMinMax>
__device__ __forceinline__ void merge10_sorted_5_5(float (&v)[10], const CE ce = CE()) {{
{pm.body(nm)}
}}
template <class CE = CeMinMax>
__device__ __forceinline__ void mid6_of_2_sorted_10(float (&v)[20], const CE ce = CE()) {{
{pc.body(nc)}
}}
}}  // namespace wm
"""
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(os.path.dirname(here), "video-watermarking-forgery-detection_b200", "csrc", "median_pair_net.cuh")
    with open(out, "w") as f:
        f.write(text)
    print(f"wrote {out}: A {len(na)}/{pa.cost(na)}, B {len(nb)}/{pb.cost(nb)}, M {len(nm)}/{pm.cost(nm)}, C {len(nc)}/{pc.cost(nc)} (comparators/ops)")


if __name__ == "__main__":
    main()

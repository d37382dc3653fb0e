#!/usr/bin/env python3
"""Emit subset_dp_sparse_gen.cuh: per-lane subset enumeration specialised on (K, F).

F = number of GPU *positions* that can be free in this warp.  The sparse kernel
(score_pairs_sparse.cuh) permutes every node's GPUs so that its free GPUs occupy positions
0..f-1 (f = popcount(free_mask), increasing GPU index), and orders the nodes so that all lanes
of a warp have f <= F with F as small as possible.  A k-subset can only be feasible if it lies
inside positions 0..F-1, so best_kK_fF enumerates the C(F,K) subsets of those positions instead
of all C(8,K) -- exactly the subsets the CPU twin visits (`if (S & ~free) continue`), at warp
granularity.  Lanes with f < F still see positions f..F-1: their pair costs carry PEN.

Same instruction-shape rules and the same per-pod multiplier as gen_subset_dp.py (every add that
can be is an IMAD by `k.one`, which the kernel derives from the pod's own request):
    K <= 4 : bottom-up with shared partial sums Q(ab,x), T(abd)
    K >= 5 : complement inside the F positions, comp = F-K in 0..3:
             key(S) = tm - max over comp-subsets D of [ sum_{i in D} (R_i + bit_i) - cost(D) ]
             with R_i = sum_{j<F} c_ij, tm = Total_F + (2^F - 1)

Run:  python gen_subset_dp_sparse.py > subset_dp_sparse_gen.cuh
"""
import itertools
import os

NACC = int(os.environ.get("KGPU_GEN_NACC", "3"))


def c(i, j):
    i, j = min(i, j), max(i, j)
    return "p.c%d%d" % (i, j)


def y(i, j):
    i, j = min(i, j), max(i, j)
    return "F2(p.c%d%d, 0x%02xu)" % (i, j, (1 << i) | (1 << j))


class Body:
    def __init__(self, op3="MIN3"):
        self.lines, self.pending, self.op3, self.turn = [], [], op3, 0
        init = "0xFFFFFFFFu" if op3 == "MIN3" else "0u"
        self.w("uint32_t " + ", ".join("b%d = %s" % (i, init) for i in range(NACC)) + ";")

    def w(self, s):
        self.lines.append("    " + s)

    def acc(self):
        a = "b%d" % (self.turn % NACC)
        self.turn += 1
        return a

    def fold(self, expr):
        self.pending.append(expr)
        if len(self.pending) == 2:
            a = self.acc()
            self.w("%s = %s(%s, %s, %s);" % (a, self.op3, a, self.pending[0], self.pending[1]))
            self.pending = []

    def addfold(self, x, yv):
        a = self.acc()
        self.w("%s = %s(%s, %s, %s);" % (a, "ADDMIN" if self.op3 == "MIN3" else "ADDMAX", x, yv, a))

    def flush(self):
        if self.pending:
            a = self.acc()
            self.w("%s = %s(%s, %s);" % (a, "min" if self.op3 == "MIN3" else "max", a, self.pending[0]))
            self.pending = []

    def finish(self):
        self.flush()
        accs = ["b%d" % i for i in range(NACC)]
        while len(accs) > 1:
            if len(accs) >= 3:
                self.w("%s = %s(%s, %s, %s);" % (accs[0], self.op3, accs[0], accs[1], accs[2]))
                accs = [accs[0]] + accs[3:]
            else:
                self.w("%s = %s(%s, %s);" % (accs[0], "min" if self.op3 == "MIN3" else "max", accs[0], accs[1]))
                accs = [accs[0]]
        self.w("const uint32_t best = b0;")


def dep_sum(terms):
    """Sum of >= 1 terms in which every instruction depends on k.one (F2 tree)."""
    level = list(terms)
    if len(level) == 1:
        return "F2(%s, 0u)" % level[0]
    while len(level) > 1:
        nxt = ["F2(%s, %s)" % (level[i], level[i + 1]) for i in range(0, len(level) - 1, 2)]
        if len(level) % 2:
            nxt.append(level[-1])
        level = nxt
    return level[0]


def gen_direct(K, F):
    b = Body()
    if K == 2:
        for a, bb in itertools.combinations(range(F), 2):
            b.fold(y(a, bb))
    elif K == 3:
        for a, bb, d in itertools.combinations(range(F), 3):
            b.fold("A3(%s, F2(%s, %s), 0x%02xu)" % (y(a, bb), c(a, d), c(bb, d), 1 << d))
    else:  # K == 4
        for a, bb in itertools.combinations(range(F), 2):
            if bb > F - 3:
                continue
            b.w("{")
            for x in range(bb + 1, F):
                b.w("    const uint32_t q%d = F2(%s, %s);" % (x, c(a, x), c(bb, x)))
            for d in range(bb + 1, F - 1):
                b.w("    const uint32_t t%d = F2(%s, q%d);" % (d, y(a, bb), d))
                for e in range(d + 1, F):
                    b.fold("A3(t%d, q%d, %s)" % (d, e, y(d, e)))
            b.flush()
            b.w("}")
    b.finish()
    b.w("return best;")
    return b.lines


def gen_complement(K, F):
    comp = F - K
    lines = []
    w = lambda s: lines.append("    " + s)
    if comp == 0:
        terms = [c(a, bb) for a, bb in itertools.combinations(range(F), 2)]
        w("return %s + 0x%02xu;" % (dep_sum(terms), (1 << F) - 1))
        return lines
    b = Body("MAX3")
    for i in range(F):
        b.w("const uint32_t r%d = %s;" % (i, dep_sum([c(i, j) for j in range(F) if j != i])))
    b.w("const uint32_t tm = (%s >> 1) + 0x%02xu;" % (dep_sum(["r%d" % i for i in range(F)]), (1 << F) - 1))
    for i in range(F):
        b.w("const uint32_t q%d = r%d + 0x%02xu;" % (i, i, 1 << i))
    if comp == 1:
        for a in range(F):
            b.fold("q%d" % a)
    elif comp == 2:
        for a, bb in itertools.combinations(range(F), 2):
            b.addfold("F2N(%s, q%d)" % (c(a, bb), bb), "q%d" % a)
    else:  # comp == 3
        for a, bb in itertools.combinations(range(F), 2):
            if bb > F - 2:
                continue
            b.w("{")
            b.w("    const uint32_t h = F2(F2N(%s, q%d), q%d);" % (c(a, bb), bb, a))
            for d in range(bb + 1, F):
                b.fold("A3(h, q%d, 0u - F2(%s, %s))" % (d, c(a, d), c(bb, d)))
            b.flush()
            b.w("}")
    b.finish()
    b.w("return tm - best;")
    return b.lines


def gen():
    out = []
    w = out.append
    w("// GENERATED by gen_subset_dp_sparse.py -- do not edit.  See that file for the scheme.")
    w("#pragma once")
    w('#include "subset_dp_gen.cuh"   // PairCosts, PipeConsts')
    w("namespace kgpu {")
    w("")
    w("#define A3(a, b, c) ((a) + (b) + (c))")
    w("#define F2(a, b) ((a) * k.one + (b))")
    w("#define F2N(a, b) ((a) * k.minus_one + (b))")
    w("#define MIN3(a, b, c) __vimin3_u32((a), (b), (c))")
    w("#define MAX3(a, b, c) __vimax3_u32((a), (b), (c))")
    w("#define ADDMIN(a, b, c) __viaddmin_u32((a), (b), (c))")
    w("#define ADDMAX(a, b, c) __viaddmax_u32((a), (b), (c))")
    w("")
    w("// min over the K-subsets S of positions 0..F-1 of (cost(S) << 8 | S); p holds COMPACTED pair costs")
    w("template <int K, int F> __device__ __forceinline__ uint32_t best_kf(const PairCosts &p, const PipeConsts k);")
    w("")
    for K in range(2, 9):
        for F in range(K, 9):
            w("template <> __device__ __forceinline__ uint32_t best_kf<%d, %d>(const PairCosts &p, const PipeConsts k) {" % (K, F))
            out.extend(gen_direct(K, F) if K <= 4 else gen_complement(K, F))
            w("}")
            w("")
    for m in ("A3", "F2", "F2N", "MIN3", "MAX3", "ADDMIN", "ADDMAX"):
        w("#undef " + m)
    w("")
    w("}  // namespace kgpu")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    import sys
    sys.stdout.write(gen())

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
    """Running min (or max) over the folded terms in NACC independent accumulators.  An accumulator is
    created from its first three terms (one 3-input min), so no instruction is spent on an identity
    element; `finish` merges the live accumulators."""

    def __init__(self, op3="MIN3"):
        self.lines, self.pending, self.op3 = [], [], op3
        self.op2 = "min" if op3 == "MIN3" else "max"
        self.live, self.turn, self.declared = [], 0, set()

    def w(self, s):
        self.lines.append("    " + s)

    def _assign(self, name, expr):
        self.declared.add(name)
        self.w("%s = %s;" % (name, expr))

    def fold(self, expr):
        self.pending.append(expr)
        if len(self.live) < NACC:                      # still creating accumulators: wait for three terms
            if len(self.pending) == 3:
                name = "b%d" % len(self.live)
                self._assign(name, "%s(%s, %s, %s)" % (self.op3, *self.pending))
                self.live.append(name)
                self.pending = []
        elif len(self.pending) == 2:
            a = self.live[self.turn % NACC]
            self.turn += 1
            self._assign(a, "%s(%s, %s, %s)" % (self.op3, a, *self.pending))
            self.pending = []

    def addfold(self, x, yv):
        """acc = min/max(acc, x + yv) in one instruction (VIADDMNMX)."""
        op = "ADDMIN" if self.op3 == "MIN3" else "ADDMAX"
        if len(self.live) < NACC:
            name = "b%d" % len(self.live)
            self._assign(name, "(%s) + (%s)" % (x, yv))
            self.live.append(name)
        else:
            a = self.live[self.turn % NACC]
            self.turn += 1
            self._assign(a, "%s(%s, %s, %s)" % (op, x, yv, a))

    def flush(self):
        """Scope boundary: pending terms refer to temporaries that are about to go out of scope."""
        if not self.pending:
            return
        if not self.live:
            expr = self.pending[0] if len(self.pending) == 1 else "%s(%s, %s)" % (self.op2, *self.pending)
            self._assign("b0", expr)
            self.live.append("b0")
        else:
            a = self.live[self.turn % len(self.live)]
            self.turn += 1
            if len(self.pending) == 1:
                self._assign(a, "%s(%s, %s)" % (self.op2, a, self.pending[0]))
            else:
                self._assign(a, "%s(%s, %s, %s)" % (self.op3, a, *self.pending))
        self.pending = []

    def finish(self):
        self.flush()
        accs = list(self.live)
        while len(accs) > 1:
            if len(accs) >= 3:
                self.w("%s = %s(%s, %s, %s);" % (accs[0], self.op3, accs[0], accs[1], accs[2]))
                accs = [accs[0]] + accs[3:]
            else:
                self.w("%s = %s(%s, %s);" % (accs[0], self.op2, accs[0], accs[1]))
                accs = [accs[0]]
        self.w("const uint32_t best = %s;" % accs[0])
        self.lines.insert(0, "    uint32_t %s;" % ", ".join(sorted(self.declared)))


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
            g = "%d%d" % (a, bb)                       # temporaries are named per (a, b) so folds can span groups
            for x in range(bb + 1, F):
                b.w("const uint32_t q%s_%d = F2(%s, %s);" % (g, x, c(a, x), c(bb, x)))
            for d in range(bb + 1, F - 1):
                b.w("const uint32_t t%s_%d = F2(%s, q%s_%d);" % (g, d, y(a, bb), g, d))
                for e in range(d + 1, F):
                    b.fold("A3(t%s_%d, q%s_%d, %s)" % (g, d, g, e, y(d, e)))
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
        b.w("const uint32_t q%d = F2(r%d, 0x%02xu);" % (i, i, 1 << i))      # an IMAD, not an ALU add
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
            b.w("const uint32_t h%d%d = F2(F2N(%s, q%d), q%d);" % (a, bb, c(a, bb), bb, a))
            for d in range(bb + 1, F):
                b.fold("A3(h%d%d, q%d, 0u - F2(%s, %s))" % (a, bb, d, c(a, d), c(bb, d)))
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

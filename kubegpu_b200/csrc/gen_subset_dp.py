#!/usr/bin/env python3
"""Emit subset_dp_gen.cuh: straight-line per-lane subset enumeration for K = 2..8.

Why generated
  * with nested `#pragma unroll` loops nvcc 12.9 leaves a residual rolled loop for part
    of the K=4 nest and indexes the pair-cost array dynamically (-> local memory);
  * the first straight-line version ran at 78% ALU-pipe / 4% FMA-pipe utilisation (ncu,
    profiles/r01_k1_v1_summary.md): every add and every min went down the ALU pipe.
    Each SM sub-partition can issue one warp instruction per clock but each of the two
    integer-capable pipes (ALU: IADD3/VIMNMX3/VIADDMNMX; FMA: IMAD) accepts one every
    two clocks, so this version assigns instructions to pipes explicitly:
        A3(a,b,c)   = a + b + c              IADD3      ALU pipe
        F2(a,b)     = a * one + b            IMAD       FMA pipe (one: opaque kernel arg == 1)
        F2N(a,b)    = a * minus_one + b      IMAD       FMA pipe (b - a)
        MIN3/MAX3   = __vimin3_u32/__vimax3_u32   VIMNMX3   ALU pipe
        ADDMIN/ADDMAX = __viaddmin_u32/__viaddmax_u32  VIADDMNMX  ALU pipe

Representation (see score_pairs.cuh): per node 28 scaled pair costs
    cXY = (W[level(X,Y)] << 8) + (PEN if X or Y is not free)
and yXY = cXY + (1<<X) + (1<<Y).  A subset's key is the sum of its pair costs plus its
bit mask == (cost << 8) | S; subsets touching a non-free GPU are >= PEN.

Every function enumerates ALL C(8,K) subsets and returns the minimum key.
K <= 4 builds keys bottom-up sharing partial sums:
    Q(ab,x)  = c_ax + c_bx                      (F2)
    T(abd)   = y_ab + Q(ab,d)                   (F2)   key of {a,b,d} minus bit d
    key{a,b,d,e} = T(abd) + Q(ab,e) + y_de      (A3)
K >= 5 uses the complement identity (exact in Z/2^32), with R_i = sum_j c_ij,
R'_i = R_i + (1<<i), tm = Total + 0xFF:
    key(S) = tm - [ sum_{i in ~S} R'_i - cost(~S) ]   -> minimise key == maximise [...]

Run:  python gen_subset_dp.py > subset_dp_gen.cuh
"""
import itertools
import os

PAIRS = list(itertools.combinations(range(8), 2))
NACC = int(os.environ.get("KGPU_GEN_NACC", "3"))          # independent min/max accumulators (ILP)
K4_FORMB_EVERY = int(os.environ.get("KGPU_GEN_K4_FORMB", "0"))   # every n-th K=4 subset as IMAD+VIADDMNMX (0 = never)
K3_FORMB_EVERY = int(os.environ.get("KGPU_GEN_K3_FORMB", "0"))   # same for K=3
STAGE_Y = os.environ.get("KGPU_GEN_STAGE_Y", "0") == "1"   # keep yXY in registers (makes K=2 hoistable: off)


def c(i, j):
    i, j = min(i, j), max(i, j)
    return "p.c%d%d" % (i, j)


def y(i, j):
    i, j = min(i, j), max(i, j)
    if STAGE_Y:
        return "p.y%d%d" % (i, j)
    return "F2(p.c%d%d, 0x%02xu)" % (i, j, (1 << i) | (1 << j))


def bits(*idx):
    m = 0
    for b in idx:
        m |= 1 << b
    return m


class Body:
    """Statement list with NACC round-robin accumulators b0..b{NACC-1} (independent
    dependency chains); `finish` combines them into `best`."""

    def __init__(self, op3="MIN3"):
        self.lines = []
        self.pending = []      # keys waiting for a MIN3/MAX3
        self.op3 = op3
        self.turn = 0
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
        """acc = min/max(x + y, acc) as one VIADDMNMX."""
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


def gen_row_sums(b):
    """r_i = sum_j c_ij (4 F2 + 1 A3, every operand of the A3 goes through an F2 so nothing
    here is invariant across pods), tm = total + 0xFF, q_i = r_i + bit_i."""
    for i in range(8):
        t = [c(i, j) for j in range(8) if j != i]
        b.w("const uint32_t r%d = A3(F2(%s, %s), F2(%s, %s), F2(%s, F2(%s, %s)));" % (i, t[0], t[1], t[2], t[3], t[4], t[5], t[6]))
    b.w("const uint32_t tm = (A3(A3(r0, r1, r2), A3(r3, r4, r5), F2(r6, r7)) >> 1) + 0xFFu;")
    for i in range(8):
        b.w("const uint32_t q%d = r%d + 0x%02xu;" % (i, i, 1 << i))   # R'_i


def gen():
    out = []
    w = out.append
    w("// GENERATED by gen_subset_dp.py -- do not edit.  See that file for the scheme.")
    w("#pragma once")
    w("#include <cstdint>")
    w("namespace kgpu {")
    w("")
    w("// Scaled pair costs of one node (cXY) and the same plus the two GPU bits (yXY).")
    w("struct PairCosts {")
    w("    uint32_t " + ", ".join("c%d%d" % p for p in PAIRS) + ";")
    if STAGE_Y:
        w("    uint32_t " + ", ".join("y%d%d" % p for p in PAIRS) + ";")
    w("};")
    w("")
    w("// Opaque multipliers (kernel arguments holding 1 and 0xFFFFFFFF) that pin F2/F2N to IMAD.")
    w("struct PipeConsts {")
    w("    uint32_t one, minus_one;")
    w("};")
    w("")
    w("#define A3(a, b, c) ((a) + (b) + (c))")
    w("#define F2(a, b) ((a) * k.one + (b))")
    w("#define F2N(a, b) ((a) * k.minus_one + (b))")
    w("#define MIN3(a, b, c) __vimin3_u32((a), (b), (c))")
    w("#define MAX3(a, b, c) __vimax3_u32((a), (b), (c))")
    w("#define ADDMIN(a, b, c) __viaddmin_u32((a), (b), (c))")
    w("#define ADDMAX(a, b, c) __viaddmax_u32((a), (b), (c))")
    w("")
    w("// static visitor used by staging: f(i, j) is called with literal constants")
    w("template <class F>")
    w("__device__ __forceinline__ void for_each_pair(PairCosts &p, F f) {")
    for a, b in PAIRS:
        if STAGE_Y:
            w("    p.c%d%d = f(%d, %d); p.y%d%d = p.c%d%d + 0x%02xu;" % (a, b, a, b, a, b, a, b, bits(a, b)))
        else:
            w("    p.c%d%d = f(%d, %d);" % (a, b, a, b))
    w("}")
    w("")

    # memory-aware path: per-pod penalties on GPUs whose memory is too small for the pod
    w("// out.cXY = in.cXY | pen[X] | pen[Y]  (pen[i] = PEN if GPU i is not eligible for this pod).  OR, not")
    w("// add: a pair must carry PEN at most once (28 * PEN stays below 2^31); the scaled cost is < PEN.")
    w("__device__ __forceinline__ void apply_pens(const PairCosts &p, PairCosts &o, const uint32_t (&pen)[8]) {")
    for a, b in PAIRS:
        w("    o.c%d%d = p.c%d%d | pen[%d] | pen[%d];" % (a, b, a, b, a, b))
    w("}")
    w("")

    # Compiler barrier: tells nvcc the pair costs may have changed, so the subset enumeration
    # below cannot be hoisted out of the per-pod loop (it IS loop invariant under snapshot
    # scoring: that shortcut is the separately reported tile-memo variant).  No instruction.
    w("__device__ __forceinline__ void per_pair_barrier(PairCosts &p) {")
    fields = ["c%d%d" % pr for pr in PAIRS] + (["y%d%d" % pr for pr in PAIRS] if STAGE_Y else [])
    for i in range(0, len(fields), 14):
        grp = fields[i:i + 14]
        w('    asm volatile("" : ' + ", ".join('"+r"(p.%s)' % f for f in grp) + ");")
    w("}")
    w("")

    w("// K = 1: the 8 single-GPU subsets all cost 0; the key is the lowest free bit (times the per-pod one).")
    w("__device__ __forceinline__ uint32_t best_k1(uint32_t free, const PipeConsts k) {")
    w("    return free ? (free & (0u - free)) * k.one : 0xFFFFFFFFu;")
    w("}")
    w("")

    sig = "__device__ __forceinline__ uint32_t best_k%d(const PairCosts &p, const PipeConsts k) {"

    # ---- K = 2: min over the 28 y values -------------------------------------------
    b = Body()
    for a, bb in PAIRS:
        b.fold(y(a, bb))
    b.finish()
    b.w("return best;")
    out += [sig % 2] + b.lines + ["}", ""]

    # ---- K = 3: key{a,b,d} = y_ab + Q(ab,d) + bit_d -----------------------------------
    b = Body()
    n = 0
    for a, bb, d in itertools.combinations(range(8), 3):
        q = "F2(%s, %s)" % (c(a, d), c(bb, d))
        if K3_FORMB_EVERY and n % K3_FORMB_EVERY == K3_FORMB_EVERY - 1:      # form B: both adds on the FMA pipe, fused add+min on the ALU pipe
            b.addfold("F2(%s, %s)" % (y(a, bb), q), "0x%02xu" % (1 << d))
        else:               # form A: one IADD3, MIN3 per two keys
            b.fold("A3(%s, %s, 0x%02xu)" % (y(a, bb), q, 1 << d))
        n += 1
    b.finish()
    b.w("return best;")
    out += [sig % 3] + b.lines + ["}", ""]

    # ---- K = 4 ----------------------------------------------------------------------------
    b = Body()
    n = 0
    for a, bb in PAIRS:
        if bb > 5:
            continue
        b.w("{")
        for x in range(bb + 1, 8):
            b.w("    const uint32_t q%d = F2(%s, %s);" % (x, c(a, x), c(bb, x)))
        for d in range(bb + 1, 7):
            b.w("    const uint32_t t%d = F2(%s, q%d);" % (d, y(a, bb), d))
            for e in range(d + 1, 8):
                if K4_FORMB_EVERY and n % K4_FORMB_EVERY == K4_FORMB_EVERY - 1:
                    b.addfold("F2(t%d, q%d)" % (d, e), y(d, e))
                else:
                    b.fold("A3(t%d, q%d, %s)" % (d, e, y(d, e)))
                n += 1
        b.flush()
        b.w("}")
    b.finish()
    b.w("return best;")
    out += [sig % 4] + b.lines + ["}", ""]

    # ---- K = 5: complement triples: maximise H_ab + R'_d - Q(ab,d) ---------------------
    b = Body("MAX3")
    gen_row_sums(b)
    for a, bb in PAIRS:
        if bb > 6:
            continue
        b.w("{")
        b.w("    const uint32_t h = F2(F2N(%s, q%d), q%d);" % (c(a, bb), bb, a))      # R'_a + R'_b - c_ab
        for d in range(bb + 1, 8):
            b.fold("A3(h, q%d, 0u - F2(%s, %s))" % (d, c(a, d), c(bb, d)))
        b.flush()
        b.w("}")
    b.finish()
    b.w("return tm - best;")
    out += [sig % 5] + b.lines + ["}", ""]

    # ---- K = 6: complement pairs: maximise R'_a + R'_b - c_ab --------------------------
    b = Body("MAX3")
    gen_row_sums(b)
    for a, bb in PAIRS:
        b.addfold("F2N(%s, q%d)" % (c(a, bb), bb), "q%d" % a)
    b.finish()
    b.w("return tm - best;")
    out += [sig % 6] + b.lines + ["}", ""]

    # ---- K = 7: complement singletons: maximise R'_a -----------------------------------
    b = Body("MAX3")
    gen_row_sums(b)
    for a in range(8):
        b.fold("q%d" % a)
    b.finish()
    b.w("return tm - best;")
    out += [sig % 7] + b.lines + ["}", ""]

    # ---- K = 8: the one subset --------------------------------------------------------------
    terms = [c(a, bb) for a, bb in PAIRS]
    b = Body()
    b.lines = []
    # a tree of F2 (x*one + y with x itself a product of `one`): the sum is a polynomial in the
    # per-pod multiplier, which no compiler factors back into hoistable invariant partial sums
    level = terms[:]
    while len(level) > 1:
        nxt = ["F2(%s, %s)" % (level[i], level[i + 1]) for i in range(0, len(level) - 1, 2)]
        if len(level) % 2:
            nxt.append(level[-1])
        level = nxt
    b.w("return %s + 0xFFu;" % level[0])
    out += [sig % 8] + b.lines + ["}", ""]

    w("#undef A3")
    w("#undef F2")
    w("#undef F2N")
    w("#undef MIN3")
    w("#undef MAX3")
    w("#undef ADDMIN")
    w("#undef ADDMAX")
    w("")
    w("}  // namespace kgpu")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    import sys
    sys.stdout.write(gen())

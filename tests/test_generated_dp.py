"""CPU check of the GENERATED enumeration code (subset_dp_gen.cuh, subset_dp_sparse_gen.cuh): the
straight-line C of every best_kK / best_kf<K,F> is transliterated to Python, evaluated in uint32
arithmetic on random pair costs (with PEN on some pairs) and compared with brute force over the subsets.
Guards the generators without a GPU; the kernels themselves are checked bit-for-bit on the B200."""
import itertools
import os
import random
import re

import pytest

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kubegpu_b200", "csrc")
M32 = 0xFFFFFFFF
PEN = 1 << 26
PAIRS = list(itertools.combinations(range(8), 2))


class U32(int):
    def __new__(cls, v):
        return int.__new__(cls, int(v) & M32)

    def __add__(self, o): return U32(int(self) + int(o))
    __radd__ = __add__
    def __sub__(self, o): return U32(int(self) - int(o))
    def __rsub__(self, o): return U32(int(o) - int(self))
    def __mul__(self, o): return U32(int(self) * int(o))
    __rmul__ = __mul__
    def __rshift__(self, o): return U32(int(self) >> int(o))
    def __or__(self, o): return U32(int(self) | int(o))


class P:
    pass


ENV = {
    "A3": lambda a, b, c: U32(a) + U32(b) + U32(c),
    "MIN3": lambda a, b, c: U32(min(int(a), int(b), int(c))),
    "MAX3": lambda a, b, c: U32(max(int(a), int(b), int(c))),
    "ADDMIN": lambda a, b, c: U32(min(int(U32(a) + U32(b)), int(c))),
    "ADDMAX": lambda a, b, c: U32(max(int(U32(a) + U32(b)), int(c))),
    "min": lambda a, b: U32(min(int(a), int(b))),
    "max": lambda a, b: U32(max(int(a), int(b))),
    "U32": U32,
}


def functions(path):
    """{name: python source} for every generated __device__ function returning uint32_t."""
    text = open(path).read()
    out = {}
    for m in re.finditer(r"uint32_t (best_k\w*(?:<\d, \d>)?)\(([^)]*)\) \{\n(.*?)\n\}\n", text, re.S):
        name, body = m.group(1), m.group(3)
        lines = []
        for ln in body.splitlines():
            ln = ln.strip()
            if ln in ("{", "}") or not ln or re.match(r"^uint32_t [\w, ]+;$", ln):   # braces, bare declarations
                continue
            ln = re.sub(r"^(const )?uint32_t ", "", ln).rstrip(";")
            ln = re.sub(r"\b(0x[0-9a-fA-F]+|\d+)u\b", r"U32(\1)", ln)
            ln = re.sub(r"\bF2N\(", "F2N(", ln)
            if ln.startswith("return "):
                ln = "result = " + ln[len("return "):]
            # "a = 1, b = 2, c = 3" declarations -> separate statements
            if re.match(r"^b0 = .*, b1 = ", ln):
                ln = "; ".join(part.strip() for part in ln.split(","))
            lines.append(ln)
        out[name] = "\n".join(lines)
    return out


def run(src, costs, one=1, free=0xFF):
    p = P()
    for (a, b) in PAIRS:
        setattr(p, "c%d%d" % (a, b), U32(costs[(a, b)]))
    env = dict(ENV)
    env["p"] = p
    env["free"] = U32(free)
    env["F2"] = lambda a, b: U32(a) * U32(one) + U32(b)
    env["F2N"] = lambda a, b: U32(a) * U32((-one) & M32) + U32(b)
    k = P()
    k.one, k.minus_one = U32(one), U32((-one) & M32)
    env["k"] = k
    exec(src, env)
    return int(env["result"])


def brute(costs, K, F=8):
    best = M32
    for comb in itertools.combinations(range(F), K):
        key = sum(costs[pr] for pr in itertools.combinations(comb, 2)) + sum(1 << i for i in comb)
        best = min(best, key & M32)
    return best


def random_costs(rng, bad_positions=(), nonzero=False):
    costs = {}
    for (a, b) in PAIRS:
        c = rng.choice([1, 2, 8, 32, 64, 4095] if nonzero else [0, 1, 2, 8, 32, 64, 4095]) << 8
        if a in bad_positions or b in bad_positions:
            c += PEN
        costs[(a, b)] = c
    return costs


def test_dense_generated_functions_match_brute_force():
    fns = functions(os.path.join(CSRC, "subset_dp_gen.cuh"))
    assert {"best_k%d" % k for k in range(2, 9)} <= set(fns)
    rng = random.Random(1)
    for trial in range(40):
        bad = rng.sample(range(8), rng.choice([0, 0, 1, 3, 5]))
        costs = random_costs(rng, bad)
        for K in range(2, 9):
            got = run(fns["best_k%d" % K], costs)
            want = brute(costs, K)
            # keys >= PEN are all "infeasible": the kernels only test key >= PEN, so compare the class
            assert (got == want) or (got >= PEN and want >= PEN), (K, trial)
    # the per-pod multiplier really is everywhere: with one = 2 the result must change for every K
    costs = random_costs(rng, nonzero=True)
    for K in range(2, 9):
        assert run(fns["best_k%d" % K], costs, one=2) != run(fns["best_k%d" % K], costs, one=1), K


@pytest.mark.parametrize("K", range(2, 9))
def test_sparse_generated_functions_match_brute_force(K):
    fns = functions(os.path.join(CSRC, "subset_dp_sparse_gen.cuh"))
    rng = random.Random(K)
    for F in range(K, 9):
        name = "best_kf<%d, %d>" % (K, F)
        assert name in fns, name
        for trial in range(12):
            f = rng.randint(0, F)                                  # this lane's own free count <= the warp's F
            costs = random_costs(rng, bad_positions=range(f, 8))   # compacted: positions >= f are not free
            got = run(fns[name], costs)
            want = brute(costs, K, F)
            assert (got == want) or (got >= PEN and want >= PEN), (K, F, trial)
            if f >= K:                                             # feasible lanes: exact key, cost and mask
                assert got == want < PEN and bin(got & 0xFF).count("1") == K and (got & 0xFF) < (1 << f)
        costs = random_costs(rng, nonzero=True)
        assert run(fns[name], costs, one=2) != run(fns[name], costs, one=1)


def test_generated_files_are_up_to_date():
    """The committed .cuh files are exactly what the generators emit (default parameters)."""
    import subprocess
    import sys
    for gen, out in (("gen_subset_dp.py", "subset_dp_gen.cuh"), ("gen_subset_dp_sparse.py", "subset_dp_sparse_gen.cuh")):
        env = {k: v for k, v in os.environ.items() if not k.startswith("KGPU_GEN_")}
        txt = subprocess.run([sys.executable, os.path.join(CSRC, gen)], capture_output=True, text=True, env=env, check=True).stdout
        assert txt == open(os.path.join(CSRC, out)).read(), out

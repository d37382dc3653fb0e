"""Oracle B: the C restatement checked against an independent pure-Python twin, its
tuned CPU variant, the committed vectors, and hand-computed known answers."""
import os

import numpy as np
import pytest

from kubegpu_b200 import synth


def test_known_answers(oracle_b):
    ob = oracle_b
    t1 = synth.shape_matrix([[2, 2], [2, 2]])          # pairs level 5 (W=2), socket 3 (W=8), cross 1 (W=32)
    assert ob.node_key(t1, 0xFF, 1) == (0 << 8) | 0x01
    assert ob.node_key(t1, 0xFF, 2) == (2 << 8) | 0x03
    assert ob.node_key(t1, 0xFF, 3) == ((2 + 8 + 8) << 8) | 0x07
    assert ob.node_key(t1, 0xFF, 4) == ((2 + 2 + 4 * 8) << 8) | 0x0F
    assert ob.node_key(t1, 0xFF, 8) == ((4 * 2 + 8 * 8 + 16 * 32) << 8) | 0xFF
    assert ob.node_key(t1, 0xFE, 2) == (2 << 8) | 0x0C      # GPU0 busy: next tight pair is {2,3}
    assert ob.node_key(t1, 0x55, 2) == (8 << 8) | 0x05      # only one GPU of each pair free
    assert ob.node_key(t1, 0x11, 2) == (32 << 8) | 0x11     # must cross sockets
    assert ob.node_key(t1, 0x01, 2) == ob.NODE_NO_FIT
    assert ob.node_key(t1, 0x00, 1) == ob.NODE_NO_FIT
    assert ob.node_key(t1, 0x00, 0) == 0
    assert ob.node_key(t1, 0xFF, 9) == ob.NODE_NO_FIT and ob.node_key(t1, 0xFF, -1) == ob.NODE_NO_FIT
    t2 = synth.shape_matrix([[4], [2, 2]])
    assert ob.node_key(t2, 0xFF, 3) == (6 << 8) | 0x07      # three GPUs of the 4-group: 3 pairs x W[5]=2
    # unknown level 0 is the most expensive (W[0]=64); NVLink levels cost 0
    k80 = np.zeros(64, dtype=np.int32)
    assert ob.node_key(k80, 0x0F, 2) == (64 << 8) | 0x03
    nvl = np.full(64, 9, dtype=np.int32)
    assert ob.node_key(nvl, 0xF0, 4) == (0 << 8) | 0xF0


def test_pod_key_order_and_no_fit(oracle_b):
    ob = oracle_b
    topo, free, pods = synth.gen_c1()
    keys = ob.score_batch(topo, free, pods, node_id_base=100)
    # k=1: cost 0 everywhere -> lowest node id (100), lowest free GPU
    assert ob.unpack_key(keys[0]) == (0, 100, 0x01)
    assert ob.unpack_key(keys[1]) == (2, 100, 0x03)
    assert ob.unpack_key(keys[2]) == (6, 101, 0x07)          # [[4],[2,2]] node wins for k=3
    assert ob.unpack_key(keys[3]) == (12, 101, 0x0F)
    none = ob.score_batch(topo, np.zeros_like(free), pods)
    assert (none == ob.NO_FIT).all()
    assert ob.score_batch(topo[:0], free[:0], pods).tolist() == [int(ob.NO_FIT)] * 4
    assert ob.score_batch(topo, free, pods[:0]).shape == (0,)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c_matches_python_twin_random(oracle_b, seed):
    ob = oracle_b
    rng = np.random.default_rng(seed)
    N, P = 40, 24
    topo = rng.integers(0, 16, size=(N, 64)).astype(np.int32)
    free = rng.integers(0, 256, size=N).astype(np.int32)
    pods = synth.make_pods(rng.integers(-1, 10, size=P).astype(np.int32))
    W = rng.integers(0, 4096, size=16).astype(np.int32)
    a = ob.score_batch(topo, free, pods, W, node_id_base=7)
    b = ob.score_batch_py(topo, free, pods, W, node_id_base=7)
    c = ob.score_batch(topo, free, pods, W, node_id_base=7, fast=True, nthreads=3)
    assert (a == b).all() and (a == c).all()


def test_fast_variant_matches_plain_on_configs(oracle_b):
    ob = oracle_b
    for gen, kw in ((synth.gen_c2, dict(N=5000, P=40)), (synth.gen_c3, dict(N=3000, P=64)),
                    (synth.gen_c4, dict(N=4000, P=64))):
        topo, free, pods = gen(**kw)
        assert (ob.score_batch(topo, free, pods) == ob.score_batch(topo, free, pods, fast=True, nthreads=4)).all()


def test_committed_vectors(oracle_b, golden_dir):
    """tests/golden/oracle_b_vectors.npz (made by make_golden.py) still reproduces."""
    z = np.load(os.path.join(golden_dir, "oracle_b_vectors.npz"))
    for tag in ("c1", "c2", "c3", "c4"):
        keys = oracle_b.score_batch(z[tag + "_topo"].astype(np.int32), z[tag + "_free"], z[tag + "_pods"])
        assert (keys == z[tag + "_keys"]).all(), tag


def test_sharding_is_transparent(oracle_b):
    """result independent of how the node list is split (G in 1,2,4,8)."""
    ob = oracle_b
    topo, free, pods = synth.gen_c4(N=1000, P=48)
    whole = ob.score_batch(topo, free, pods)
    for G in (2, 4, 8):
        per = (1000 + G - 1) // G
        parts = [ob.score_batch(topo[g * per:(g + 1) * per], free[g * per:(g + 1) * per], pods, node_id_base=g * per)
                 for g in range(G)]
        assert (ob.reduce_shards(np.stack(parts)) == whole).all()


def test_weights_monotone_property(oracle_b):
    """Scaling every weight by c scales every cost by c and keeps node/mask."""
    ob = oracle_b
    topo, free, pods = synth.gen_c4(N=500, P=32)
    k1 = ob.score_batch(topo, free, pods, ob.DEFAULT_WEIGHTS)
    k3 = ob.score_batch(topo, free, pods, ob.DEFAULT_WEIGHTS * 3)
    fit = k1 != ob.NO_FIT
    assert ((k3 == ob.NO_FIT) == ~fit).all()
    assert ((k3[fit] >> np.uint64(40)) == (k1[fit] >> np.uint64(40)) * np.uint64(3)).all()
    assert ((k3[fit] & np.uint64((1 << 40) - 1)) == (k1[fit] & np.uint64((1 << 40) - 1))).all()


def test_relabelling_gpus_permutes_the_mask_and_keeps_the_cost(oracle_b):
    """The definition does not depend on how a node's GPUs are numbered (what K1s' compaction relies on):
    permuting rows/columns of the matrix and the bits of the free mask permutes the chosen mask when
    the optimum is unique, and always keeps the optimal cost."""
    ob = oracle_b
    rng = np.random.default_rng(3)
    for _ in range(200):
        M = rng.integers(0, 13, size=(8, 8)).astype(np.int32)
        M = np.triu(M, 1)
        M = M + M.T
        free = int(rng.integers(0, 256))
        k = int(rng.integers(1, 9))
        perm = rng.permutation(8)                    # new position i holds old GPU perm[i]
        M2 = M[np.ix_(perm, perm)]
        free2 = sum(((free >> int(perm[i])) & 1) << i for i in range(8))
        a, b = ob.node_key(M.reshape(64), free, k), ob.node_key(M2.reshape(64), free2, k)
        assert (a == ob.NODE_NO_FIT) == (b == ob.NODE_NO_FIT)
        if a != ob.NODE_NO_FIT:
            assert a >> 8 == b >> 8
            back = sum(((b >> i) & 1) << int(perm[i]) for i in range(8))      # b's mask in the old numbering
            assert bin(back).count("1") == k and (back & ~free) == 0
            assert ob.lib().kgpu_oracle_subset_cost is not None


def test_freeing_gpus_never_hurts_and_monotone_in_k(oracle_b):
    ob = oracle_b
    rng = np.random.default_rng(4)
    for _ in range(200):
        M = rng.integers(0, 7, size=64).astype(np.int32)
        free = int(rng.integers(0, 256))
        more = free | int(rng.integers(0, 256))
        for k in range(1, 9):
            a, b = ob.node_key(M, free, k), ob.node_key(M, more, k)
            if a != ob.NODE_NO_FIT:
                assert b != ob.NODE_NO_FIT and (b >> 8) <= (a >> 8)
        costs = [ob.node_key(M, more, k) >> 8 for k in range(1, bin(more).count("1") + 1)]
        assert costs == sorted(costs)                # with non-negative weights a bigger set never costs less


def test_cpu_twins_for_the_bench_agree_with_the_plain_oracle(oracle_b):
    """The fair CPU baselines bench.py times next to the GPU (two-level-minima sequential placement, k-memoised
    snapshot scoring) give the plain oracle's bits."""
    for gen, n, p in ((synth.gen_c2, 5000, 700), (synth.gen_c4, 4133, 500), (synth.gen_c2, 100, 300)):
        topo, free, pods = gen(N=n, P=p)
        pods[3, 0], pods[4, 0] = 0, 9
        want, wf = oracle_b.place_batch(topo, free, pods, node_id_base=5)
        got, gf = oracle_b.place_batch(topo, free, pods, node_id_base=5, tiled=True)
        assert (got == want).all() and (gf == wf).all()
        for nt in (1, 3):
            assert (oracle_b.score_batch_memo(topo, free, pods, node_id_base=5, nthreads=nt)
                    == oracle_b.score_batch(topo, free, pods, node_id_base=5, fast=True, nthreads=2)).all()

"""GPU parity (run with `-m gpu` on the B200 box): every K1 variant, through the C ABI,
bit-exact against Oracle B on the same seeded inputs, against the committed golden
vectors, and -- at BASELINE's full sizes -- through size-independent properties."""
import os

import numpy as np
import pytest

from kubegpu_b200 import _lib, synth

pytestmark = pytest.mark.gpu

VARIANTS = [_lib.VARIANT_AUTO, _lib.VARIANT_LANE_PER_NODE, _lib.VARIANT_WARP_PER_PAIR, _lib.VARIANT_MEMO_BY_K, _lib.VARIANT_TILE_MEMO,
            _lib.VARIANT_SPARSE]


@pytest.fixture(scope="module")
def scorer():
    from kubegpu_b200.scorer import Scorer
    s = Scorer((0,))
    yield s
    s.close()


def _score(scorer, variant, topo, free, pods, base=0, W=None):
    scorer.set_variant(variant)
    if W is not None:
        scorer.set_weights(W)
    scorer.upload_nodes(topo, free, node_id_base=base)
    out = scorer.score_batch(pods)
    if W is not None:
        scorer.set_weights([64, 32, 16, 8, 4, 2, 1] + [0] * 9)
    return out


@pytest.mark.parametrize("variant", VARIANTS)
def test_committed_golden_vectors(scorer, golden_dir, variant):
    z = np.load(os.path.join(golden_dir, "oracle_b_vectors.npz"))
    for tag in ("c1", "c2", "c3", "c4"):
        keys = _score(scorer, variant, z[tag + "_topo"].astype(np.int32), z[tag + "_free"], z[tag + "_pods"])
        assert (keys == z[tag + "_keys"]).all(), tag


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("gen,kw", [
    (synth.gen_c2, dict(N=20_000, P=257)),
    (synth.gen_c3, dict(N=12_345, P=300)),
    (synth.gen_c4, dict(N=33_000, P=200)),
])
def test_parity_vs_oracle_seeded(scorer, oracle_b, variant, gen, kw):
    topo, free, pods = gen(**kw)
    want = oracle_b.score_batch(topo, free, pods, node_id_base=1000, fast=True, nthreads=8)
    got = _score(scorer, variant, topo, free, pods, base=1000)
    assert (got == want).all()


@pytest.mark.parametrize("variant", VARIANTS)
def test_edge_cases(scorer, oracle_b, variant):
    rng = np.random.default_rng(5)
    # ragged sizes around the tile (128/64 nodes) and chunk (512 pods) boundaries, all k incl. invalid
    for N, P in [(1, 1), (127, 3), (128, 31), (129, 33), (640, 513), (1, 1100)]:
        topo = rng.integers(0, 16, size=(N, 64)).astype(np.int32)
        free = rng.integers(0, 256, size=N).astype(np.int32)
        pods = synth.make_pods(rng.integers(-2, 11, size=P).astype(np.int32))
        W = rng.integers(0, 4096, size=16).astype(np.int32)
        want = oracle_b.score_batch(topo, free, pods, W, node_id_base=2**32 - 1 - N)
        got = _score(scorer, variant, topo, free, pods, base=2**32 - 1 - N, W=W)
        assert (got == want).all(), (N, P)
    # nothing free anywhere / empty node list / empty pod list
    topo, free, pods = synth.gen_c2(N=300, P=40)
    assert (_score(scorer, variant, topo, np.zeros_like(free), pods) == np.uint64(_lib.NO_FIT)).all()
    assert (_score(scorer, variant, topo[:0], free[:0], pods) == np.uint64(_lib.NO_FIT)).all()
    assert _score(scorer, variant, topo, free, pods[:0]).shape == (0,)
    # maximum weights: cost field at its largest, no overflow into the penalty range
    wmax = np.full(16, 4095, dtype=np.int32)
    full = np.full(200, 0xFF, dtype=np.int32)
    got = _score(scorer, variant, topo[:200], full, synth.make_pods(np.arange(0, 9, dtype=np.int32)), W=wmax)
    want = oracle_b.score_batch(topo[:200], full, synth.make_pods(np.arange(0, 9, dtype=np.int32)), wmax)
    assert (got == want).all() and int(got[8]) >> 40 == 28 * 4095


@pytest.mark.parametrize("variant", [_lib.VARIANT_SPARSE, _lib.VARIANT_LANE_PER_NODE])
def test_update_and_remove_node(scorer, oracle_b, variant):
    """Also the stale-order case of K1s: masks change after upload, the order does not."""
    scorer.set_variant(variant)
    topo, free, pods = synth.gen_c2(N=2000, P=64)
    topo, free = topo.copy(), free.copy()
    scorer.upload_nodes(topo, free)
    base = scorer.score_batch(pods)
    assert (base == oracle_b.score_batch(topo, free, pods)).all()
    # make node 1500 perfect (all NVLink, all free): every pod must move there
    topo[1500] = 9
    free[1500] = 0xFF
    scorer.update_node(1500, topo[1500], 0xFF)
    got = scorer.score_batch(pods)
    assert (got == oracle_b.score_batch(topo, free, pods)).all()
    assert all(((int(k) >> 8) & 0xFFFFFFFF) in (1500, ((int(b) >> 8) & 0xFFFFFFFF)) for k, b in zip(got, base))
    scorer.remove_node(1500)
    free[1500] = 0
    assert (scorer.score_batch(pods) == oracle_b.score_batch(topo, free, pods)).all()
    scorer.set_free_mask(7, 0x0F)
    free[7] = 0x0F
    assert (scorer.score_batch(pods) == oracle_b.score_batch(topo, free, pods)).all()
    assert scorer.num_nodes == 2000


def test_error_paths(scorer):
    from kubegpu_b200.scorer import KgpuError
    topo, free, pods = synth.gen_c2(N=10, P=4)
    bad = topo.copy()
    bad[3, 5] = 16
    with pytest.raises(KgpuError) as e:
        scorer.upload_nodes(bad, free)
    assert e.value.code == _lib.ERR_INVALID and "topo[3][5]" in str(e.value)
    with pytest.raises(KgpuError):
        scorer.set_weights([4096] + [0] * 15)
    with pytest.raises(KgpuError):
        scorer.set_variant(9)
    scorer.upload_nodes(topo, free)
    with pytest.raises(KgpuError):
        scorer.update_node(10, topo[0], 0xFF)
    with pytest.raises(KgpuError):
        scorer.upload_nodes(topo, free, node_id_base=2**32)


def test_variants_agree_at_full_c2_size(scorer, oracle_b):
    """BASELINE config 2 (100k nodes x 10k pods): the oracle is too slow for all of it, so
    check (a) the three variants agree bit for bit, (b) an oracle-checked pod sample,
    (c) per-k collapse: pods with equal k get equal keys (snapshot scoring), and
    (d) every reported (node, mask) re-scores to the reported cost."""
    topo, free, pods = synth.gen_c2()
    outs = {v: _score(scorer, v, topo, free, pods) for v in VARIANTS}
    assert all((outs[VARIANTS[0]] == outs[v]).all() for v in VARIANTS[1:])
    keys = outs[VARIANTS[0]]
    sample = np.arange(0, 10_000, 625)
    want = oracle_b.score_batch(topo, free, pods[sample], fast=True, nthreads=8)
    assert (keys[sample] == want).all()
    for k in (1, 2, 4, 8):
        assert len(set(keys[pods[:, 0] == k].tolist())) == 1
    for key, k in zip(keys[:64], pods[:64, 0]):
        cost, node, mask = int(key) >> 40, (int(key) >> 8) & 0xFFFFFFFF, int(key) & 0xFF
        assert bin(mask).count("1") == k and (mask & ~int(free[node])) == 0
        assert oracle_b.lib().kgpu_oracle_subset_cost(
            topo[node].ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int32)), mask,
            oracle_b.DEFAULT_WEIGHTS.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int32))) == cost


def test_heterogeneous_256k_bit_exact(scorer, oracle_b):
    """BASELINE config 4: 256k heterogeneous nodes, every key compared with Oracle B."""
    topo, free, pods = synth.gen_c4(N=262_144, P=64)
    want = oracle_b.score_batch(topo, free, pods, fast=True, nthreads=8)
    for v in (VARIANTS[0], VARIANTS[1], VARIANTS[4]):
        assert (_score(scorer, v, topo, free, pods) == want).all()


def test_device_buffer_entry_point_and_k2(scorer, oracle_b):
    import torch
    scorer.set_variant(_lib.VARIANT_SPARSE)
    topo, free, pods = synth.gen_c3(N=5000, P=333)
    G, per = 4, 1250
    d_pods = torch.from_numpy(pods).cuda()
    gathered = torch.empty((G, 333), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for g in range(G):
        scorer.upload_nodes(topo[g * per:(g + 1) * per], free[g * per:(g + 1) * per], node_id_base=g * per)
        scorer.score_batch_device(d_pods.data_ptr(), 333, gathered[g].data_ptr(), st)
        torch.cuda.synchronize()
    out = torch.empty(333, dtype=torch.int64, device="cuda")
    scorer.reduce_shards_device(gathered.data_ptr(), G, 333, out.data_ptr(), st)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint64)
    assert (got == oracle_b.score_batch(topo, free, pods, fast=True, nthreads=4)).all()
    assert scorer.kernel_launches > 0


def test_one_million_nodes_c3_scale(scorer, oracle_b):
    """BASELINE config 3's node count on one GPU (1M nodes = 256 MB of topology), k in 1..8:
    an oracle-checked pod sample, per-k collapse, and agreement of the sharded flow (8 shards
    scored one after the other + K2) with the unsharded launch."""
    import torch
    topo, free, pods = synth.gen_c3(N=1_000_000, P=2048)
    scorer.set_variant(_lib.VARIANT_SPARSE)
    scorer.upload_nodes(topo, free)
    keys = scorer.score_batch(pods)
    sample = np.arange(0, 2048, 128)
    assert (keys[sample] == oracle_b.score_batch(topo, free, pods[sample], fast=True, nthreads=8)).all()
    for k in range(1, 9):
        assert len(set(keys[pods[:, 0] == k].tolist())) == 1
    G, per = 8, 125_000
    d_pods = torch.from_numpy(pods).cuda()
    gathered = torch.empty((G, 2048), dtype=torch.int64, device="cuda")
    out = torch.empty(2048, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for g in range(G):
        scorer.upload_nodes(topo[g * per:(g + 1) * per], free[g * per:(g + 1) * per], node_id_base=g * per)
        scorer.score_batch_device(d_pods.data_ptr(), 2048, gathered[g].data_ptr(), st)
        torch.cuda.synchronize()
    scorer.reduce_shards_device(gathered.data_ptr(), G, 2048, out.data_ptr(), st)
    torch.cuda.synchronize()
    assert (out.cpu().numpy().view(np.uint64) == keys).all()


def test_ten_million_nodes_sweep_point(scorer, oracle_b):
    """Largest point of BASELINE config 5's sweep: 10M nodes (2.56 GB of topology in HBM).
    Few pods (the oracle has to follow); checks 64-bit indexing and node ids near 10^7."""
    N = 10_000_000
    topo, free, pods = synth.gen_c2(N=N, P=64, seed=synth.SEED_C5)
    # make the very last node the unique best home for k=8 so the answer must come from the far end
    topo[N - 1] = 9
    free[N - 1] = 0xFF
    scorer.set_variant(_lib.VARIANT_SPARSE)
    scorer.upload_nodes(topo, free)
    keys = scorer.score_batch(pods)
    sample = np.array([0, 1, 2, 3, 17, 40, 63])
    assert (keys[sample] == oracle_b.score_batch(topo, free, pods[sample], fast=True, nthreads=8)).all()
    k8 = keys[pods[:, 0] == 8]
    assert len(k8) and all(((int(k) >> 8) & 0xFFFFFFFF) == N - 1 and (int(k) >> 40) == 0 for k in k8)
    scorer.upload_nodes(topo[:1], free[:1])          # release the 2.5 GB before the next test


def test_few_pod_instantiations_at_their_boundaries(scorer, oracle_b):
    """The host picks the K1s instantiation by pod count: TMA-staged tiles with 7 blocks per SM up to 4 pods, with 8
    blocks up to 64, the register-prefetch one up to 512, the plain one above.  Every boundary, on heterogeneous nodes
    (runs of tiles across class boundaries, ragged permutations), with a k = 0 and an invalid pod riding along."""
    topo, free, _ = synth.gen_c4(N=300_000, P=1)
    scorer.set_variant(_lib.VARIANT_SPARSE)
    scorer.upload_nodes(topo, free, node_id_base=5)
    for P in (1, 2, 4, 5, 31, 32, 33, 63, 64, 65, 200, 512, 513):
        pods = synth.make_pods((1 + synth.rand_below(77, P, P, 8)).astype(np.int32))
        pods[0, 0] = 0
        if P > 3:
            pods[3, 0] = 9
        want = oracle_b.score_batch(topo, free, pods, node_id_base=5, fast=True, nthreads=8)
        assert (scorer.score_batch(pods) == want).all(), P
    scorer.upload_nodes(topo[:1], free[:1])


def test_agree_set_on_gpu_against_reference_greedy(golden_dir):
    """the chain GPU -> Oracle A -> reference goldens on hardware.  All 223 two-level
    8-GPU shapes x k = 1..8 (1784 cases) go through kgpu_score_pairs; the GPU's optimal cost is compared
    with the cost of the subset the reference's greedy tree fill picks (Oracle A, which reproduces
    gpu_test.go:61-109): equal on 1759 cases, and the divergent set is exactly the committed 25
    (tests/golden/agree_set.json), every one a case where the GPU is cheaper."""
    import json
    import sys
    sys.path.insert(0, golden_dir)
    import make_golden
    from oracle import oracle_a as oa
    from oracle import oracle_b as ob
    from kubegpu_b200.scorer import Scorer
    W = ob.DEFAULT_WEIGHTS
    shapes = make_golden.two_level_shapes(8)
    assert len(shapes) == 223
    mats, trees = [], []
    for shp in shapes:
        tree = oa.add_to_node(None, oa.shape_to_resources([list(g) for g in shp]), "gpugrp", "cards", 1)
        trees.append(tree)
        mats.append(np.asarray(oa.tree_to_matrix(tree), dtype=np.int32).reshape(64))
    topo = np.stack(mats)
    free = np.full(len(shapes), 0xFF, dtype=np.int32)
    idx = np.repeat(np.arange(len(shapes), dtype=np.int64), 8)
    ks = np.tile(np.arange(1, 9, dtype=np.int32), len(shapes))
    with Scorer((0,)) as s:
        s.upload_nodes(topo, free)
        got = s.score_pairs(idx, ks)
    divergent = []
    for i, k, nk in zip(idx, ks, got):
        assert nk != 0xFFFFFFFF                       # all 8 GPUs are free: every k fits
        greedy = oa.greedy_fill_mask(trees[i], int(k))
        M = mats[i]
        gcost = sum(int(W[M[a * 8 + b]]) for a in range(8) for b in range(a + 1, 8) if (greedy >> a) & 1 and (greedy >> b) & 1)
        gpu_cost = int(nk) >> 8
        assert gpu_cost <= gcost                      # the subset search is never worse than the greedy fill
        if gpu_cost != gcost:
            divergent.append({"shape": [list(g) for g in shapes[i]], "k": int(k), "greedy_cost": gcost, "optimal_cost": gpu_cost})
    with open(os.path.join(golden_dir, "agree_set.json")) as f:
        committed = json.load(f)
    assert len(idx) == committed["cases"] == 1784
    assert len(idx) - len(divergent) == committed["agree"] == 1759
    assert divergent == [{k: d[k] for k in ("shape", "k", "greedy_cost", "optimal_cost")} for d in committed["divergent"]]

"""State changes on the device: batched free-mask updates with an
incremental refresh of the scorer's records, the device-side re-sort, the (node, k) fit table that serves
PodFitsDevice, the dry-run placement, the upload's device-side value check.  All through the C ABI, bit-exact
against the oracle."""
import numpy as np
import pytest

from kubegpu_b200 import _lib, synth

pytestmark = pytest.mark.gpu


@pytest.fixture()
def scorer():
    from kubegpu_b200.scorer import Scorer
    s = Scorer((0,))
    yield s
    s.close()


def test_batched_mask_updates_track_the_oracle(scorer, oracle_b):
    topo, free, pods = synth.gen_c4(N=20_000, P=300)
    f = free.copy()
    scorer.upload_nodes(topo, f)
    assert scorer.last_upload_ms > 0.0
    assert (scorer.score_batch(pods) == oracle_b.score_batch(topo, f, pods, fast=True, nthreads=8)).all()
    rng = np.random.default_rng(7)
    for rnd in range(8):                       # 8 x 1500 changed nodes: crosses the re-sort threshold (n / 200) every round
        idx = rng.integers(0, len(f), size=1500).astype(np.int64)         # duplicates happen: the last one wins
        masks = rng.integers(0, 256, size=1500).astype(np.int32)
        scorer.set_free_masks(idx, masks)
        for i, m in zip(idx, masks):
            f[i] = m
        got = scorer.score_batch(pods)
        assert (got == oracle_b.score_batch(topo, f, pods, fast=True, nthreads=8)).all(), "round %d" % rnd
    assert (scorer.get_free_masks() == f).all()
    # single-node calls ride the same path
    scorer.set_free_mask(123, 0xFF)
    scorer.remove_node(124)
    f[123], f[124] = 0xFF, 0
    assert (scorer.score_batch(pods) == oracle_b.score_batch(topo, f, pods, fast=True, nthreads=8)).all()


def test_fit_table_serves_pod_fits_device(scorer, oracle_b):
    topo, free, mem, pods = synth.gen_c6(N=5_000, P=10)
    f = free.copy()
    scorer.upload_nodes(topo, f)
    scorer.upload_gpu_memory(mem)
    rng = np.random.default_rng(11)
    launches0 = scorer.kernel_launches
    scorer.build_fit_table()
    built = scorer.kernel_launches
    assert built == launches0 + 1
    for _ in range(300):
        n, k = int(rng.integers(0, len(f))), int(rng.integers(0, 9))
        assert scorer.fit_lookup(n, k) == oracle_b.node_key(topo[n], int(f[n]), k)
    assert scorer.fit_lookup(5, 9) == 0xFFFFFFFF
    assert scorer.kernel_launches == built                 # lookups never launch
    # Take / Return keep the table current row by row
    idx = rng.choice(len(f), size=64, replace=False).astype(np.int64)
    masks = rng.integers(0, 256, size=64).astype(np.int32)
    scorer.set_free_masks(idx, masks)
    f[idx] = masks
    t2 = topo[17].copy()
    t2[:] = 9
    scorer.update_node(17, t2, 0x3F)
    topo = topo.copy()
    topo[17], f[17] = t2, 0x3F
    for n in list(idx[:20]) + [17, 0, 4999]:
        for k in range(9):
            assert scorer.fit_lookup(int(n), k) == oracle_b.node_key(topo[n], int(f[n]), k)
    # kgpu_score_pairs: table for plain pairs, one packed launch for pairs with a memory requirement
    qn = rng.integers(0, len(f), size=200).astype(np.int64)
    qk = rng.integers(0, 9, size=200).astype(np.int32)
    got = scorer.score_pairs(qn, qk)
    assert [int(x) for x in got] == [oracle_b.node_key(topo[n], int(f[n]), int(k)) for n, k in zip(qn, qk)]
    qm = np.array(synth.POD_MIN_MEM_CHOICES_MIB, dtype=np.int32)[rng.integers(0, 7, size=200)]
    got = scorer.score_pairs(qn, qk, qm)
    for n, k, m, g in zip(qn, qk, qm, got):
        ok = sum(1 << i for i in range(8) if mem[n, i] >= m) if m > 0 else 0xFF
        assert int(g) == oracle_b.node_key(topo[n], int(f[n]) & ok, int(k))
    # weights change -> table rebuilt lazily
    w = np.array([100, 50, 25, 12, 6, 3, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=np.int32)
    scorer.set_weights(w)
    assert scorer.fit_lookup(42, 3) == oracle_b.node_key(topo[42], int(f[42]), 3, w)


def test_dry_run_placement_leaves_the_state_alone(scorer, oracle_b):
    topo, free, pods = synth.gen_c2(N=3_000, P=400)
    scorer.upload_nodes(topo, free)
    before = scorer.score_batch(pods)
    want, wf = oracle_b.place_batch(topo, free, pods)
    got = scorer.place_batch(pods, dry_run=True)
    assert (got == want).all()
    assert (scorer.get_free_masks() == free).all()
    assert (scorer.score_batch(pods) == before).all()
    # proposals are conflict free: no GPU handed out twice
    used = {}
    for key in got:
        if key != np.uint64(_lib.NO_FIT):
            node, mask = int(key >> np.uint64(8)) & 0xFFFFFFFF, int(key & np.uint64(0xFF))
            assert used.get(node, 0) & mask == 0
            used[node] = used.get(node, 0) | mask
    # committing them through the batched Take gives the state the real placement leaves
    idx = np.array(sorted(used), dtype=np.int64)
    scorer.set_free_masks(idx, np.array([int(free[i]) & ~used[i] for i in idx], dtype=np.int32))
    assert (scorer.get_free_masks() == wf).all()
    assert (scorer.place_batch(pods) == oracle_b.place_batch(topo, wf, pods)[0]).all()


def test_upload_rejects_out_of_domain_levels_and_stays_usable(scorer, oracle_b):
    from kubegpu_b200.scorer import KgpuError
    topo, free, pods = synth.gen_c2(N=2_000, P=16)
    scorer.upload_nodes(topo, free)
    bad = topo.copy()
    bad[1234, 17] = 16
    with pytest.raises(KgpuError) as e:
        scorer.upload_nodes(bad, free)
    assert e.value.code == _lib.ERR_INVALID and "topo[1234][17] = 16" in str(e.value)
    assert scorer.num_nodes == 0                            # never half old / half new
    assert (scorer.score_batch(pods) == np.uint64(_lib.NO_FIT)).all()
    scorer.upload_nodes(topo, free)
    assert (scorer.score_batch(pods) == oracle_b.score_batch(topo, free, pods, fast=True, nthreads=8)).all()

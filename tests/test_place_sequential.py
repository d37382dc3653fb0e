"""K3 stateful sequential placement: oracle self-checks on CPU, and the
device path (kgpu_place_batch) bit-exact against it on the GPU."""
import numpy as np
import pytest

from kubegpu_b200 import synth


def test_oracle_cached_equals_plain_and_conserves_gpus(oracle_b):
    ob = oracle_b
    for gen, kw in ((synth.gen_c2, dict(N=300, P=400)), (synth.gen_c4, dict(N=257, P=300)), (synth.gen_c1, {})):
        topo, free, pods = gen(**kw)
        if len(pods) < 100:
            pods = synth.make_pods(np.array([3, 4, 1, 2, 8, 4, 4, 2, 0, 9, 1, 1] * 3, dtype=np.int32))
        k1, f1 = ob.place_batch(topo, free, pods, plain=True)
        k2, f2 = ob.place_batch(topo, free, pods)
        assert (k1 == k2).all() and (f1 == f2).all()
        # every placement used GPUs that were free when it happened, and the totals add up
        fm = free.astype(np.int64).copy() & 0xFF
        for key, k in zip(k1, pods[:, 0]):
            u = ob.unpack_key(key)
            if u is None:
                continue
            cost, node, mask = u
            assert bin(mask).count("1") == k and (mask & ~fm[node]) == 0
            assert ob.lib().kgpu_oracle_subset_cost is not None
            fm[node] &= ~mask
        assert (fm == f1).all()
        # a pod that did not fit really has no home in the final state or was blocked earlier: re-score it
        unfit = np.flatnonzero(k1 == ob.NO_FIT)
        if len(unfit):
            last = unfit[-1]
            later_frees_nothing = ob.score_batch(topo, f1, pods[last:last + 1])
            assert later_frees_nothing[0] == ob.NO_FIT     # masks only shrink, so it still cannot fit at the end


def test_sequential_differs_from_snapshot(oracle_b):
    """With snapshot scoring every pod of equal k gets the same answer; sequentially they spread."""
    topo, free, pods = synth.gen_c2(N=64, P=40)
    snap = oracle_b.score_batch(topo, free, pods)
    seq, _ = oracle_b.place_batch(topo, free, pods)
    same_k = pods[:, 0] == 4
    assert len(set(snap[same_k].tolist())) == 1
    fit = seq[same_k][seq[same_k] != oracle_b.NO_FIT]
    assert len(set(fit.tolist())) == len(fit) > 1


@pytest.mark.gpu
@pytest.mark.parametrize("gen,kw", [
    (synth.gen_c1, {}),
    (synth.gen_c2, dict(N=5000, P=3000)),
    (synth.gen_c3, dict(N=12_345, P=4000)),
    (synth.gen_c4, dict(N=20_000, P=2500)),
    (synth.gen_c4, dict(N=1, P=40)),
    (synth.gen_c4, dict(N=129, P=700)),
])
def test_place_batch_matches_oracle(oracle_b, gen, kw):
    from kubegpu_b200.scorer import Scorer
    topo, free, pods = gen(**kw)
    if len(pods) < 20:
        pods = synth.make_pods(np.array([3, 4, 1, 2, 8, 4, 4, 2, 0, 9, -1, 1, 1] * 4, dtype=np.int32))
    want_keys, want_free = oracle_b.place_batch(topo, free, pods, node_id_base=77)
    with Scorer((0,)) as s:
        s.upload_nodes(topo, free, node_id_base=77)
        got = s.place_batch(pods)
        assert (got == want_keys).all()
        assert (s.get_free_masks() == want_free).all()
        # the state persists: a second cycle continues from the first one's masks
        more_keys, more_free = oracle_b.place_batch(topo, want_free, pods[:200], node_id_base=77)
        assert (s.place_batch(pods[:200]) == more_keys).all() and (s.get_free_masks() == more_free).all()
        # and snapshot scoring now sees the reduced cluster
        assert (s.score_batch(pods[:64]) == oracle_b.score_batch(topo, more_free, pods[:64], node_id_base=77)).all()


@pytest.mark.gpu
def test_place_batch_full_c2(oracle_b):
    """BASELINE config 2 sequentially: 10k pods onto 100k nodes, every key and every mask."""
    from kubegpu_b200.scorer import Scorer
    topo, free, pods = synth.gen_c2()
    want_keys, want_free = oracle_b.place_batch(topo, free, pods)
    with Scorer((0,)) as s:
        s.upload_nodes(topo, free)
        got = s.place_batch(pods)
        assert (got == want_keys).all() and (s.get_free_masks() == want_free).all()
        print("K3 place_batch: %.3f ms for %d pods -> %.0f placements/s" % (s.last_kernel_ms, len(pods), len(pods) / s.last_kernel_ms * 1e3))


@pytest.mark.gpu
@pytest.mark.parametrize("N,P", [(300, 150), (20_000, 3000), (128 * 40 + 3, 2000)])
def test_place_batch_memory_aware(oracle_b, N, P):
    """Pods with min_mem placed sequentially: one table set per distinct requirement (4 + the plain one)."""
    from kubegpu_b200.scorer import Scorer
    topo, free, mem, pods = synth.gen_c6(N=N, P=P)
    want_keys, want_free = oracle_b.place_batch(topo, free, pods, node_id_base=11, mem=mem)
    plain = pods.copy()
    plain[:, 3] = 0
    plain_keys, plain_free = oracle_b.place_batch(topo, free, plain, node_id_base=11)
    assert (plain_keys != want_keys).any()
    with Scorer((0,)) as s:
        s.upload_nodes(topo, free, node_id_base=11)
        # no GPU memory uploaded: the requirement excludes nothing
        assert (s.place_batch(pods) == plain_keys).all() and (s.get_free_masks() == plain_free).all()
        s.upload_nodes(topo, free, node_id_base=11)
        s.upload_gpu_memory(mem)
        assert (s.place_batch(pods) == want_keys).all() and (s.get_free_masks() == want_free).all()
        # next cycle continues on the reduced cluster, snapshot scoring sees it too
        more_keys, more_free = oracle_b.place_batch(topo, want_free, pods[:100], node_id_base=11, mem=mem)
        assert (s.place_batch(pods[:100]) == more_keys).all() and (s.get_free_masks() == more_free).all()
        assert (s.score_batch(pods[:64]) == oracle_b.score_batch(topo, more_free, pods[:64], node_id_base=11, mem=mem)).all()

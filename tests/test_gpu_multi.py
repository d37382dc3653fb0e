"""Multi-GPU paths on the GPU box (need >= 2 devices; skipped otherwise):
 - the single-process handle kgpu_create(devs, n>1): shards + NCCL all-gather + K2 inside libkgpu
 - per-pair queries routed to the right shard."""
import numpy as np
import pytest

from kubegpu_b200 import synth

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("G", [2, 4, 8])
def test_multi_device_handle_matches_oracle(oracle_b, G):
    if _ndev() < G:
        pytest.skip("needs %d GPUs" % G)
    from kubegpu_b200.scorer import Scorer
    topo, free, pods = synth.gen_c4(N=50_001, P=321)
    want = oracle_b.score_batch(topo, free, pods, node_id_base=17, fast=True, nthreads=8)
    with Scorer(tuple(range(G))) as s:
        s.upload_nodes(topo, free, node_id_base=17)
        assert s.num_nodes == 50_001
        assert (s.score_batch(pods) == want).all()
        # per-pair queries land on whichever shard owns the node
        idx = np.array([0, 1, 25_000, 25_001, 49_999, 50_000], dtype=np.int64)
        ks = np.array([1, 2, 3, 4, 5, 8], dtype=np.int32)
        got = s.score_pairs(idx, ks)
        for i, k, g in zip(idx, ks, got):
            assert int(g) == oracle_b.node_key(topo[i], int(free[i]), int(k))
        # update a node on the last shard and rescore
        topo2, free2 = topo.copy(), free.copy()
        topo2[50_000] = 9
        free2[50_000] = 0xFF
        s.update_node(50_000, topo2[50_000], 0xFF)
        assert (s.score_batch(pods) == oracle_b.score_batch(topo2, free2, pods, node_id_base=17, fast=True, nthreads=8)).all()

"""N > 1 host logic on CPU: world_size-2 gloo processes shard the node list, exchange
per-pod bests with one all-gather and pick the column minimum; the result must equal the
unsharded answer.  (Local scoring uses the oracle here; on the GPU box the same flow runs
with K1/K2 in bench.py and tests/test_gpu_multi.py.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from kubegpu_b200.distributed import shard_range
    for n in (0, 1, 7, 100_000, 262_144, 1_000_003):
        for G in (1, 2, 3, 4, 8):
            rs = [shard_range(n, G, r) for r in range(G)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(hi - lo for lo, hi in rs) <= (n + G - 1) // G
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kubegpu_b200 import synth
    from kubegpu_b200.distributed import all_gather_keys, shard_range
    from oracle import oracle_b
    N, P = 3001, 77
    lo, hi = shard_range(N, world, rank)
    topo, free, pods = synth.gen_c4(N=hi - lo, P=P, node_start=lo)       # each rank generates only its shard
    local = oracle_b.score_batch(topo, free, pods, node_id_base=lo)
    gathered = all_gather_keys(torch.from_numpy(local.view(np.int64)))
    final = oracle_b.reduce_shards(gathered.numpy().view(np.uint64))
    # the one-collective alternative must agree (sign-flipped signed MIN == unsigned min, NO_FIT included)
    from kubegpu_b200.distributed import all_reduce_min_keys
    reduced = all_reduce_min_keys(torch.from_numpy(local.copy().view(np.int64))).numpy().view(np.uint64)
    assert (reduced == final).all()
    if rank == 0:
        np.save(result_path, final)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_flow_equals_unsharded(tmp_path, oracle_b, world):
    from kubegpu_b200 import synth
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "final.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    topo, free, pods = synth.gen_c4(N=3001, P=77)
    assert (np.load(out) == oracle_b.score_batch(topo, free, pods)).all()

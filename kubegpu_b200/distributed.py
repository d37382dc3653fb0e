"""Multi-GPU plumbing for the one-process-per-GPU launch (torchrun): how the candidate
node list is sharded and how shard results are exchanged.

Nodes are independent and a pod's answer is a min over nodes, so rank r of G scores the
contiguous node range ``shard_range(N, G, r)`` (global node ids stay in the keys), one
all-gather moves every rank's uint64[P] bests, and K2 (``kgpu_reduce_shards_device``)
takes the column minimum on every rank.  The result is independent of G.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_nodes: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the nodes rank `rank` of `world` holds: ceil(N/G) per rank, contiguous
    (the same split ``kgpu_upload_nodes`` uses inside a multi-device handle)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = (n_nodes + world - 1) // world
    lo = min(n_nodes, rank * per)
    return lo, min(n_nodes, lo + per)


def all_gather_keys(local_keys, group=None):
    """One all-gather of each rank's per-pod best keys (a torch int64 view of the uint64
    keys, CPU/gloo or CUDA/NCCL).  Returns a [world, P] tensor, identical on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world, local_keys.numel()), dtype=local_keys.dtype, device=local_keys.device)
    dist.all_gather_into_tensor(out.view(-1), local_keys.contiguous(), group=group)
    return out


_SIGN = -(1 << 63)


def all_reduce_min_keys(local_keys, group=None):
    """One all-reduce instead of all-gather + K2: the per-pod minimum over the ranks' uint64 keys, given
    and returned as their int64 view (in place).  Collectives reduce SIGNED integers, so the sign bit is
    flipped around the reduction (an order-preserving map of uint64 onto int64; NO_FIT stays the largest)."""
    import torch
    import torch.distributed as dist
    sign = torch.tensor(_SIGN, dtype=torch.int64, device=local_keys.device)
    local_keys.bitwise_xor_(sign)
    dist.all_reduce(local_keys, op=dist.ReduceOp.MIN, group=group)
    local_keys.bitwise_xor_(sign)
    return local_keys

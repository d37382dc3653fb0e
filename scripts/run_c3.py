#!/usr/bin/env python3
"""BASELINE config 3 at full size: 1M nodes x 100k pods, k uniform 1..8, node list sharded over the
ranks, one NCCL all-gather + K2.  Launch: python -m torch.distributed.run --nproc-per-node 8
--master-addr 127.0.0.1 scripts/run_c3.py [--nodes N --pods P].  Rank 0 checks a pod sample against
the oracle and prints one JSON line."""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from kubegpu_b200 import _lib, synth
from kubegpu_b200.distributed import shard_range
from kubegpu_b200.scorer import Scorer

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=1_000_000)
ap.add_argument("--pods", type=int, default=100_000)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
lo, hi = shard_range(a.nodes, world, rank)
topo, free, pods = synth.gen_c3(hi - lo, a.pods, node_start=lo)
s = Scorer((local,))
s.upload_nodes(topo, free, node_id_base=lo)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
d_pods = torch.from_numpy(pods).to(dev)
d_local = torch.empty(a.pods, dtype=torch.int64, device=dev)
d_gather = torch.empty((world, a.pods), dtype=torch.int64, device=dev)
d_final = torch.empty(a.pods, dtype=torch.int64, device=dev)

def step():
    s.score_batch_device(d_pods.data_ptr(), a.pods, d_local.data_ptr(), st.cuda_stream, _lib.BATCH_NO_MIN_MEM)
    if world > 1:
        dist.all_gather_into_tensor(d_gather.view(-1), d_local)
        s.reduce_shards_device(d_gather.data_ptr(), world, a.pods, d_final.data_ptr(), st.cuda_stream)
    else:
        d_final.copy_(d_local)

s.score_batch_device(d_pods.data_ptr(), 1024, d_local.data_ptr(), st.cuda_stream)   # warm-up on a slice
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
step()
e1.record(st)
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
keys = d_final.cpu().numpy().view(np.uint64)
if rank == 0:
    from oracle import oracle_b
    full_topo, full_free, _ = synth.gen_c3(a.nodes, 0)
    sample = np.arange(0, a.pods, max(1, a.pods // 24))[:24]
    want = oracle_b.score_batch(full_topo, full_free, pods[sample], fast=True, nthreads=min(24, os.cpu_count() or 1))
    ok = bool((keys[sample] == want).all())
    per_k = all(len(set(keys[pods[:, 0] == k].tolist())) == 1 for k in range(1, 9))
    t = float(ms.item())
    print(json.dumps({"config": "C3: %d nodes x %d pods, k uniform 1..8" % (a.nodes, a.pods), "n_gpus": world, "ms": t,
                      "placements_per_s": a.pods / (t * 1e-3), "pairs_per_s": a.nodes * a.pods / (t * 1e-3),
                      "algorithmic_GBps_total": 260.0 * a.nodes * a.pods / (t * 1e-3) / 1e9,
                      "oracle_sample_ok": ok, "per_k_collapse_ok": per_k, "keys_sha256_12": hashlib.sha256(keys.tobytes()).hexdigest()[:12]}))
s.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

#!/usr/bin/env python3
"""BASELINE config C5 on one GPU: node-count sweep x pod-count sweep of the headline kernel.
For every N the node array is generated once (C2's distributions, seed 0xB2000005), uploaded (timed:
upload_ms) and scored for P in --pods; prints one JSON line per (N, P) with placements/s, pairs/s,
algorithmic GB/s (260 B per pair + 24 B per pod) and the GB/s the kernel must at least stream from DRAM
(its node records once per launch).  Small P at large N is the regime where HBM is the roof.
Usage: python scripts/c5_time.py [--nodes 10000,100000,1000000,10000000] [--pods 1,32,10000] [--reps 5]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kubegpu_b200 import _lib, synth
from kubegpu_b200.scorer import Scorer

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", default="10000,100000,1000000,10000000")
ap.add_argument("--pods", default="1,32,10000")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--variant", type=int, default=_lib.VARIANT_SPARSE)
ap.add_argument("--lib", default=None)
ap.add_argument("--stream-bytes", type=int, default=124, help="bytes per node the kernel reads once per launch")
a = ap.parse_args()
if a.lib:
    _lib.LIB_PATH = os.path.abspath(a.lib)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N in [int(x) for x in a.nodes.split(",")]:
    t0 = time.perf_counter()
    topo, free, _ = synth.gen_c2(N, 0, seed=synth.SEED_C5)
    gen_s = time.perf_counter() - t0
    s = Scorer((0,))
    s.set_variant(a.variant)
    t0 = time.perf_counter()
    s.upload_nodes(topo, free)
    upload_ms = 1e3 * (time.perf_counter() - t0)
    for P in [int(x) for x in a.pods.split(",")]:
        _, _, pods = synth.gen_c2(0, P, seed=synth.SEED_C5)
        d_pods = torch.from_numpy(pods).cuda()
        d_keys = torch.empty(P, dtype=torch.int64, device="cuda")
        for _ in range(3):
            s.score_batch_device(d_pods.data_ptr(), P, d_keys.data_ptr(), st.cuda_stream, _lib.BATCH_NO_MIN_MEM)
        torch.cuda.synchronize()
        ms = []
        reps = a.reps if N * P < 2e11 else 2
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            s.score_batch_device(d_pods.data_ptr(), P, d_keys.data_ptr(), st.cuda_stream, _lib.BATCH_NO_MIN_MEM)
            e1.record(st)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        med = float(np.median(ms))
        print(json.dumps({"N": N, "P": P, "ms": med, "ms_min": min(ms), "placements_per_s": P / med * 1e3,
                          "gpairs_per_s": N * P / med / 1e6,
                          "algorithmic_gbs": (260.0 * N * P + 24.0 * P) / med / 1e6,
                          "streamed_gbs_min": a.stream_bytes * N / med / 1e6,
                          "upload_ms": upload_ms, "gen_s": gen_s}), flush=True)
    s.close()
    del topo, free

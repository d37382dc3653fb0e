#!/usr/bin/env python3
"""Times K1 alone (device-resident pods, CUDA events on the launch stream) on config C2
or C3 for one or more variants; checks the keys against a committed digest of the oracle.
Usage: python scripts/k1_time.py [--config c2|c3] [--variants 2,1,3] [--reps 10]"""
import argparse, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kubegpu_b200 import _lib, synth
from kubegpu_b200.scorer import Scorer

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--variants", default="2")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--nodes", type=int, default=100_000)
ap.add_argument("--pods", type=int, default=10_000)
ap.add_argument("--lib", default=None, help="time this libkgpu build instead of kubegpu_b200/lib/libkgpu.so (sweeps)")
a = ap.parse_args()
if a.lib:
    _lib.LIB_PATH = os.path.abspath(a.lib)
mem = None
if a.config == "c6":
    topo, free, mem, pods = synth.gen_c6(a.nodes, a.pods)
else:
    gen = synth.gen_c2 if a.config == "c2" else synth.gen_c3
    topo, free, pods = gen(a.nodes, a.pods)
s = Scorer((0,))
s.upload_nodes(topo, free)
if mem is not None:
    s.upload_gpu_memory(mem)
d_pods = torch.from_numpy(pods).cuda()
d_keys = torch.empty(a.pods, dtype=torch.int64, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for v in [int(x) for x in a.variants.split(",")]:
    s.set_variant(v)
    for _ in range(3):
        s.score_batch_device(d_pods.data_ptr(), a.pods, d_keys.data_ptr(), st.cuda_stream, 0 if mem is not None else _lib.BATCH_NO_MIN_MEM)
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        s.score_batch_device(d_pods.data_ptr(), a.pods, d_keys.data_ptr(), st.cuda_stream, 0 if mem is not None else _lib.BATCH_NO_MIN_MEM)
        e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    keys = d_keys.cpu().numpy()
    dig = hashlib.sha256(keys.tobytes()).hexdigest()[:12]
    if ref is None:
        ref = dig
    med = float(np.median(ms))
    pairs = a.nodes * a.pods
    print("variant %d  %s  N=%d P=%d  median %.4f ms  min %.4f  -> %.3f Mplacements/s, %.1f Gpairs/s, alg %.1f TB/s  keys %s %s"
          % (v, a.config, a.nodes, a.pods, med, min(ms), a.pods / med / 1e3, pairs / med / 1e6, 260 * pairs / med / 1e9, dig,
             "" if dig == ref else "MISMATCH vs first variant"), flush=True)

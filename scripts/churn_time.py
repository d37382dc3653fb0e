#!/usr/bin/env python3
"""Cost of state changes on C2: kgpu_set_free_masks of 1 % of the nodes, then a scoring step, for several re-sort
thresholds (KGPU_RESORT_DIV: the order is re-sorted when stale * DIV > n).  Run once per DIV value (env is read once)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kubegpu_b200 import _lib, synth
from kubegpu_b200.scorer import Scorer
topo, free, pods = synth.gen_c2()
s = Scorer((0,)); s.set_variant(_lib.VARIANT_SPARSE); s.upload_nodes(topo, free)
hp = torch.from_numpy(pods).pin_memory(); hk = torch.empty(len(pods), dtype=torch.int64).pin_memory()
for _ in range(3): s.score_batch_ptr(hp.data_ptr(), len(pods), hk.data_ptr())
t0 = time.perf_counter()
for _ in range(20): s.score_batch_ptr(hp.data_ptr(), len(pods), hk.data_ptr())
base = (time.perf_counter() - t0) / 20
rng = np.random.default_rng(1)
upd, step, kms = [], [], []
for it in range(30):
    idx = rng.choice(len(free), size=len(free) // 100, replace=False).astype(np.int64)
    masks = rng.integers(0, 256, size=len(idx)).astype(np.int32)
    t0 = time.perf_counter(); s.set_free_masks(idx, masks); t1 = time.perf_counter()
    s.score_batch_ptr(hp.data_ptr(), len(pods), hk.data_ptr()); t2 = time.perf_counter()
    upd.append(1e3 * (t1 - t0)); step.append(1e3 * (t2 - t1)); kms.append(s.last_kernel_ms)
print("DIV=%s  baseline step %.3f ms | set_free_masks median %.3f ms | following step: median %.3f max %.3f min %.3f ms | kernel median %.3f ms"
      % (os.environ.get("KGPU_RESORT_DIV", "8"), 1e3 * base, np.median(upd), np.median(step), max(step), min(step), np.median(kms)))
print("   steps:", " ".join("%.2f" % x for x in step))

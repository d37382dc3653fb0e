#!/usr/bin/env python3
"""Times K3 (kgpu_place_batch) on config C2 and checks it against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kubegpu_b200 import _lib, synth
from kubegpu_b200.scorer import Scorer
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])        # time this build instead (A/B on one box)
from oracle import oracle_b
topo, free, pods = synth.gen_c2()
want, wf = oracle_b.place_batch(topo, free, pods)
with Scorer((0,)) as s:
    ms = []
    for _ in range(4):
        s.upload_nodes(topo, free)
        got = s.place_batch(pods)
        ms.append(s.last_kernel_ms)
    ok = bool((got == want).all() and (s.get_free_masks() == wf).all())
    print("K3 place_batch C2: median %.3f ms -> %.0f placements/s, %.2f us/pod, exact=%s" % (np.median(ms), 1e4 / np.median(ms) * 1e3, np.median(ms) * 1e3 / 1e4, ok))

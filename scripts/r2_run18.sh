#!/bin/bash
# One box: parity of the few-pod paths, then the TMA instantiations (1 stage, 7 / 8 blocks per SM, permutation expanded in the
# flush, bucket mask) at 10M nodes, the register-prefetch path (KGPU_SP_TMA=0) beside it, and K3 with the prefetching idle warp.
set -u
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_state_changes.py tests/test_place_sequential.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from kubegpu_b200 import _lib, synth
from kubegpu_b200.scorer import Scorer
from oracle import oracle_b
topo, free, _ = synth.gen_c4(N=300_000, P=1)
with Scorer((0,)) as s:
    s.set_variant(_lib.VARIANT_SPARSE); s.upload_nodes(topo, free, node_id_base=5)
    for P in (1, 2, 4, 5, 31, 32, 33, 63, 64, 65, 200):
        pods = synth.make_pods((1 + synth.rand_below(77, P, P, 8)).astype(np.int32)); pods[0, 0] = 0
        if P > 3: pods[3, 0] = 9
        got = s.score_batch(pods); want = oracle_b.score_batch(topo, free, pods, node_id_base=5, fast=True, nthreads=8)
        print("P=%d %s" % (P, "ok" if (got == want).all() else "MISMATCH"))
PY
for tma in 1 0; do echo "== KGPU_SP_TMA=$tma"; KGPU_SP_TMA=$tma timeout 300 python scripts/c5_time.py --nodes 10000000 --pods 1,4,8,16,32,64 --stream-bytes 120 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('P=%d ms=%.4f min=%.4f hbm_frac=%.3f' % (d['P'], d['ms'], d['ms_min'], d['streamed_gbs_min'] / 6587.7))
"; done
echo "== K3"; timeout 200 python scripts/k3_time.py 2>&1 | tail -1

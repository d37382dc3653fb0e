#!/bin/bash
# One box: K3 with / without the prefetching idle warp; the TMA ring's stages x blocks per SM at 10M nodes.
set -u
V=kubegpu_b200/lib/variants
echo "== K3 default / helper / default / helper"
for l in kubegpu_b200/lib/libkgpu.so $V/libkgpu_k3helper.so kubegpu_b200/lib/libkgpu.so $V/libkgpu_k3helper.so; do timeout 200 python scripts/k3_time.py $l 2>&1 | tail -1; done
echo "== TMA ring variants, 10M nodes"
for t in default tma_s1_b8 tma_s1_b7 tma_s2_b5 tma_s3_b4; do
  lib=$V/libkgpu_$t.so; [ $t = default ] && lib=kubegpu_b200/lib/libkgpu.so
  echo "-- $t"
  timeout 200 python scripts/c5_time.py --lib $lib --nodes 10000000 --pods 1,16,32,64 --stream-bytes 120 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('P=%d ms=%.4f min=%.4f hbm_frac=%.3f' % (d['P'], d['ms'], d['ms_min'], d['streamed_gbs_min'] / 6587.7))
"
done

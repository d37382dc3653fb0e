#!/bin/bash
# A/B on one box: final flush builds the key once per item (new) against per tile (base), 10M nodes, few pods; parity of the new build first.
set -u
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "few_pod or edge or seeded" 2>&1 | tail -2
for lib in kubegpu_b200/lib/variants/libkgpu_base.so kubegpu_b200/lib/libkgpu.so; do echo "-- $lib"
timeout 200 python scripts/c5_time.py --lib $lib --nodes 10000000 --pods 1,16,32,64,200 --stream-bytes 120 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('P=%d ms=%.4f min=%.4f hbm_frac=%.3f' % (d['P'], d['ms'], d['ms_min'], d['streamed_gbs_min'] / 6587.7))
"; done
for lib in kubegpu_b200/lib/variants/libkgpu_base.so kubegpu_b200/lib/libkgpu.so; do timeout 100 python scripts/k1_time.py --lib $lib --config c2 --variants 5 --reps 10 | cut -c1-110; done

#!/bin/bash
set -u
for rep in 1 2; do for v in head ordbit; do echo "== $v"; timeout 200 python scripts/c5_time.py --lib kubegpu_b200/lib/variants/libkgpu_$v.so --nodes 10000000 --pods 1,32 --stream-bytes 120 2>&1 | cut -c1-110; python scripts/k1_time.py --lib kubegpu_b200/lib/variants/libkgpu_$v.so --config c2 --variants 5 --reps 10 | cut -c1-100; done; done

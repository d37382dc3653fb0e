#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== stream"; timeout 300 python scripts/c5_time.py --nodes 10000000 --pods 1,32,512 --stream-bytes 120 2>&1 | cut -c1-150
echo "== C2/C3"; python scripts/k1_time.py --config c2 --variants 5 --reps 10 | cut -c1-100; python scripts/k1_time.py --config c3 --variants 5 --reps 6 | cut -c1-100
echo "== shard"; python scripts/k1_time.py --config c2 --nodes 12500 --variants 5 --reps 10 | cut -c1-100

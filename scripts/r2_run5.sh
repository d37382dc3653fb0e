#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== K3 v3"; timeout 300 python scripts/k3_time.py 2>&1 | tail -2 | tee gpurun_out/k3_time_v3.txt
echo "== K3 tests"; timeout 600 python -m pytest tests/test_place_sequential.py tests/test_memory_aware.py -x -q -m gpu 2>&1 | tail -3
for div in 8 32 100 400; do KGPU_RESORT_DIV=$div timeout 200 python scripts/churn_time.py 2>&1 | tail -2; done | tee gpurun_out/churn.txt
echo "== shard 12500"; python scripts/k1_time.py --config c2 --nodes 12500 --variants 5 --reps 10 | cut -c1-120
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size --clock-control none -k regex:score_pairs_sparse -s 3 -c 1 python scripts/k1_time.py --config c2 --nodes 12500 --variants 5 --reps 1 2>&1 | grep -E "inst_executed|duration|issue_active|throughput|grid_size"
for mr in 1 4 16 64; do echo "== stream MAXRUN=$mr"; KGPU_SP_MAXRUN=$mr timeout 300 python scripts/c5_time.py --nodes 10000000 --pods 1,32 --stream-bytes 120 2>&1 | cut -c1-150; done | tee gpurun_out/stream_maxrun.txt

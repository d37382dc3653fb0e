#!/bin/bash
# ncu --set full of K1s on a 12.5k-node shard of C2 (what one of eight ranks scores) + the counters of the same launch
set -u
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_shard12k python scripts/k1_time.py --config c2 --variants 5 --reps 1 --nodes 12500 > gpurun_out/ncu5.log 2>&1; tail -1 gpurun_out/ncu5.log | cut -c1-200
timeout 200 python scripts/k1_time.py --config c2 --variants 5 --reps 10 --nodes 12500 | cut -c1-120
timeout 200 python scripts/k1_time.py --config c2 --variants 5 --reps 10 --nodes 25000 | cut -c1-120

#!/bin/bash
# Round-2 first GPU call: tests + smoke + bench, the K1s build sweep, K3 timing, the C5 (N, P) sweep.
set -u
mkdir -p gpurun_out
bash scripts/gpu_check.sh quick 2>&1 | tail -40
echo "== K3"; timeout 300 python scripts/k3_time.py 2>&1 | tail -3 | tee gpurun_out/k3_time.txt
echo "== K1s sweep"; timeout 900 bash scripts/k1s_round2_sweep.sh > gpurun_out/k1s_sweep.log 2>&1; grep -E "^==|median" gpurun_out/k1s_round2_sweep.txt | cut -c1-150
echo "== C5"; timeout 600 python scripts/c5_time.py --nodes 100000,1000000,10000000 --pods 1,32,1000,10000 2>&1 | tee gpurun_out/c5_time.jsonl | cut -c1-220

#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== K3 v4"; timeout 300 python scripts/k3_time.py 2>&1 | tail -2 | tee gpurun_out/k3_time_v4.txt
echo "== K3 tests"; timeout 600 python -m pytest tests/test_place_sequential.py tests/test_memory_aware.py tests/test_state_changes.py tests/test_host_scheduler.py tests/test_cabi_sequence.py -x -q -m gpu 2>&1 | tail -3
echo "== stream (STREAM build, max_run 16)"; timeout 300 python scripts/c5_time.py --nodes 10000000 --pods 1,32,512 --stream-bytes 120 2>&1 | cut -c1-150 | tee gpurun_out/stream_prefetch.txt
echo "== churn"; timeout 200 python scripts/churn_time.py 2>&1 | tail -2
echo "== K3 ncu"; timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:place_sequential -s 1 -c 1 -o gpurun_out/r02_k3v4_c2 python scripts/k3_time.py > gpurun_out/ncu_k3v4.log 2>&1; tail -2 gpurun_out/ncu_k3v4.log
echo "== ncu stream P=32"; timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_stream_p32_v2 python scripts/c5_time.py --nodes 10000000 --pods 32 --reps 1 > gpurun_out/ncu_stream32v2.log 2>&1; tail -1 gpurun_out/ncu_stream32v2.log
echo "== shard 12500 work-list knobs"
for cfg in "25 3 8" "0 3 8" "15 3 8" "0 4 8" "0 2 8" "0 3 4" "25 2 8"; do set -- $cfg; echo "tail=$1 waves=$2 floor=$3"; KGPU_SP_TAIL=$1 KGPU_SP_WAVES=$2 KGPU_SP_FLOOR=$3 python scripts/k1_time.py --config c2 --nodes 12500 --variants 5 --reps 10 | cut -c1-90; done
echo "== full C2 knobs"
for cfg in "25 3 8" "15 3 8" "0 3 8" "15 4 8"; do set -- $cfg; echo "tail=$1 waves=$2 floor=$3"; KGPU_SP_TAIL=$1 KGPU_SP_WAVES=$2 KGPU_SP_FLOOR=$3 python scripts/k1_time.py --config c2 --variants 5 --reps 10 | cut -c1-90; done

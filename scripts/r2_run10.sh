#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== K3 v6"; timeout 200 python scripts/k3_time.py 2>&1 | tail -1 | tee gpurun_out/k3_time_v6.txt
echo "== K3 tests"; timeout 600 python -m pytest tests/test_place_sequential.py tests/test_memory_aware.py tests/test_state_changes.py tests/test_host_scheduler.py tests/test_cabi_sequence.py -x -q -m gpu 2>&1 | tail -3
echo "== K3 ncu"; timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:place_sequential -s 1 -c 1 -o gpurun_out/r02_k3_final python scripts/k3_time.py > gpurun_out/ncu_k3v6.log 2>&1; tail -1 gpurun_out/ncu_k3v6.log
echo "== K3 again"; timeout 200 python scripts/k3_time.py 2>&1 | tail -1

#!/bin/bash
# Runs on the B200 box under gpurun: GPU tests, smoke, a short bench and the ncu launch list.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick]'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  grep -c score_pairs gpurun_out/launches.csv
  echo "== ncu full (K1)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_pairs_sparse -s 3 -c 2 -f -o gpurun_out/k1_full \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi

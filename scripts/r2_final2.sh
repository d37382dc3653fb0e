#!/bin/bash
# Final evidence run of round 2, second session (1 GPU): tests, smoke, counters of the final kernel sources, bench line + reference arm,
# ncu launch list of the bench command, ncu --set full of the TMA few-pod kernel (32 pods, 1 pod), the few-pod rows of the C5 sweep.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== counters"; timeout 300 python scripts/refresh_counts.py 2>&1 | tail -1; cp gpurun_out/k1s_counts.json profiles/k1s_counts.json
echo "== bench N=1"; (time timeout 900 python bench.py 2>gpurun_out/bench1.err > gpurun_out/r02_bench_1gpu.json); echo rc=$?; tail -2 gpurun_out/bench1.err; cut -c1-300 gpurun_out/r02_bench_1gpu.json
echo "== bench reference arm"; (time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2>gpurun_out/benchref.err); cut -c1-300 gpurun_out/r02_bench_reference.json
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/bench_under_ncu.log 2>&1
grep -c score_pairs_sparse gpurun_out/r02_launches_bench.csv
NCU="ncu --set full --clock-control none --import-source on -f"
echo "== ncu TMA P=32"; timeout 600 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_tma_p32 python scripts/c5_time.py --nodes 10000000 --pods 32 --reps 1 > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log | cut -c1-200
echo "== ncu TMA P=1"; timeout 600 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_tma_p1 python scripts/c5_time.py --nodes 10000000 --pods 1 --reps 1 > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log | cut -c1-200
echo "== C5 few-pod rows"; timeout 600 python scripts/c5_time.py --nodes 10000,100000,1000000,10000000 --pods 1,32 --stream-bytes 120 2>&1 | cut -c1-260 | tee gpurun_out/r02_c5_fewpods_final.jsonl | cut -c1-120

#!/bin/bash
# Final evidence run of round 2 (1 GPU): tests, counters, bench line, ncu launch list of the bench command, ncu --set full captures.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== counters"; timeout 300 python scripts/refresh_counts.py 2>&1 | tail -1; cp gpurun_out/k1s_counts.json profiles/k1s_counts.json
echo "== bench N=1"; (time timeout 900 python bench.py 2>gpurun_out/bench1.err > gpurun_out/r02_bench_1gpu.json); echo rc=$?; tail -2 gpurun_out/bench1.err; cut -c1-400 gpurun_out/r02_bench_1gpu.json
echo "== bench reference arm"; (time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2>gpurun_out/benchref.err); cut -c1-300 gpurun_out/r02_bench_reference.json
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/bench_under_ncu.log 2>&1
grep -c score_pairs_sparse gpurun_out/r02_launches_bench.csv
NCU="ncu --set full --clock-control none --import-source on -f"
echo "== ncu K1s C2"; timeout 600 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_c2_final python scripts/k1_time.py --config c2 --variants 5 --reps 1 > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
echo "== ncu K3"; timeout 600 $NCU -k regex:place_sequential -s 1 -c 1 -o gpurun_out/r02_k3_final python scripts/k3_time.py > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
echo "== ncu stream P=32"; timeout 900 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_stream_p32_final python scripts/c5_time.py --nodes 10000000 --pods 32 --reps 1 > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
echo "== ncu stream P=1"; timeout 900 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_stream_p1_final python scripts/c5_time.py --nodes 10000000 --pods 1 --reps 1 > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
echo "== C5 (N, P) sweep"; timeout 600 python scripts/c5_time.py --nodes 10000,100000,1000000,10000000 --pods 1,32,10000 --stream-bytes 120 2>&1 | cut -c1-260 | tee gpurun_out/r02_c5_np_sweep_final.jsonl | cut -c1-120

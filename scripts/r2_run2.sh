#!/bin/bash
# Round-2 second GPU call: ncu --set full captures of K1s (C2, and the streaming regime N=1e7 P=32 / P=1) and of K3.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
echo "== K1s C2"
timeout 600 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_c2 python scripts/k1_time.py --config c2 --variants 5 --reps 1 > gpurun_out/ncu_k1s_c2.log 2>&1; tail -2 gpurun_out/ncu_k1s_c2.log
echo "== K3"
timeout 600 $NCU -k regex:place_sequential -s 1 -c 1 -o gpurun_out/r02_k3_c2 python scripts/k3_time.py > gpurun_out/ncu_k3.log 2>&1; tail -2 gpurun_out/ncu_k3.log
echo "== K1s stream N=1e7 P=32"
timeout 900 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_stream_p32 python scripts/c5_time.py --nodes 10000000 --pods 32 --reps 1 > gpurun_out/ncu_stream32.log 2>&1; tail -2 gpurun_out/ncu_stream32.log
echo "== K1s stream N=1e7 P=1"
timeout 900 $NCU -k regex:score_pairs_sparse -s 3 -c 1 -o gpurun_out/r02_k1s_stream_p1 python scripts/c5_time.py --nodes 10000000 --pods 1 --reps 1 > gpurun_out/ncu_stream1.log 2>&1; tail -2 gpurun_out/ncu_stream1.log
ls -la gpurun_out/*.ncu-rep

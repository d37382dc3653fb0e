#!/bin/bash
# K1s build-time sweep: min-blocks/SM and bucket-loop unroll; restores the default build at the end.
set -u
mkdir -p gpurun_out
: > gpurun_out/k1s_sweep.txt
for un in 1 2; do for mb in 5 6 8; do
  regs=$(make -s EXTRA="-DKGPU_SP_MINBLOCKS=$mb -DKGPU_SP_UNROLL=$un" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A2 'score_pairs_sparseILb1ELb0' | grep -E 'Used|spill' | sed 's/ptxas info    : //; s/, used 1 barriers.*//; s/bytes stack frame, //' | tr '\n' ' ')
  echo "unroll=$un minblocks=$mb :: $regs" | tee -a gpurun_out/k1s_sweep.txt
  python scripts/k1_time.py --config c2 --variants 5 --reps 8 | cut -c1-100 | tee -a gpurun_out/k1s_sweep.txt
  python scripts/k1_time.py --config c3 --variants 5 --reps 6 | cut -c1-100 | tee -a gpurun_out/k1s_sweep.txt
done; done
make -s -B kubegpu_b200/lib/libkgpu.so >/dev/null 2>&1

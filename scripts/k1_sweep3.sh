#!/bin/bash
# Sweep min-blocks/SM for the current generator settings; restores the default build at the end.
set -u
mkdir -p gpurun_out
for mb in ${MBS:-5 6 8}; do
  regs=$(make -s EXTRA="-DKGPU_LPN_MINBLOCKS=$mb" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A2 lane_per_node | grep -E 'Used' | sed 's/ptxas info    : //; s/, used 1 barriers.*//' | tr '\n' ' ')
  echo "minblocks=$mb :: $regs" | tee -a gpurun_out/sweep3.txt
  python scripts/k1_time.py --config c2 --variants 2,4 --reps 6 | sed "s/N=100000 P=10000  //; s/-> .*alg/alg/" | tee -a gpurun_out/sweep3.txt
  python scripts/k1_time.py --config c3 --variants 2,4 --reps 4 | sed "s/N=100000 P=10000  //; s/-> .*alg/alg/" | tee -a gpurun_out/sweep3.txt
done
make -s -B kubegpu_b200/lib/libkgpu.so >/dev/null 2>&1

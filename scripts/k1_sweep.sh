#!/bin/bash
# Build-time sweep of K1's min-blocks/SM (register budget) on the GPU box.
set -u
mkdir -p gpurun_out
for mb in ${@:-3 4 5}; do
  make -s EXTRA="-DKGPU_LPN_MINBLOCKS=$mb" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A1 lane_per_node | grep -E 'Used' 
  echo "minblocks=$mb" | tee -a gpurun_out/sweep.txt
  python scripts/k1_time.py --config c2 --variants 2 | tee -a gpurun_out/sweep.txt
  python scripts/k1_time.py --config c3 --variants 2 --reps 5 | tee -a gpurun_out/sweep.txt
done
make -s -B kubegpu_b200/lib/libkgpu.so >/dev/null 2>&1

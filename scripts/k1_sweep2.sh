#!/bin/bash
# Sweep generator options x min-blocks/SM on the GPU box; restores the default build at the end.
set -u
mkdir -p gpurun_out
cp kubegpu_b200/csrc/subset_dp_gen.cuh /tmp/subset_dp_gen.cuh.orig
for sy in 1 0; do
 for nacc in ${NACCS:-4}; do
  KGPU_GEN_STAGE_Y=$sy KGPU_GEN_NACC=$nacc python kubegpu_b200/csrc/gen_subset_dp.py > kubegpu_b200/csrc/subset_dp_gen.cuh
  for mb in ${MBS:-4 5 6 8}; do
    regs=$(make -s EXTRA="-DKGPU_LPN_MINBLOCKS=$mb" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A2 lane_per_node | grep -E 'Used|spill' | tr '\n' ' ')
    echo "stage_y=$sy nacc=$nacc minblocks=$mb :: $regs" | tee -a gpurun_out/sweep2.txt
    python scripts/k1_time.py --config c2 --variants 2,4 --reps 6 | tee -a gpurun_out/sweep2.txt
    python scripts/k1_time.py --config c3 --variants 2,4 --reps 4 | tee -a gpurun_out/sweep2.txt
  done
 done
done
cp /tmp/subset_dp_gen.cuh.orig kubegpu_b200/csrc/subset_dp_gen.cuh
make -s -B kubegpu_b200/lib/libkgpu.so >/dev/null 2>&1

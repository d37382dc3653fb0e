#!/bin/bash
# Generator-parameter sweep (K4/K3 form-B share, accumulators, min-blocks); restores default at the end.
set -u
mkdir -p gpurun_out
cp kubegpu_b200/csrc/subset_dp_gen.cuh /tmp/subset_dp_gen.cuh.orig
run() {  # $1 = label, env already set
  python kubegpu_b200/csrc/gen_subset_dp.py > kubegpu_b200/csrc/subset_dp_gen.cuh
  regs=$(make -s EXTRA="-DKGPU_LPN_MINBLOCKS=${MB:-6}" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A2 'lane_per_nodeILb1' | grep -E 'Used' | sed 's/ptxas info    : //; s/, used 1 barriers.*//' | tr '\n' ' ')
  echo "$1 mb=${MB:-6} :: $regs" | tee -a gpurun_out/sweep4.txt
  python scripts/k1_time.py --config c2 --variants 2 --reps 6 | sed "s/N=100000 P=10000  //; s/-> .*alg/alg/" | tee -a gpurun_out/sweep4.txt
  python scripts/k1_time.py --config c3 --variants 2 --reps 4 | sed "s/N=100000 P=10000  //; s/-> .*alg/alg/" | tee -a gpurun_out/sweep4.txt
}
for fb in 0 9 4 2 1; do KGPU_GEN_K4_FORMB=$fb run "k4formB=$fb"; done
for na in 2 8; do KGPU_GEN_NACC=$na run "nacc=$na"; done
for k3 in 0 2 1; do KGPU_GEN_K3_FORMB=$k3 run "k3formB=$k3"; done
cp /tmp/subset_dp_gen.cuh.orig kubegpu_b200/csrc/subset_dp_gen.cuh
make -s -B kubegpu_b200/lib/libkgpu.so >/dev/null 2>&1

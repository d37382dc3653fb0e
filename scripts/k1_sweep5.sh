#!/bin/bash
set -u
mkdir -p gpurun_out
cp kubegpu_b200/csrc/subset_dp_gen.cuh /tmp/subset_dp_gen.cuh.orig
run() {
  python kubegpu_b200/csrc/gen_subset_dp.py > kubegpu_b200/csrc/subset_dp_gen.cuh
  make -s EXTRA="-DKGPU_LPN_MINBLOCKS=${MB:-6}" -B kubegpu_b200/lib/libkgpu.so >/dev/null 2>&1
  echo "$1 mb=${MB:-6}" | tee -a gpurun_out/sweep5.txt
  python scripts/k1_time.py --config c2 --variants 2 --reps 8 | sed "s/N=100000 P=10000  //; s/-> .*alg/alg/" | tee -a gpurun_out/sweep5.txt
  python scripts/k1_time.py --config c3 --variants 2 --reps 4 | sed "s/N=100000 P=10000  //; s/-> .*alg/alg/" | tee -a gpurun_out/sweep5.txt
}
for na in 2 3 4; do for k3 in 0 3; do for k4 in 0 9; do
  KGPU_GEN_NACC=$na KGPU_GEN_K3_FORMB=$k3 KGPU_GEN_K4_FORMB=$k4 run "nacc=$na k3formB=$k3 k4formB=$k4"
done; done; done
cp /tmp/subset_dp_gen.cuh.orig kubegpu_b200/csrc/subset_dp_gen.cuh
make -s -B kubegpu_b200/lib/libkgpu.so >/dev/null 2>&1

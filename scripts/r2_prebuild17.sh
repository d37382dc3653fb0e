#!/bin/bash
# Variants timed by scripts/r2_run17.sh, built HERE (no GPU) into kubegpu_b200/lib/variants/ (git-ignored, travels with gpurun).
set -u
dir=kubegpu_b200/lib/variants
mkdir -p $dir
: > $dir/variants17.txt
build() {   # tag flags...
  tag=$1; shift
  flags="$*"
  out=$(make -s -B LIB=$dir/libkgpu_$tag.so EXTRA="$flags" $dir/libkgpu_$tag.so 2>&1)
  regs=$(echo "$out" | grep -A2 'score_pairs_sparseILb1ELb0ELb1ELb1ELb1' | grep -E 'Used|spill' | sed 's/ptxas info    : //; s/, used 1 barriers.*//; s/bytes stack frame, //' | tr '\n' ' ')
  echo "$tag | $flags | $regs" | tee -a $dir/variants17.txt
}
build k3helper    -DKGPU_PLACE_HELPER=1
build tma_s1_b8   -DKGPU_SP_TMA_STAGES=1 -DKGPU_SP_TMA_MINBLOCKS=8
build tma_s1_b7   -DKGPU_SP_TMA_STAGES=1 -DKGPU_SP_TMA_MINBLOCKS=7
build tma_s1_b6   -DKGPU_SP_TMA_STAGES=1 -DKGPU_SP_TMA_MINBLOCKS=6
build tma_s2_b5   -DKGPU_SP_TMA_STAGES=2 -DKGPU_SP_TMA_MINBLOCKS=5
build tma_s3_b4   -DKGPU_SP_TMA_STAGES=3 -DKGPU_SP_TMA_MINBLOCKS=4

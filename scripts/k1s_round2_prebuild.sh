#!/bin/bash
# Build the K1s variants of the round-2 sweep HERE (no GPU needed, ~40 s each) into
# kubegpu_b200/lib/variants/, so that the sweep on the GPU box only times them (the .so files are
# git-ignored but travel with gpurun).  Writes variants.txt: tag | flags | registers / spills.
set -u
dir=kubegpu_b200/lib/variants
mkdir -p $dir
: > $dir/variants.txt
build() {   # tag flags...
  tag=$1; shift
  flags="$*"
  out=$(make -s -B LIB=$dir/libkgpu_$tag.so EXTRA="$flags" $dir/libkgpu_$tag.so 2>&1)
  regs=$(echo "$out" | grep -A2 'score_pairs_sparseILb1ELb0ELb1' | grep -E 'Used|spill' | sed 's/ptxas info    : //; s/, used 1 barriers.*//; s/bytes stack frame, //' | tr '\n' ' ')
  echo "$tag | $flags | $regs" | tee -a $dir/variants.txt
}
build g2            -DKGPU_SP_GROUP=2
build g1            -DKGPU_SP_GROUP=1
build g4            -DKGPU_SP_GROUP=4
build g2_pf         -DKGPU_SP_GROUP=2 -DKGPU_SP_PREFETCH=1
build g4_pf         -DKGPU_SP_GROUP=4 -DKGPU_SP_PREFETCH=1
build g2_mb7        -DKGPU_SP_GROUP=2 -DKGPU_SP_MINBLOCKS=7
build g2_mb6        -DKGPU_SP_GROUP=2 -DKGPU_SP_MINBLOCKS=6
build g2_t256       -DKGPU_SP_GROUP=2 -DKGPU_SP_THREADS=256 -DKGPU_SP_MINBLOCKS=4

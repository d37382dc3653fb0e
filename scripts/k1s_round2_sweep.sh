#!/bin/bash
# Round-2 K1s sweep: times every prebuilt variant of kubegpu_b200/lib/variants/ (scripts/k1s_round2_prebuild.sh,
# run on the CPU box beforehand) on C2 and C3, then the work list on/off on an 8-GPU-sized shard.
# Falls back to building on the spot when the variants are missing.
set -u
mkdir -p gpurun_out
out=gpurun_out/k1s_round2_sweep.txt
: > $out
dir=kubegpu_b200/lib/variants
[ -f $dir/variants.txt ] || bash scripts/k1s_round2_prebuild.sh
while IFS='|' read -r tag flags regs; do
  tag=$(echo $tag); [ -z "$tag" ] && continue
  echo "== $tag :: $flags :: $regs" | tee -a $out
  python scripts/k1_time.py --lib $dir/libkgpu_$tag.so --config c2 --variants 5 --reps 8 | cut -c1-100 | tee -a $out
  python scripts/k1_time.py --lib $dir/libkgpu_$tag.so --config c3 --variants 5 --reps 6 | cut -c1-100 | tee -a $out
done < $dir/variants.txt
# small shard (what one of 8 GPUs holds of C2): plain grid against the work list
for wl in 0 1; do
  echo "== default build, 12500 nodes, KGPU_SP_WORKLIST=$wl" | tee -a $out
  KGPU_SP_WORKLIST=$wl python scripts/k1_time.py --config c2 --nodes 12500 --variants 5 --reps 10 | cut -c1-100 | tee -a $out
done

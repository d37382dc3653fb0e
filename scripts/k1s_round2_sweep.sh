#!/bin/bash
# Round-2 K1s sweep (pointer-walk pod table build): min-blocks/SM x small-loop unroll x "small" threshold.
# Prints registers/spills and the C2 / C3 kernel times per build; restores the default build at the end.
set -u
mkdir -p gpurun_out
out=gpurun_out/k1s_round2_sweep.txt
: > $out
for mb in 8 7 6; do for us in 1 2; do for sm in 6 15; do
  [ "$us" = 1 ] && [ "$sm" = 15 ] && continue
  flags="-DKGPU_SP_MINBLOCKS=$mb -DKGPU_SP_UNROLL_SMALL=$us -DKGPU_SP_SMALL=$sm"
  regs=$(make -s EXTRA="$flags" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A2 'score_pairs_sparseILb1ELb0ELb1' | grep -E 'Used|spill' | sed 's/ptxas info    : //; s/, used 1 barriers.*//; s/bytes stack frame, //' | tr '\n' ' ')
  echo "$flags :: $regs" | tee -a $out
  python scripts/k1_time.py --config c2 --variants 5 --reps 8 | cut -c1-100 | tee -a $out
  python scripts/k1_time.py --config c3 --variants 5 --reps 6 | cut -c1-100 | tee -a $out
done; done; done
# pods per trip of a bucket loop: 1, 2 (default), 4 (7.8k instructions of code: instruction cache?)
for g in 1 2 4; do
  flags="-DKGPU_SP_GROUP=$g"
  regs=$(make -s EXTRA="$flags" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A2 'score_pairs_sparseILb1ELb0ELb1' | grep -E 'Used|spill' | sed 's/ptxas info    : //; s/, used 1 barriers.*//; s/bytes stack frame, //' | tr '\n' ' ')
  echo "$flags :: $regs" | tee -a $out
  python scripts/k1_time.py --config c2 --variants 5 --reps 8 | cut -c1-100 | tee -a $out
  python scripts/k1_time.py --config c3 --variants 5 --reps 6 | cut -c1-100 | tee -a $out
done
# multipliers loaded one trip ahead (LDS latency off the per-trip chain), with 2 and 4 pods per trip
for g in 2 4; do
  flags="-DKGPU_SP_PREFETCH=1 -DKGPU_SP_GROUP=$g"
  regs=$(make -s EXTRA="$flags" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A2 'score_pairs_sparseILb1ELb0ELb1' | grep -E 'Used|spill' | sed 's/ptxas info    : //; s/, used 1 barriers.*//; s/bytes stack frame, //' | tr '\n' ' ')
  echo "$flags :: $regs" | tee -a $out
  python scripts/k1_time.py --config c2 --variants 5 --reps 8 | cut -c1-100 | tee -a $out
  python scripts/k1_time.py --config c3 --variants 5 --reps 6 | cut -c1-100 | tee -a $out
done
# 256-thread blocks: half as many per-pod block flushes and atomics
for us in 1 2; do
  flags="-DKGPU_SP_THREADS=256 -DKGPU_SP_MINBLOCKS=4 -DKGPU_SP_UNROLL_SMALL=$us"
  regs=$(make -s EXTRA="$flags" -B kubegpu_b200/lib/libkgpu.so 2>&1 | grep -A2 'score_pairs_sparseILb1ELb0ELb1' | grep -E 'Used|spill' | sed 's/ptxas info    : //; s/, used 1 barriers.*//; s/bytes stack frame, //' | tr '\n' ' ')
  echo "$flags :: $regs" | tee -a $out
  python scripts/k1_time.py --config c2 --variants 5 --reps 8 | cut -c1-100 | tee -a $out
  python scripts/k1_time.py --config c3 --variants 5 --reps 6 | cut -c1-100 | tee -a $out
done
make -s -B kubegpu_b200/lib/libkgpu.so >/dev/null 2>&1

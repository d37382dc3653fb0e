#!/bin/bash
set -u
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_state_changes.py tests/test_memory_aware.py -x -q -m gpu 2>&1 | tail -2
for pad in 32 128; do
  for n in 12500 25000 50000 100000; do echo -n "pad=$pad nodes=$n: "; KGPU_ORDER_PAD=$pad python scripts/k1_time.py --config c2 --nodes $n --variants 5 --reps 10 | cut -c20-75; done
  echo -n "pad=$pad c3: "; KGPU_ORDER_PAD=$pad python scripts/k1_time.py --config c3 --variants 5 --reps 6 | cut -c20-75
done

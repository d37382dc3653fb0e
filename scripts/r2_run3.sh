#!/bin/bash
# Round-2 third GPU call: tests after the node-state rewrite, memcheck of a small case, streaming sweep, C2 worklist.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== memcheck smoke"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/memcheck_smoke.txt
echo "== C5"; timeout 600 python scripts/c5_time.py --nodes 100000,1000000,10000000 --pods 1,32,1000,10000 --stream-bytes 120 2>&1 | tee gpurun_out/c5_time.jsonl | cut -c1-230
for wl in 0 1; do
  echo "== C2 KGPU_SP_WORKLIST=$wl"; KGPU_SP_WORKLIST=$wl python scripts/k1_time.py --config c2 --variants 5 --reps 10 | cut -c1-120
  echo "== C3 KGPU_SP_WORKLIST=$wl"; KGPU_SP_WORKLIST=$wl python scripts/k1_time.py --config c3 --variants 5 --reps 6 | cut -c1-120
done

#!/bin/bash
# N-GPU check of the final build (gpurun --gpus N): multi-device tests, bench under torchrun (push exchange) and its reference arm.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/bench_$N.err > gpurun_out/bench_$N.json
tail -3 gpurun_out/bench_$N.err; cut -c1-350 gpurun_out/bench_$N.json
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads([l for l in open('gpurun_out/bench_%s.json' % n) if l.startswith('{')][-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'e2e', 'gpu_launches', 'keys_sha256_12', 'parity_in_run', 'parity_all_subruns_ok')})
print('multi_device_handle', d.get('multi_device_handle'))
print('k1_ms', d.get('config', {}).get('k1_ms_max_over_ranks'), [k for k in d.keys()])
PY

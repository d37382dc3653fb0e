#!/bin/bash
# 8-GPU box, final build: the driver's N = 8 command, then the 1-GPU headline on the same box (efficiency denominator).
set -u
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 8 --steps 20 --warmup 3 2>gpurun_out/bench_8.err > gpurun_out/bench_8.json
tail -2 gpurun_out/bench_8.err | cut -c1-300
timeout 60 python bench.py --steps 20 --warmup 3 --no-variants --no-cpu-baseline 2>/dev/null > gpurun_out/bench_8box_1.json
python - <<'PY'
import json
for f in ('gpurun_out/bench_8.json', 'gpurun_out/bench_8box_1.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches', 'keys_sha256_12', 'parity_in_run', 'parity_all_subruns_ok')})
    print('  e2e', d.get('e2e', {}).get('value'), 'mdh', d.get('multi_device_handle'), 'c3', json.dumps(d.get('c3'))[:300])
PY

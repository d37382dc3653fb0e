#!/bin/bash
# Multi-GPU check on the box (gpurun --gpus N): host facts, multi-device handle test, torchrun bench.
set -u
N=${1:-2}
mkdir -p gpurun_out
{ echo "nproc=$(nproc)"; python -c "import os; print('affinity', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nvidia-smi -L; } | tee gpurun_out/host_facts.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_multi.txt
for n in $(seq 2 $N); do
  if [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/bench_$n.err | tee gpurun_out/bench_$n.json
  tail -3 gpurun_out/bench_$n.err
  fi
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-variants 2>/dev/null | tee gpurun_out/bench_1.json

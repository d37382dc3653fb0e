#!/bin/bash
# 2-GPU call: the store-based exchange (tests + bench push vs nccl), K3 v5.
set -u
mkdir -p gpurun_out
echo "== K3 v5"; timeout 200 python scripts/k3_time.py 2>&1 | tail -1 | tee gpurun_out/k3_time_v5.txt
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_place_sequential.py tests/test_memory_aware.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
run() { n=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 30 --warmup 5 "$@"; }
for ex in push nccl; do
  echo "== N=2 $ex"; run 2 --exchange $ex --no-variants --no-cpu-baseline 2>gpurun_out/b2_$ex.err > gpurun_out/b2_$ex.json; echo rc=$?; python -c "
import json; d=json.load(open('gpurun_out/b2_$ex.json')); print('ms/step %.4f kernel_ms %.4f value %.1fM e2e %.1fM' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value']/1e6, d['e2e']['value']/1e6))" || tail -5 gpurun_out/b2_$ex.err
done
echo "== K3 ncu"; timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:place_sequential -s 1 -c 1 -o gpurun_out/r02_k3v5_c2 python scripts/k3_time.py > gpurun_out/ncu_k3v5.log 2>&1; tail -1 gpurun_out/ncu_k3v5.log

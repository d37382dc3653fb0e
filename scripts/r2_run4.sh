#!/bin/bash
# Round-2 fourth GPU call (2 GPUs): tests incl. multi-GPU, counters, bench N=1, bench N=2 nccl vs push.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== counters"; timeout 300 python scripts/refresh_counts.py 2>&1 | tail -2
cp gpurun_out/k1s_counts.json profiles/k1s_counts.json 2>/dev/null
echo "== C5 small-P"; timeout 600 python scripts/c5_time.py --nodes 10000000 --pods 1,32,512 --stream-bytes 120 2>&1 | cut -c1-230 | tee gpurun_out/c5_smallp.jsonl
echo "== bench N=1"; (time timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench1.err > gpurun_out/bench1.json); tail -3 gpurun_out/bench1.err; cut -c1-1500 gpurun_out/bench1.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
echo "== bench N=2 nccl"; timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --exchange nccl --no-variants 2>gpurun_out/bench2n.err > gpurun_out/bench2_nccl.json; tail -3 gpurun_out/bench2n.err; cut -c1-700 gpurun_out/bench2_nccl.json
echo "== bench N=2 push"; timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --exchange push --no-variants 2>gpurun_out/bench2p.err > gpurun_out/bench2_push.json; echo rc=$?; tail -3 gpurun_out/bench2p.err; cut -c1-700 gpurun_out/bench2_push.json
echo "== bench N=2 full (auto)"; (time timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/bench2.err > gpurun_out/bench2.json); echo rc=$?; tail -3 gpurun_out/bench2.err; cut -c1-600 gpurun_out/bench2.json

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/refresh_counts.py 2>&1 | tail -1; cp gpurun_out/k1s_counts.json profiles/k1s_counts.json
(time timeout 600 python bench.py 2>gpurun_out/bench1.err > gpurun_out/r02_bench_1gpu.json); echo rc=$?; cut -c1-300 gpurun_out/r02_bench_1gpu.json

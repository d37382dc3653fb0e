#!/bin/bash
set -u
mkdir -p gpurun_out
for rep in 1 2; do for v in v4 pf0 pf1 pf2; do echo -n "$v: "; timeout 120 python scripts/k3_time.py kubegpu_b200/lib/variants/libkgpu_$v.so 2>&1 | tail -1; done; done | tee gpurun_out/k3_ab.txt

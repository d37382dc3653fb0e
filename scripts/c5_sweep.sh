#!/bin/bash
# BASELINE config 5: node-count sweep on one GPU (K1 headline variant), placements/s and algorithmic GB/s.
set -u
mkdir -p gpurun_out
: > gpurun_out/c5_sweep.txt
python scripts/k1_time.py --config c2 --nodes 10000 --pods 10000 --reps 6 | tee -a gpurun_out/c5_sweep.txt
python scripts/k1_time.py --config c2 --nodes 100000 --pods 10000 --reps 6 | tee -a gpurun_out/c5_sweep.txt
python scripts/k1_time.py --config c2 --nodes 1000000 --pods 10000 --reps 3 | tee -a gpurun_out/c5_sweep.txt
python scripts/k1_time.py --config c2 --nodes 10000000 --pods 1000 --reps 2 | tee -a gpurun_out/c5_sweep.txt

#!/bin/bash
# ncu --set full capture of K1 (current build) on config C2; prints key metrics.
set -u
mkdir -p gpurun_out
TAG=${1:-k1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-score_pairs_lane} -s ${SKIP:-3} -c 1 -f -o gpurun_out/$TAG \
    python scripts/k1_time.py --config ${2:-c2} --variants ${VARIANT:-2} --reps 2 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log

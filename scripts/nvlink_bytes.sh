#!/bin/bash
# NVLink bytes per step of the peer-memory key exchange (2 GPUs): driver NVLink counters before / after K steps of
# bench.py --exchange push (and, for comparison, --exchange nccl).  Writes gpurun_out/r02_nvlink_bytes.txt.
set -u
mkdir -p gpurun_out
out=gpurun_out/r02_nvlink_bytes.txt
: > $out
counters() { nvidia-smi nvlink -gt d -i 0 2>&1 | awk '/Data Tx/ {tx += $(NF-1)} /Data Rx/ {rx += $(NF-1)} END {print tx+0, rx+0}'; }
nvidia-smi nvlink -gt d -i 0 2>&1 | head -8 >> $out
for ex in push nccl; do
  K=4000
  read tx0 rx0 < <(counters)
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps $K --warmup 3 --exchange $ex --no-variants --no-cpu-baseline > gpurun_out/nvl_$ex.json 2>gpurun_out/nvl_$ex.err
  read tx1 rx1 < <(counters)
  python - <<PY | tee -a $out
import json
d = json.load(open("gpurun_out/nvl_$ex.json"))
steps = $K * 2 + 3 + 3 + 2      # timed steps of `value` and of the kernel-only pass do not both exchange: see below
tx, rx = ($tx1 - $tx0), ($rx1 - $rx0)
print("exchange=$ex  GPU0 NVLink counters over the whole run: tx %d KiB, rx %d KiB; timed steps %d (+ %d e2e steps + warm-up); ms/step %.4f" % (tx, rx, $K, $K, d["ms_per_step"]))
print("   per exchanging step (2 x %d + warm-ups): tx %.1f KiB, rx %.1f KiB; expected for the store exchange: 10000 keys x 8 B = 78.1 KiB tx and rx per rank at G = 2" % ($K, tx / (2.0 * $K + 8), rx / (2.0 * $K + 8)))
PY
done
cat $out

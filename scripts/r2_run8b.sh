#!/bin/bash
# 8-GPU validation of the store-based exchange + the scaling numbers of the final build (short).
set -u
mkdir -p gpurun_out
run() { n=$1; shift; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 30 --warmup 5 "$@"; }
echo "== N=8 full (auto)"; (time run 8 2>gpurun_out/b8_full.err > gpurun_out/b8_full.json); echo rc=$?; tail -2 gpurun_out/b8_full.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/b8_full.json"))
print("ms/step %.4f kernel_ms %.4f value %.1fM e2e %.1fM" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"] / 1e6, d["e2e"]["value"] / 1e6), d["config"]["parallelism"])
for k in ("parity_in_run", "multi_device_handle"):
    print(k, json.dumps(d.get(k))[:400])
print("c3 ms", d["c3"]["ms_per_step"], d["c3"]["parity"], "c5", [(p["nodes"], round(p["ms_per_step"], 4)) for p in d["c5"]["points"]])
PY
for n in 8 4 2; do echo "== N=$n push"; run $n --exchange push --no-variants --no-cpu-baseline 2>gpurun_out/b${n}_push.err > gpurun_out/b${n}_push.json; python -c "
import json; d=json.load(open('gpurun_out/b${n}_push.json')); print('ms/step %.4f kernel_ms %.4f value %.1fM' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value']/1e6))" || tail -3 gpurun_out/b${n}_push.err; done
echo "== N=8 nccl"; run 8 --exchange nccl --no-variants --no-cpu-baseline 2>gpurun_out/b8_nccl.err > gpurun_out/b8_nccl.json; python -c "
import json; d=json.load(open('gpurun_out/b8_nccl.json')); print('ms/step %.4f kernel_ms %.4f value %.1fM' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value']/1e6))"
echo "== N=1"; timeout 120 python bench.py --steps 30 --warmup 5 --no-variants --no-cpu-baseline 2>/dev/null > gpurun_out/b1.json; python -c "
import json; d=json.load(open('gpurun_out/b1.json')); print('ms/step %.4f value %.1fM' % (d['ms_per_step'], d['value']/1e6))"

#!/bin/bash
# 8-GPU call: multi-GPU tests (G = 2, 4, 8 handles, torchrun paths), bench at N = 8 with the three exchanges, at N = 4 / 2 push,
# and the full N = 8 line (c3 / c5 / multi_device_handle).
set -u
mkdir -p gpurun_out
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 30 --warmup 5 "$@"; }
for ex in nccl push allreduce; do
  echo "== N=8 $ex"; run 8 --exchange $ex --no-variants --no-cpu-baseline 2>gpurun_out/b8_$ex.err > gpurun_out/b8_$ex.json; echo rc=$?; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/b8_$ex.json"))
    print("ms/step %.4f  kernel_ms %.4f  value %.1fM  e2e %.1fM  parity %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d.get("parity_in_run")))
except Exception as e:
    print("no line:", e); print(open("gpurun_out/b8_$ex.err").read()[-1500:])
PY
done
echo "== N=8 nccl graph"; run 8 --exchange nccl --graph --no-variants --no-cpu-baseline 2>gpurun_out/b8_graph.err > gpurun_out/b8_graph.json; python -c "
import json; d=json.load(open('gpurun_out/b8_graph.json')); print('ms/step %.4f value %.1fM' % (d['ms_per_step'], d['value']/1e6))" || tail -5 gpurun_out/b8_graph.err
for n in 4 2; do echo "== N=$n push"; run $n --exchange push --no-variants --no-cpu-baseline 2>gpurun_out/b${n}_push.err > gpurun_out/b${n}_push.json; python -c "
import json; d=json.load(open('gpurun_out/b${n}_push.json')); print('ms/step %.4f kernel_ms %.4f value %.1fM' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value']/1e6))"; done
echo "== N=1"; timeout 300 python bench.py --steps 30 --warmup 5 --no-variants --no-cpu-baseline 2>/dev/null > gpurun_out/b1.json; python -c "
import json; d=json.load(open('gpurun_out/b1.json')); print('ms/step %.4f value %.1fM' % (d['ms_per_step'], d['value']/1e6))"
echo "== N=8 full"; (time run 8 2>gpurun_out/b8_full.err > gpurun_out/b8_full.json); echo rc=$?; tail -3 gpurun_out/b8_full.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/b8_full.json"))
for k in ("ms_per_step", "value", "parity_in_run", "c3", "multi_device_handle"):
    print(k, json.dumps(d.get(k))[:600])
print("c5", [(p["nodes"], round(p["ms_per_step"], 4)) for p in d["c5"]["points"]])
PY

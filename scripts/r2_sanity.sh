#!/bin/bash
# last check of the round's final tree (2 GPUs): all GPU tests, smoke, bench at N = 1 and N = 2 (full lines).
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench N=1"; (time timeout 600 python bench.py 2>gpurun_out/s1.err > gpurun_out/sanity_bench1.json); echo rc=$?; tail -2 gpurun_out/s1.err
echo "== bench N=2"; (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 2>gpurun_out/s2.err > gpurun_out/sanity_bench2.json); echo rc=$?; tail -2 gpurun_out/s2.err
python - <<'PY'
import json
for f in ("sanity_bench1", "sanity_bench2"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, "ms/step %.4f value %.1fM e2e %.1fM parity %s all_sub %s launches %s" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["parity_in_run"]["ok"], d["parity_all_subruns_ok"], d["gpu_launches"]))
    if "multi_device_handle" in d: print("  mdh", d["multi_device_handle"])
    if "hbm_regime" in d: print("  hbm", [(p["pods"], round(p["frac_of_hbm_peak"], 3)) for p in d["hbm_regime"]["points"]], "seq %.2f ms" % d["stateful_sequential"]["ms_per_batch"], "roof", round(d["roofline"]["frac"], 3), d["roofline"]["inst_matches_this_build"])
PY

#!/usr/bin/env python3
"""Run ON THE GPU BOX after a kernel change: one ncu pass (counters only, --clock-control none) over the headline
kernel on C2, written as gpurun_out/k1s_counts.json; copy it to profiles/k1s_counts.json (bench.py reads it for
roofline.achieved / roofline.traffic and says whether the kernel sources still match)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha
metrics = "smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed"
cmd = ["ncu", "--metrics", metrics, "--clock-control", "none", "-k", "regex:score_pairs_sparse", "-s", "3", "-c", "1", "--csv",
       sys.executable, os.path.join(ROOT, "scripts", "k1_time.py"), "--config", "c2", "--variants", "5", "--reps", "1"]
out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout
rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 5]
hdr = next(r for r in rows if "Metric Name" in r)
mi, vi, ui = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
vals = {}
for r in rows:
    if r is hdr or len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    unit = r[ui]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    vals[r[mi]] = v * scale
res = {"kernel_source_sha": kernel_source_sha(), "workload": "C2, 1 GPU, one launch of score_pairs_sparse<true,false,true>",
       "inst_per_launch_c2": vals.get("smsp__inst_executed.sum"),
       "dram_bytes_per_launch_c2": (vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)),
       "ncu_duration_ns": vals.get("gpu__time_duration.sum"),
       "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "sm_throughput_pct_elapsed": vals.get("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
       "how": "ncu --metrics ... --clock-control none -k regex:score_pairs_sparse -s 3 -c 1 python scripts/k1_time.py --config c2 --variants 5"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "k1s_counts.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))

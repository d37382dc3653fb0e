#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py rep [rep...]"""
import csv, io, subprocess, sys
WANT = [
 "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
 "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
 "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
 "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
 "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
 "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
 "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
]
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("==", rep, "(%d launches)" % len(data))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("  %-82s %-10s %s" % (w, units[i], " | ".join(d[i] for d in data)))

#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: the metrics DESIGN.md / profiles/ quote."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active",
        "smsp__issue_active.avg.pct", "smsp__inst_issued.avg.per_cycle_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp", "smsp__warps_eligible.avg.per_cycle_active", "smsp__pcsamp_warps_issue_stalled",
        "sm__cycles_elapsed.avg ", "smsp__cycles_active.avg", "sm__inst_executed_pipe_", "smsp__inst_executed_op_",
        "smsp__sass_thread_inst_executed_op_d", "sm__throughput.avg.pct", "gpu__dram_throughput",
        "sm__sass_thread_inst_executed_op_dfma", "sm__sass_thread_inst_executed_op_dadd", "sm__sass_thread_inst_executed_op_dmul",
        "smsp__sass_average_branch_targets", "derived__smsp__sass_thread_inst_executed_op", "local"]
for vals in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k) or (k in h and k.endswith("_")) for k in keys):
            try:
                fv = float(v.replace(",", ""))
                if fv == 0 and "pcsamp" in h:
                    continue
            except ValueError:
                pass
            print(f"{h:95s} {u:14s} {v}")

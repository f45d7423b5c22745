#!/usr/bin/env python
"""print the key metrics of an ncu report (first kernel): time, occupancy limiters, issue/fp64 utilisation, stall mix"""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k]}")
st = {k: float(v) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "average_warps" in k}
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]:
    print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:22s} {v:.2f}")

#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per launch duration, issue
utilisation, FP64 pipe utilisation, stall reasons, DRAM traffic."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
         and "not_issued" not in h]
for r in rows[2:]:
    print("=" * 100)
    print(r[idx["Kernel Name"]], r[idx.get("Grid Size", 0)] if "Grid Size" in idx else "")
    for k in keys:
        if k in idx:
            print(f"  {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")
    st = sorted(((float(r[idx[h]].replace(',', '')), h) for h in stall if r[idx[h]]), reverse=True)
    print("  stall reasons (warps stalled per issue-active cycle):")
    for v, h in st[:8]:
        print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:8.3f}")

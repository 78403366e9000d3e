#!/usr/bin/env python
"""Print the metrics the profiles/ summaries quote from an ncu report:  python tools/ncu_pick.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
print(f"{'Kernel Name':84s} {vals[col['Kernel Name']]}")
for k in KEEP:
    if k in col:
        print(f"{k:84s} {vals[col[k]]:>18s} {units[col[k]]}")

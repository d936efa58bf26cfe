#!/usr/bin/env python
"""Key numbers of every kernel in an .ncu-rep (needs ncu on PATH):  python tools/ncu_summary.py prof.ncu-rep"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = {"Kernel Name": "kernel", "gpu__time_duration.sum": "us", "launch__registers_per_thread": "regs", "launch__grid_size": "grid",
        "launch__block_size": "block", "smsp__inst_executed.sum": "warp_inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma%",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu%",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu%", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_tp%", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_conf",
        "smsp__average_warp_latency_per_inst_issued.ratio": "lat/inst"}
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    out = []
    for k, name in want.items():
        if k in idx:
            v = r[idx[k]]
            if name == "kernel":
                v = v[:48]
            out.append(f"{name}={v} {rows[1][idx[k]] if name in ('us', 'dram_rd', 'dram_wr') else ''}".strip())
    print(" | ".join(out))

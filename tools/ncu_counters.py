#!/usr/bin/env python
"""Profile constants of one kernel from an `ncu --set full` report -> JSON (what bench.py quotes as roofline.traffic / issue_frac):
    python tools/ncu_counters.py gpurun_out/prof.ncu-rep <kernel-substring> "<how it was captured>" > profiles/r2_kernel_counters.json"""
import csv
import json
import subprocess
import sys

rep, pat, how = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
row = next(r for r in rows[2:] if pat in r[hdr.index("Kernel Name")])
val = lambda name: float(row[hdr.index(name)].replace(",", ""))
unit = lambda name: units[hdr.index(name)]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {
    "kernel": row[hdr.index("Kernel Name")],
    "dram_bytes_read": val("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")],
    "dram_bytes_write": val("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")],
    "issue_slots_busy_pct": val("sm__inst_issued.avg.pct_of_peak_sustained_active") if "sm__inst_issued.avg.pct_of_peak_sustained_active" in hdr else val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warp_instructions": val("smsp__inst_executed.sum"),
    "duration_us_under_ncu": val("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit("gpu__time_duration.sum"), 1),
    "registers_per_thread": val("launch__registers_per_thread"),
    "pipe_fma_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "pipe_fmaheavy_pct": val("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "pipe_alu_pct": val("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    "pipe_xu_pct": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    "source": how,
}
print(json.dumps(out, indent=1))

#!/usr/bin/env python
"""Warp-stall sample totals of one kernel from an `ncu --page source --csv` dump (optionally restricted to an address range).
    python tools/ncu_stalls.py sass.csv [lo_hex hi_hex]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 62
cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot = collections.Counter()
base = None
n_inst = 0
ci = hdr.index("Instructions Executed")
for r in rows[h + 1:]:
    if len(r) != len(hdr):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    off = addr - base
    if not (lo <= off < hi):
        continue
    n_inst += int(r[ci] or 0)
    for i in cols:
        tot[hdr[i]] += int(r[i] or 0)
s = sum(tot.values())
print(f"range [{lo:#x},{hi:#x}) warp-instructions {n_inst}, samples {s}")
for k, v in tot.most_common():
    if v:
        print(f"  {k:28s} {v:7d} {100*v/s:5.1f}%")

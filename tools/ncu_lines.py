#!/usr/bin/env python
"""Executed warp-instructions and stall samples per CUDA source line from
    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
Usage: python tools/ncu_lines.py src.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr = None, None
agg, samp, src = collections.Counter(), collections.Counter(), {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and cur and len(r) >= len(hdr) - 2 and r[0].isdigit():
        try:
            n, s = int(r[ci] or 0), int(r[cs] or 0)
        except ValueError:
            continue
        k = (cur, int(r[0]))
        agg[k] += n
        samp[k] += s
        src[k] = r[1].strip()[:100]
tot, ts = sum(agg.values()), sum(samp.values())
print(f"warp-instructions {tot}, samples {ts}")
byf, sbyf = collections.Counter(), collections.Counter()
for k, n in agg.items():
    byf[k[0]] += n
    sbyf[k[0]] += samp[k]
print("by file:", {k: f"{100 * v / tot:.1f}% i / {100 * sbyf[k] / ts:.1f}% s" for k, v in byf.most_common()})
for k, n in agg.most_common(top):
    print(f"{100 * n / tot:5.2f}% i {100 * samp[k] / ts:5.2f}% s {k[0]}:{k[1]}  {src[k]}")

#!/usr/bin/env python
"""Warp-stall samples per executed instruction, by opcode: how long a warp sits in front of each kind of instruction.
    python tools/ncu_cost_by_op.py sass.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci, cs, csrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = {c: i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c}
n, s = collections.Counter(), collections.Counter()
why = collections.defaultdict(collections.Counter)
for r in rows[h + 1:]:
    if len(r) != len(hdr):
        continue
    op = re.sub(r"^@!?U?P\w+\s+", "", r[csrc].strip()).split()[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "IMAD")) else op.split(".")[0]
    n[op] += int(r[ci] or 0)
    s[op] += int(r[cs] or 0)
    for c, i in stall_cols.items():
        why[op][c[6:]] += int(r[i] or 0)
tn, ts = sum(n.values()), sum(s.values())
print(f"{'op':14s} {'inst%':>6s} {'samp%':>6s} {'ratio':>6s}  top stall reasons")
for op, v in s.most_common(22):
    top = ", ".join(f"{k} {100*c/max(v,1):.0f}%" for k, c in why[op].most_common(4))
    print(f"{op:14s} {100*n[op]/tn:6.1f} {100*v/ts:6.1f} {(v/ts)/(n[op]/tn) if n[op] else 0:6.2f}  {top}")

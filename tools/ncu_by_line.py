#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with `nvdisasm -g -c` line info: executed warp-instructions and stall
samples per source line / per file / per opcode.  Usage:
    ncu -i prof.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g -c x.cubin > all.dis   (cut the kernel's .text section)
    python tools/ncu_by_line.py sass.csv kernel.dis [top_n]
"""
import collections
import csv
import re
import sys

sass_csv, dis, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci, cs, csrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
ins = [(r[csrc].strip(), int(r[ci] or 0), int(r[cs] or 0)) for r in rows[h + 1:] if len(r) == len(hdr)]

loc, locs = None, []
for line in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        locs.append(loc)
assert len(locs) == len(ins), (len(locs), len(ins))
by_line, by_file, by_op = collections.Counter(), collections.Counter(), collections.Counter()
s_line = collections.Counter()
tot = sum(i[1] for i in ins)
tots = sum(i[2] for i in ins)
for (src, n, s), l in zip(ins, locs):
    key = (l[0], l[1]) if l else ("?", 0)
    by_line[key] += n
    s_line[key] += s
    by_file[key[0]] += n
    op = re.sub(r"^@!?U?P\w+\s+", "", src).split()[0].split(".")[0]
    by_op[op] += n
print(f"total warp-instructions {tot}, samples {tots}")
print("by file:", {k: f"{100 * v / tot:.1f}%" for k, v in by_file.most_common()})
print("by opcode:", {k: f"{100 * v / tot:.1f}%" for k, v in by_op.most_common(24)})
for (f, l), n in by_line.most_common(top):
    print(f"{100 * n / tot:6.2f}% inst {100 * s_line[(f, l)] / max(tots, 1):6.2f}% samples  {f}:{l}")

#!/usr/bin/env python
"""Where does a kernel spill?  Counts STL/LDL (local-memory stores/loads) per source line from nvdisasm line info.
    python tools/spills_by_line.py <object-or-so> <kernel-name-substring> [top_n]"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

obj, pat, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    cur, loc, per, total, ops = None, None, collections.Counter(), 0, collections.Counter()
    for line in dis.split("\n"):
        m = re.match(r"//-+ \.text\.(\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or pat not in cur:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            loc = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            total += 1
            op = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?", "", line).split()[0].split(".")[0]
            ops[op] += 1
            if op in ("STL", "LDL"):
                per[(loc, op)] += 1
    if total:
        print(f"{os.path.basename(cub)}: {total} instructions matching '{pat}'", dict(ops.most_common(12)))
        for (l, op), n in per.most_common(top):
            print(f"  {n:5d} {op} {l}")

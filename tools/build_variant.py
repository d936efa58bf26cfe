#!/usr/bin/env python
"""Side build of libofdmgan with different -D flags for ONE translation unit (default csrc/sim_ws.cu): the other objects are the
default build's.  Result: ofdm-gan-sr_b200/lib/libofdmgan_<name>.so, selected at run time with OFDMGAN_LIB=<path>.
    python tools/build_ws_variant.py <name> [-DOG_WS_CWG=3 ...] [--unit sim_ws.cu]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ofdm-gan-sr_b200")
name, flags, unit = sys.argv[1], [], "sim_lean.cu"
args = sys.argv[2:]
while args:
    a = args.pop(0)
    if a == "--unit":
        unit = args.pop(0)
    else:
        flags.append(a)
arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
obj_dir = os.path.join(PKG, "csrc", "_obj")
var_obj = os.path.join(obj_dir, unit[:-3] + "_" + name + ".var.o")
cmd = ["nvcc"] + arch + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + flags + ["-c", os.path.join(PKG, "csrc", unit), "-o", var_obj]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr)
for line in r.stderr.split("\n"):
    if "k_sim_lean" in line or "spill" in line or "Used" in line:
        print(line)
objs = [os.path.join(obj_dir, f) for f in sorted(os.listdir(obj_dir)) if f.endswith(".o") and not f.endswith(".var.o") and f != unit[:-3] + ".o" and f != "train.o"]
lib = os.path.join(PKG, "lib", "libofdmgan_%s.so" % name)
r = subprocess.run(["nvcc"] + arch + ["-shared", "-o", lib] + objs + [var_obj], capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr)
print(lib)

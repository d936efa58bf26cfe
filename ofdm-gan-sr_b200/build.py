"""Build libofdmgan.so (hand-written sm_100a CUDA behind the C ABI of include/ofdmgan.h) with nvcc, in-tree.

    python ofdm-gan-sr_b200/build.py [--force]

Each .cu is one translation unit (its own __constant__ bank); objects are cached under csrc/_obj by source mtime.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment hooks: OG_NVCC_FLAGS adds compiler flags (e.g. -DOG_SIM_MINB=4), OG_VARIANT names a side build that leaves
# the default library untouched (lib/libofdmgan_<variant>.so, selected at run time with OFDMGAN_LIB=<path>)
VARIANT = os.environ.get("OG_VARIANT", "")
OBJ = os.path.join(CSRC, "_obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "lib", "libofdmgan" + ("_" + VARIANT if VARIANT else "") + ".so")
UNITS = ["runtime.cu", "infer.cu", "sim_gauss.cu", "sim_gauss_ext.cu", "sim_qpsk.cu", "sim_lean.cu", "critic_step.cu", "critic_api.cu", "gen_train.cu", "ofdm_api.cu", "critic_q.cu", "peer_comm.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"] + os.environ.get("OG_NVCC_FLAGS", "").split()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "ofdmgan.h"))
    return hs


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(p) > t for p in [src] + _headers())


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    units = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u))]
    jobs = []
    for u in units:
        src, obj = os.path.join(CSRC, u), os.path.join(OBJ, u[:-3] + ".o")
        if force or _stale(src, obj):
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        cmd = ["nvcc"] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for log in ex.map(cc, jobs):
                if verbose and log:
                    sys.stderr.write(log)
    objs = [os.path.join(OBJ, u[:-3] + ".o") for u in units]
    if jobs or not os.path.exists(LIB):
        r = subprocess.run(["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

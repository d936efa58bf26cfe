#!/usr/bin/env python3
"""Extract fixed-point critic known-answer data from the reference's committed RTL simulation.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_rtl_critic_vectors.py

Inputs (read-only):
  /root/reference/rtl/ofdmGAN/tb_discriminator_mini.vcd   Icarus dump of tb_discriminator_mini.v
  (the ROM literals are already in tests/golden/rtl_generator_vectors.json, from weight_rom.v)

Output: tests/golden/rtl_critic_vectors.json
  vectors[i].candidate / .condition : 32 int16 each, order ch0[0..15], ch1[0..15]  (discriminator_mini.v:233-256)
  vectors[i].score                  : score_out when score_valid                     (discriminator_mini.v:488-497)
  vectors[i].dense_acc              : the accumulator that was saturated into the score
  vectors[i].trace                  : every (state, out_ch, out_pos, last_in_ch, ksum) presented to the accumulate
                                      stage in CONV1 / CONV2 / DENSE, in cycle order - pins the cycle-level emulator
                                      far more tightly than the five (degenerate: the committed ROM is almost all
                                      zeros) scores do.
Sampling rule as in make_rtl_vectors.py: the DUT sees the values in effect before the timestamp of the clk edge.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_rtl_vectors import s  # noqa: E402

REF = "/root/reference/rtl/ofdmGAN"
HERE = os.path.dirname(os.path.abspath(__file__))
TB = "tb_discriminator_mini"
WANT = ["clk", "cand_in", "cand_valid", "cond_in", "cond_valid", "test_num", "score_out", "score_valid",
        "dut.state", "dut.pipe_s3_valid", "dut.pipe_s3_out_ch", "dut.pipe_s3_out_pos", "dut.pipe_s3_last_in_ch",
        "dut.pipe_s3_ksum", "dut.dense_acc"]


def snapshots(path):
    """Yield the pre-edge signal values at every rising clk edge."""
    ids, scope, cur = {}, [], {}
    want = {TB + "." + w for w in WANT}
    f = open(path)
    for line in f:
        t = line.split()
        if not t:
            continue
        if t[0] == "$scope":
            scope.append(t[2])
        elif t[0] == "$upscope":
            scope.pop()
        elif t[0] == "$var":
            name = ".".join(scope + [t[4]])
            if name in want:
                ids.setdefault(t[3], []).append(name[len(TB) + 1:])
        elif t[0] == "$enddefinitions":
            break
    pending, clk_prev = {}, 0
    for line in f:
        line = line.strip()
        if not line or line[0] == "$":
            continue
        if line[0] == "#":
            if pending:
                if pending.get("clk", clk_prev) == 1 and clk_prev == 0:
                    yield dict(cur)
                cur.update(pending)
                clk_prev = cur.get("clk", 0)
                pending = {}
            continue
        if line[0] in "01xz":
            val, vid = line[0], line[1:]
        elif line[0] == "b":
            val, vid = line[1:].split()
        else:
            continue
        if vid in ids:
            v = 0 if ("x" in val or "z" in val) else int(val, 2)
            for n in ids[vid]:
                pending[n] = v


def main():
    ST_LOAD_CAND, ST_LOAD_COND, ST_CONV1, ST_CONV2, ST_DENSE, ST_OUTPUT = 1, 2, 3, 4, 6, 7
    runs, cur = [], None
    for snap in snapshots(os.path.join(REF, TB + ".vcd")):
        st = snap.get("dut.state", 0)
        if st == ST_LOAD_CAND and snap.get("cand_valid", 0):
            if cur is None or len(cur["candidate"]) == 32:
                cur = {"test": snap.get("test_num", 0), "candidate": [], "condition": [], "trace": [], "score": None,
                       "dense_acc": None}
                runs.append(cur)
            cur["candidate"].append(s(snap["cand_in"], 16))
        elif st == ST_LOAD_COND and snap.get("cond_valid", 0):
            cur["condition"].append(s(snap["cond_in"], 16))
        elif st in (ST_CONV1, ST_CONV2, ST_DENSE) and snap.get("dut.pipe_s3_valid", 0):
            cur["trace"].append([st, snap["dut.pipe_s3_out_ch"], snap["dut.pipe_s3_out_pos"],
                                 snap["dut.pipe_s3_last_in_ch"], s(snap["dut.pipe_s3_ksum"], 32)])
        elif st == ST_OUTPUT:
            cur["dense_acc"] = s(snap["dut.dense_acc"], 32)
        if snap.get("score_valid", 0) and cur is not None and cur["score"] is None:
            cur["score"] = s(snap["score_out"], 16)
    for r in runs:
        assert len(r["candidate"]) == 32 and len(r["condition"]) == 32 and r["score"] is not None, r["test"]
    out = {"source": "rtl/ofdmGAN/tb_discriminator_mini.vcd (reference, committed Icarus run)", "vectors": runs}
    path = os.path.join(HERE, "rtl_critic_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, "runs:", len(runs), "bytes:", os.path.getsize(path))
    for r in runs:
        print("test", r["test"], "score", r["score"], "dense_acc", r["dense_acc"], "trace", len(r["trace"]),
              "cand[:6]", r["candidate"][:6], "cond[:6]", r["condition"][:6])


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""Extract fixed-point generator known-answer vectors from the reference's committed RTL simulation.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_rtl_vectors.py

Inputs (read-only):
  /root/reference/rtl/ofdmGAN/tb_generator_mini.vcd   Icarus dump of tb_generator_mini.v (10 tests)
  /root/reference/rtl/ofdmGAN/weight_rom.v            ROM literals (weights[] :42-159, biases[] :190-258)

Output: tests/golden/rtl_generator_vectors.json
  rom.weights / rom.biases : {address: signed value} for every literal assignment in weight_rom.v
  vectors[i].input/output  : 32 int16 each, order I[0..15], Q[0..15]
  vectors[i].total_cycles  : the testbench's total_cycles
  vectors[i].trace         : (tests 4 and 7 only) every (state, out_ch, out_pos, last_in_ch, ksum) presented
                             to the accumulate stage, in cycle order - pins the cycle-level emulator's internals.

Sampling rule: the DUT reads its inputs at posedge clk, i.e. it sees the values that were in effect *before*
the timestamp of the edge.  data_in is taken when state==LOAD_IN and valid_in (generator_mini.v:266-267),
data_out when valid_out && ready_out (tb_generator_mini.v:516-519).  The loaded frame is cross-checked against
the data_k0/1/2 fetch registers seen during ST_ENC1 (generator_mini.v:331-333).
"""
import json
import os
import re
import sys

REF = "/root/reference/rtl/ofdmGAN"
HERE = os.path.dirname(os.path.abspath(__file__))


def s(v, bits):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def parse_rom(path):
    w, b = {}, {}
    for line in open(path):
        m = re.search(r"weights\[(\d+)\]\s*=\s*8'h([0-9A-Fa-f]+)", line)
        if m:
            w[int(m.group(1))] = s(int(m.group(2), 16), 8)
        m = re.search(r"biases\[(\d+)\]\s*=\s*16'h([0-9A-Fa-f]+)", line)
        if m:
            b[int(m.group(1))] = s(int(m.group(2), 16), 16)
    return w, b


def parse_vcd(path):
    """Yield (pre_edge_values: dict name->int) for every rising clk edge."""
    ids = {}          # vcd id -> list of names
    scope = []
    want = {
        "tb_generator_mini.clk", "tb_generator_mini.data_in", "tb_generator_mini.valid_in",
        "tb_generator_mini.ready_out", "tb_generator_mini.valid_out", "tb_generator_mini.data_out",
        "tb_generator_mini.test_num", "tb_generator_mini.total_cycles",
        "tb_generator_mini.dut.state", "tb_generator_mini.dut.in_ch_cnt", "tb_generator_mini.dut.in_pos_cnt",
        "tb_generator_mini.dut.data_k0", "tb_generator_mini.dut.data_k1", "tb_generator_mini.dut.data_k2",
        "tb_generator_mini.dut.pipe_s3_valid", "tb_generator_mini.dut.pipe_s3_out_ch",
        "tb_generator_mini.dut.pipe_s3_out_pos", "tb_generator_mini.dut.pipe_s3_last_in_ch",
        "tb_generator_mini.dut.pipe_s3_ksum", "tb_generator_mini.dut.out_ch_cnt",
        "tb_generator_mini.dut.out_pos_cnt", "tb_generator_mini.dut.in_ch_iter",
        "tb_generator_mini.dut.weight_addr_base", "tb_generator_mini.dut.pipe_flush",
    }
    cur = {}
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
                ids.setdefault(t[3], []).append(name.replace("tb_generator_mini.", ""))
        elif t[0] == "$enddefinitions":
            break
    pending = {}
    clk_prev = 0
    for line in f:
        line = line.strip()
        if not line or line[0] == "$":
            continue
        if line[0] == "#":
            # apply pending changes of the previous timestamp; emit a snapshot if clk rose in them
            if pending:
                rose = pending.get("clk", clk_prev) == 1 and clk_prev == 0
                if rose:
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
            continue  # reals
        if vid not in ids:
            continue
        v = 0 if ("x" in val or "z" in val) else int(val, 2)
        for n in ids[vid]:
            pending[n] = v


def main():
    w, b = parse_rom(os.path.join(REF, "weight_rom.v"))
    tests = {}
    order = []
    ST_LOAD, ST_ENC1 = 1, 2
    for snap in parse_vcd(os.path.join(REF, "tb_generator_mini.vcd")):
        tn = snap.get("test_num", 0)
        if tn == 0:
            continue
        if tn not in tests:
            tests[tn] = {"input": [], "output": [], "enc_fetch": {}, "trace": [], "total_cycles": None}
            order.append(tn)
        T = tests[tn]
        st = snap.get("dut.state", 0)
        if st == ST_LOAD and snap.get("valid_in", 0):
            T["input"].append(s(snap["data_in"], 16))
        if snap.get("valid_out", 0) and snap.get("ready_out", 0) and len(T["output"]) < 32:
            T["output"].append(s(snap["data_out"], 16))
        if snap.get("dut.pipe_s3_valid", 0) and 2 <= st <= 8:
            T["trace"].append([st, snap["dut.pipe_s3_out_ch"], snap["dut.pipe_s3_out_pos"],
                               snap["dut.pipe_s3_last_in_ch"], s(snap["dut.pipe_s3_ksum"], 32)])
    # total_cycles: take the final value per test from a second pass (it is assigned after done)
    last_tc = {}
    for snap in parse_vcd(os.path.join(REF, "tb_generator_mini.vcd")):
        tn = snap.get("test_num", 0)
        if tn:
            last_tc[tn] = s(snap.get("total_cycles", 0), 32)
    vectors = []
    for tn in order:
        T = tests[tn]
        assert len(T["input"]) == 32, (tn, len(T["input"]))
        assert len(T["output"]) == 32, (tn, len(T["output"]))
        v = {"test": tn, "input": T["input"], "output": T["output"]}
        # total_cycles of test k is visible once test k+1 has started (or at the end for the last one)
        v["total_cycles"] = last_tc.get(tn + 1, last_tc[tn]) if tn + 1 in last_tc else last_tc[tn]
        if tn in (4, 7):
            v["trace"] = T["trace"]
        vectors.append(v)
    out = {
        "source": "rtl/ofdmGAN/tb_generator_mini.vcd + weight_rom.v (reference, Icarus run of 2025-12-19)",
        "rom": {"weights": {str(k): v for k, v in sorted(w.items())},
                "biases": {str(k): v for k, v in sorted(b.items())}},
        "vectors": vectors,
    }
    path = os.path.join(HERE, "rtl_generator_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, "tests:", order, "bytes:", os.path.getsize(path))
    for v in vectors:
        print("T%d in :" % v["test"], v["input"])
        print("T%d out:" % v["test"], v["output"], "cycles", v["total_cycles"])


if __name__ == "__main__":
    sys.exit(main())

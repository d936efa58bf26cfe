"""Oracle pinning, fixed-point generator: closed form vs the reference's RTL known answers and vs the
cycle-level emulation of generator_mini.v (CPU only)."""
import numpy as np

import oracle
from oracle.rtl_cycle_emulator import GeneratorMiniRTL


def _rom(rtl_vectors):
    return oracle.rom_arrays(rtl_vectors["rom"]["weights"], rtl_vectors["rom"]["biases"])


def test_rom_literals(rtl_vectors):
    W, Bq = _rom(rtl_vectors)
    # rtl/ofdmGAN/weight_rom.v:59-61,110-117,210-213
    assert list(W[0:3]) == [0x20, 0x40, 0x20] and list(W[216:224]) == [0x50, 0x40, 0x48, 0x38, 0x58, 0x30, 0x44, 0x3C]
    assert list(Bq[0:4]) == [16, -16, 8, -8] and Bq[16] == 0
    assert np.count_nonzero(W[24:120]) == 8 and np.count_nonzero(W[120:216]) == 6


def test_rtl_literal_matches_vcd_known_answers(rtl_vectors):
    W, Bq = _rom(rtl_vectors)
    x = np.array([v["input"] for v in rtl_vectors["vectors"]], np.int16)
    y = np.array([v["output"] for v in rtl_vectors["vectors"]], np.int16)
    assert len(x) == 10
    got = oracle.gen_fwd_q(x, W, Bq, mode=1).reshape(-1, 32)
    assert np.array_equal(got, y)
    assert all(v["total_cycles"] == 1055 for v in rtl_vectors["vectors"])


def test_cycle_emulator_matches_vcd(rtl_vectors):
    rtl = GeneratorMiniRTL(rtl_vectors["rom"]["weights"], rtl_vectors["rom"]["biases"])
    for v in rtl_vectors["vectors"]:
        rtl.trace = []
        out, cycles = rtl.run_frame(v["input"])
        assert out == v["output"], v["test"]
        assert abs(cycles - v["total_cycles"]) <= 2          # testbench handshake adds two cycles
        if "trace" in v:                                      # accumulate-stage inputs, cycle by cycle
            assert rtl.trace == [t for t in v["trace"] if t[0] in (2, 3, 5, 8)]


def test_closed_form_equals_cycle_emulator_on_random_roms():
    rng = np.random.default_rng(0)
    for trial in range(6):
        W = rng.integers(-128, 128, 2048).astype(np.int8)
        Bq = np.zeros(64, np.int16)
        Bq[:18] = rng.integers(-2000, 2000, 18)
        scale = [300, 3000, 32767][trial % 3]
        frames = rng.integers(-scale, scale + 1, (3, 32)).astype(np.int16)
        rtl = GeneratorMiniRTL(list(W), list(Bq))
        exp = np.array([rtl.run_frame([int(t) for t in f])[0] for f in frames], np.int16)
        steady = oracle.gen_fwd_q(frames, W, Bq, mode=1).reshape(-1, 32)
        first = oracle.gen_fwd_q(frames[:1], W, Bq, mode=2).reshape(-1, 32)
        assert np.array_equal(first[0], exp[0])               # first frame after reset: enc1 stale address 0
        assert np.array_equal(steady[1:], exp[1:])            # steady state: stale address 223


def test_spec_mode_primitives():
    """spec mode = RTL primitives on the textbook dataflow: check against a direct numpy evaluation."""
    rng = np.random.default_rng(1)
    W = rng.integers(-128, 128, 2048).astype(np.int8)
    Bq = np.zeros(64, np.int16)
    Bq[:18] = rng.integers(-500, 500, 18)
    x = rng.integers(-400, 401, (5, 2, 16)).astype(np.int16)
    got = oracle.gen_fwd_q(x, W, Bq, mode=0)

    def sat(v):
        return np.clip(v, -32768, 32767)

    def lrelu(v):
        return np.where(v < 0, (v >> 2) + (v >> 4), v)

    def conv(src, IN, OC, OL, stride, WA, BA, K, act):
        out = np.zeros((OC, OL), np.int64)
        for oc in range(OC):
            for p in range(OL):
                acc = 0
                for ic in range(IN):
                    for k in range(K):
                        acc += (int(src[ic][p * stride + k]) * int(W[WA + (oc * IN + ic) * K + k])) >> 7
                v = sat(acc + int(Bq[BA + oc]))
                out[oc, p] = lrelu(np.int64(v)) if act else v
        return out

    for b in range(5):
        xp = np.pad(x[b].astype(np.int64), ((0, 0), (1, 1)))
        enc = conv(xp, 2, 4, 8, 2, 0, 0, 3, True)
        bn = conv(np.pad(enc, ((0, 0), (1, 1))), 4, 8, 4, 2, 24, 4, 3, True)
        up1 = np.pad(np.repeat(bn, 2, axis=1), ((0, 0), (1, 1)))
        dec = sat(conv(up1, 8, 4, 8, 1, 120, 12, 3, True) + enc)
        out = conv(np.repeat(dec, 2, axis=1), 4, 2, 16, 1, 216, 16, 1, False)
        out = np.where(out > 256, 255, np.where(out < -256, -255, out))
        assert np.array_equal(got[b], out.astype(np.int16))


def test_q88_truncation_matches_verification_golden():
    """proof/verification.py:297-298: Q8.8 conversion is truncation toward zero; pinned by the committed vectors."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "verification_golden.npz"))
    assert np.array_equal(oracle.quantize_q88(g["input_float"]), g["input_q88"])
    assert np.array_equal(oracle.quantize_q88(g["output_float"]), g["output_q88"])
    # hex files: channel-major I[0..15] then Q[0..15], 4-digit two's complement (proof/verification.py:306-312)
    assert np.array_equal(g["input_q88"].reshape(-1).view(np.uint16), g["input_hex"])
    assert np.array_equal(g["output_q88"].reshape(-1).view(np.uint16), g["output_hex"])
    assert list(g["input_q88"].reshape(-1)[:4]) == [322, 257, -74, -10]
    # rounding would NOT reproduce the fixture
    assert not np.array_equal(np.round(g["input_float"] * 256).astype(np.int16), g["input_q88"])


def test_digest_is_order_salted():
    y = np.arange(64, dtype=np.int16)
    a = oracle.digest_i16(y)
    y2 = y.copy()
    y2[[3, 4]] = y2[[4, 3]]
    assert oracle.digest_i16(y2) != a and oracle.digest_i16(y) == a

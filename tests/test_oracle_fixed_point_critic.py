"""Oracle pinning, fixed-point critic: closed form vs the reference's RTL run (scores and the accumulate-stage trace of
tb_discriminator_mini.vcd) and vs the cycle-level emulation of discriminator_mini.v on random ROMs (CPU only)."""
import numpy as np

import oracle
from oracle.rtl_cycle_emulator import DiscriminatorMiniRTL


def _rom(rtl_vectors):
    return oracle.rom_arrays(rtl_vectors["rom"]["weights"], rtl_vectors["rom"]["biases"])


def test_committed_rom_is_nearly_empty(rtl_vectors):
    W, Bq = _rom(rtl_vectors)
    # rtl/ofdmGAN/weight_rom.v:122-159: six conv1 literals, three conv2 literals, the dense row
    assert np.count_nonzero(W[256:352]) == 6 and np.count_nonzero(W[352:736]) == 3 and np.count_nonzero(W[736:752]) == 16
    assert list(Bq[32:36]) == [8, -8, 16, -16] and Bq[55] == -2 and Bq[56] == 0


def test_cycle_emulator_matches_vcd(rtl_vectors, rtl_critic_vectors):
    rtl = DiscriminatorMiniRTL(rtl_vectors["rom"]["weights"], rtl_vectors["rom"]["biases"])
    assert len(rtl_critic_vectors["vectors"]) == 5
    for v in rtl_critic_vectors["vectors"]:
        rtl.trace = []
        score, cycles = rtl.run_frame(v["candidate"], v["condition"])
        assert score == v["score"] == -4 and rtl.dense_acc == v["dense_acc"]
        assert len(v["trace"]) == 846 and rtl.trace == v["trace"]          # every (state, ch, pos, last, ksum), in order
        assert any(t[4] != 0 for t in v["trace"]) or v["test"] == 1        # not a vacuous trace (test 1 is all zeros)


def test_rtl_literal_matches_vcd_scores(rtl_vectors, rtl_critic_vectors):
    W, Bq = _rom(rtl_vectors)
    cand = np.array([v["candidate"] for v in rtl_critic_vectors["vectors"]], np.int16)
    cond = np.array([v["condition"] for v in rtl_critic_vectors["vectors"]], np.int16)
    want = np.array([v["score"] for v in rtl_critic_vectors["vectors"]], np.int16)
    assert np.array_equal(oracle.disc_fwd_q(cand, cond, W, Bq, mode=1), want)
    assert np.array_equal(oracle.disc_fwd_q(cand, cond, W, Bq, mode=2), want)


def test_closed_form_equals_cycle_emulator_on_random_roms():
    rng = np.random.default_rng(0)
    seen = set()
    for trial in range(8):
        W = rng.integers(-128, 128, 2048).astype(np.int8)
        Bq = rng.integers(-2000, 2000, 64).astype(np.int16) if trial % 2 else rng.integers(-30000, 30000, 64).astype(np.int16)
        scale = [40, 300, 3000, 32767][trial % 4]
        cand = rng.integers(-scale, scale + 1, (3, 32)).astype(np.int16)
        cond = rng.integers(-scale, scale + 1, (3, 32)).astype(np.int16)
        rtl = DiscriminatorMiniRTL(list(W), list(Bq))
        exp = np.array([rtl.run_frame([int(t) for t in a], [int(t) for t in b])[0] for a, b in zip(cand, cond)], np.int16)
        assert oracle.disc_fwd_q(cand[:1], cond[:1], W, Bq, mode=2)[0] == exp[0]      # first frame after reset
        assert np.array_equal(oracle.disc_fwd_q(cand, cond, W, Bq, mode=1)[1:], exp[1:])   # steady state
        seen.update(int(e) for e in exp)
    assert len(seen) > 8                                                                # scores actually vary


def test_spec_mode_primitives():
    """spec mode = RTL primitives on the textbook critic dataflow: check against a direct numpy evaluation."""
    rng = np.random.default_rng(2)
    W = rng.integers(-128, 128, 2048).astype(np.int8)
    Bq = rng.integers(-500, 500, 64).astype(np.int16)
    cand = rng.integers(-600, 601, (64, 2, 16)).astype(np.int16)
    cond = rng.integers(-600, 601, (64, 2, 16)).astype(np.int16)

    def conv(x, wbase, bbase, oc_n):                    # x [C][L] int64 -> [oc_n][L/2]
        C, L = x.shape
        xp = np.pad(x, ((0, 0), (1, 1)))
        out = np.zeros((oc_n, L // 2), np.int64)
        for oc in range(oc_n):
            for p in range(L // 2):
                acc = 0
                for ic in range(C):
                    for k in range(3):
                        acc += (int(xp[ic, 2 * p + k]) * int(W[wbase + (oc * C + ic) * 3 + k])) >> 7
                v = max(-32768, min(32767, acc + int(Bq[bbase + oc])))
                out[oc, p] = ((v >> 2) + (v >> 4)) if v < 0 else v
        return out

    want = []
    for a, b in zip(cand, cond):
        c2 = conv(conv(np.concatenate([a, b]).astype(np.int64), 256, 32, 8), 352, 40, 16)
        pool = c2.sum(axis=1)
        d = sum((((int(pool[oc]) + 32768) % 65536 - 32768) * int(W[736 + oc])) >> 7 for oc in range(16)) + int(Bq[56])
        want.append(max(-32768, min(32767, d)))
    assert np.array_equal(oracle.disc_fwd_q(cand, cond, W, Bq, mode=0), np.array(want, np.int16))

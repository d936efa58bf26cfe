"""GPU parity, integer critic: ofdmgan_disc_fwd_q (through the C ABI) vs the CPU oracle, bit-exact, and vs the scores of
the reference's committed RTL run (rtl/ofdmGAN/tb_discriminator_mini.vcd)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import ofdm_gan_sr_b200 as pkg
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return pkg.ops


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def _random_roms(rng, big_bias=False):
    W = rng.integers(-128, 128, 2048).astype(np.int8)
    Bq = rng.integers(-30000, 30000, 64).astype(np.int16) if big_bias else rng.integers(-600, 600, 64).astype(np.int16)
    return W, Bq


def test_rtl_literal_matches_vcd_scores(ops, rtl_vectors, rtl_critic_vectors):
    W, Bq = oracle.rom_arrays(rtl_vectors["rom"]["weights"], rtl_vectors["rom"]["biases"])
    cand = np.array([v["candidate"] for v in rtl_critic_vectors["vectors"]], np.int16).reshape(-1, 2, 16)
    cond = np.array([v["condition"] for v in rtl_critic_vectors["vectors"]], np.int16).reshape(-1, 2, 16)
    want = np.array([v["score"] for v in rtl_critic_vectors["vectors"]], np.int16)
    got = ops.disc_fwd_q(cu(cand), cu(cond), W, Bq, mode=ops.GEN_Q_RTL).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("scale,big_bias", [(40, False), (600, False), (5000, False), (32767, False), (32767, True)])
def test_bit_exact_vs_oracle(ops, mode, scale, big_bias):
    rng = np.random.default_rng(100 * mode + scale % 97 + big_bias)
    W, Bq = _random_roms(rng, big_bias)
    B = 4099                                                   # ragged: not a multiple of the 128-frame tile
    cand = rng.integers(-scale, scale + 1, (B, 2, 16)).astype(np.int16)
    cond = rng.integers(-scale, scale + 1, (B, 2, 16)).astype(np.int16)
    cand[0], cond[0] = -32768, -32768                          # extreme frames
    cand[1], cond[1] = 32767, -32768
    cand[2], cond[2] = 0, 0
    got = ops.disc_fwd_q(cu(cand), cu(cond), W, Bq, mode=[ops.GEN_Q_SPEC, ops.GEN_Q_RTL][mode]).cpu().numpy()
    want = oracle.disc_fwd_q(cand, cond, W, Bq, mode)
    assert np.array_equal(got, want)
    assert len(np.unique(want)) > 3 or scale >= 32767          # not a degenerate comparison (big inputs saturate)


def test_spec_full_size_matches_oracle(ops):
    """2^20 frames, both modes: every score equal (the oracle does this in about a second with OpenMP)."""
    rng = np.random.default_rng(5)
    W, Bq = _random_roms(rng)
    W[256:752] = rng.integers(-40, 41, 496)                    # moderate weights: scores spread instead of saturating
    B = 1 << 20
    cand = rng.integers(-700, 701, (B, 2, 16)).astype(np.int16)
    cond = rng.integers(-700, 701, (B, 2, 16)).astype(np.int16)
    a, c = cu(cand), cu(cond)
    for mode, omode in ((ops.GEN_Q_SPEC, 0), (ops.GEN_Q_RTL, 1)):
        got = ops.disc_fwd_q(a, c, W, Bq, mode=mode).cpu().numpy()
        want = oracle.disc_fwd_q(cand, cond, W, Bq, omode)
        assert np.array_equal(got, want)
        assert len(np.unique(want)) > 100


def test_empty_and_argument_errors(ops):
    W, Bq = _random_roms(np.random.default_rng(0))
    e = torch.empty((0, 2, 16), dtype=torch.int16, device="cuda")
    assert ops.disc_fwd_q(e, e, W, Bq).shape == (0,)
    x = torch.zeros((4, 2, 16), dtype=torch.int16, device="cuda")
    with pytest.raises(Exception):
        ops.disc_fwd_q(x, x, W, Bq, mode=7)
    with pytest.raises(Exception):
        ops.disc_fwd_q(x, x[:2], W, Bq)
    with pytest.raises(Exception):
        ops.disc_fwd_q(x, x, W[:100], Bq)


def test_rom_change_between_calls_takes_effect(ops):
    rng = np.random.default_rng(9)
    cand = rng.integers(-500, 501, (256, 2, 16)).astype(np.int16)
    cond = rng.integers(-500, 501, (256, 2, 16)).astype(np.int16)
    a, c = cu(cand), cu(cond)
    for _ in range(3):
        W, Bq = _random_roms(rng)
        assert np.array_equal(ops.disc_fwd_q(a, c, W, Bq).cpu().numpy(), oracle.disc_fwd_q(cand, cond, W, Bq, 0))

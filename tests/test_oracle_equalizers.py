"""Oracle pinning, genie-aided ZF / MMSE equalisers: the C restatement vs fixtures recorded from the reference
(tests/golden/make_eq_fixtures.py).  CPU only."""
import os

import numpy as np

import oracle
from conftest import GOLDEN, assert_close


def _ref():
    return dict(np.load(os.path.join(GOLDEN, "ref_eq.npz")))


def test_zero_forcing_is_bit_identical_to_the_reference():
    r = _ref()
    zf = oracle.equalize(r["noisy"], r["clean"], None, method=2)
    assert np.array_equal(zf, r["zf"])                       # complex64 Smith division, operation for operation
    m = oracle.frame_metrics(zf, r["clean"], None, method=2, n_snr=1)[0, 2]
    assert m[0] == 700
    assert_close(m[1], r["metrics"][:, 0].sum(), 1e-5, "sum mse")
    assert_close(m[3], r["metrics"][:, 1].sum(), 1e-5, "sum evm")
    assert -151 < r["metrics"][:, 1].min() and r["metrics"][:, 1].max() < -136     # rounding noise of complex64, not -200 dB


def test_mmse_matches_the_reference():
    r = _ref()
    mm = oracle.equalize(r["noisy"], r["clean"], r["snr"], method=3)
    assert_close(mm, r["mmse"], 1e-6, "MMSE frames")         # np.abs is a SIMD hypot: last bit not reproduced
    bins = (r["snr"] / 5).astype(np.int32)
    m = oracle.frame_metrics(mm, r["clean"], bins, method=3, n_snr=7)[:, 3]
    ref = r["metrics"].reshape(7, 100, 4)
    assert_close(m[:, 1], ref[:, :, 2].sum(1), 1e-5, "sum mse per snr")
    assert_close(m[:, 3], ref[:, :, 3].sum(1), 1e-5, "sum evm per snr")


def test_fused_restatement_fills_the_equaliser_rows(ref_fp32):
    kw = dict(nonlinear=True, pa_saturation=0.8, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=50, normalize=2)
    m = oracle.sim_gen_metrics(oracle.make_cfg(equalizers=True, **kw), 0, 700, gparams=ref_fp32["gparams"], seed=3)
    m0 = oracle.sim_gen_metrics(oracle.make_cfg(**kw), 0, 700, gparams=ref_fp32["gparams"], seed=3)
    assert_close(m[:, :2], m0[:, :2], 1e-12, "GAN / NoEQ rows unchanged")      # (OpenMP merge order differs run to run)
    assert np.all(m0[:, 2:] == 0) and np.all(m[:, 2:, 0] == 100)
    s = oracle.metrics_summary(m)
    assert np.all(s["evm"][:, 2] < -130)                     # ZF with a genie channel: rounding noise only
    assert np.all(np.diff(s["evm"][:, 3]) < 0)               # MMSE improves with SNR
    clean, noisy, snr = oracle.chan_sim(oracle.make_cfg(**kw), 700, seed=3)
    bins = ((np.arange(700) // 50) % 7).astype(np.int32)
    for method in (2, 3):
        m2 = oracle.frame_metrics(oracle.equalize(noisy, clean, snr, method), clean, bins, method, 7)
        assert_close(m[:, method, :5], m2[:, method, :5], 1e-9, "fused vs unfused")

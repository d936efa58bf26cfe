"""GPU parity of the lean headline simulator kernel (csrc/sim_lean.cu) against the general one (csrc/sim_kernel.cuh, selected
with OFDMGAN_SIM_IMPL=general) and against the CPU oracle, over every option the lean kernel accepts.  Both kernels implement
utils/dataset.py:236-293 / benchmark_comparison.py:179-250; they differ only in where scales are folded (~1e-7)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import ofdm_gan_sr_b200 as pkg
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return pkg.ops


def host(t):
    return t.detach().cpu().numpy()


CASES = {
    "awgn_joint": dict(),
    "awgn10_none": dict(snr_lo=10.0, snr_hi=10.0, normalize=0),
    "nl08_joint": dict(nonlinear=True, pa_saturation=0.8),
    "nl_sep_grid": dict(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=100),
    "nl_p2_bigpn": dict(nonlinear=True, pa_saturation=1.3, pa_smoothness=2.0, phase_noise_dbchz=-66.0, iq_imbalance_db=-0.7, iq_phase_deg=-4.0),
    "pa_only_scaleN": dict(pa=True, pa_saturation=0.6, ifft_scale=1, normalize=2),
    "pn_only_nosnr": dict(pn=True, snr_mode=2),
    "iq_only_grid1": dict(iq=True, snr_mode=1, snr_lo=7.0, snr_step=0.0, n_snr=1, frames_per_snr=1),
}


@pytest.mark.parametrize("tag", sorted(CASES))
@pytest.mark.parametrize("B,frame0", [(1, 0), (33, 5), (1000, (1 << 33) + 17), (20011, 123456789)])
def test_lean_frames_equal_general_and_oracle(ops, monkeypatch, tag, B, frame0):
    kw = CASES[tag]
    cfg = ops.make_cfg(**kw)
    clean, noisy, snr = (host(t) for t in ops.chan_sim(cfg, B, seed=11, frame0=frame0))
    monkeypatch.setenv("OFDMGAN_SIM_IMPL", "general")
    gclean, gnoisy, gsnr = (host(t) for t in ops.chan_sim(cfg, B, seed=11, frame0=frame0))
    monkeypatch.delenv("OFDMGAN_SIM_IMPL")
    assert np.array_equal(snr, gsnr)
    if kw.get("snr_mode") == 2:
        assert np.all(np.isinf(snr))                                 # "no noise" is reported as +inf dB
    assert_close(clean, gclean, 2e-6, f"lean vs general clean [{tag}]")
    assert_close(noisy, gnoisy, 2e-6, f"lean vs general noisy [{tag}]")
    if B <= 1000:
        oc, on, _ = oracle.chan_sim(oracle.make_cfg(**kw), B, seed=11, frame0=frame0)
        # the oracle draws in float64 libm; the device in MUFU approximations: 5e-5 to scale on its own Philox stream
        assert_close(clean, oc, 5e-5, f"lean vs oracle clean [{tag}]")
        assert_close(noisy, on, 5e-5, f"lean vs oracle noisy [{tag}]")


@pytest.mark.parametrize("tag", ["awgn_joint", "nl08_joint", "nl_sep_grid", "nl_p2_bigpn"])
def test_lean_injected_draws_equal_oracle(ops, tag):
    """host-generated draws (the reference's np.random order): the injected-draw instantiation against the float64 oracle"""
    kw = CASES[tag]
    B = 777
    rng = np.random.default_rng(5)
    sym, pn, noise = rng.standard_normal((B, 32)), rng.standard_normal((B, 16)), rng.standard_normal((B, 32))
    snr_db = rng.uniform(0, 30, B)
    cfg = ops.make_cfg(**kw)
    grid = kw.get("snr_mode", 0) == 1
    clean, noisy, snr = (host(t) for t in ops.chan_sim(cfg, B, sym=sym, pn=pn, noise=noise, snr_db=None if grid else snr_db))
    oc, on, osnr = oracle.chan_sim(oracle.make_cfg(**kw), B, sym=sym, pn=pn, noise=noise, snr_db=None if grid else snr_db)
    assert_close(clean, oc, 1e-5, f"lean injected clean [{tag}]")
    assert_close(noisy, on, 1e-5, f"lean injected noisy [{tag}]")
    assert_close(snr, osnr, 1e-6, f"lean injected snr [{tag}]")


METRIC_CASES = dict(CASES, grid16_ragged=dict(nonlinear=True, pa_saturation=0.8, normalize=1, snr_mode=1, snr_lo=-2.0, snr_step=2.0, n_snr=16,
                                               frames_per_snr=7),
                    uniform_snr=dict(nonlinear=True, pa_saturation=0.9, normalize=2, snr_lo=3.0, snr_hi=27.0))


@pytest.mark.parametrize("tag", ["nl_sep_grid", "awgn_joint", "nl_p2_bigpn", "grid16_ragged", "uniform_snr"])
def test_lean_metrics_equal_general_and_oracle(ops, monkeypatch, tag):
    kw = dict(METRIC_CASES[tag])
    if kw.get("snr_mode", 0) != 1 and tag != "uniform_snr":
        kw.update(snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=64)
    rng = np.random.default_rng(2)
    gp = (rng.standard_normal(258) * 0.3).astype(np.float32)
    B, frame0 = 7 * 300 + 13, 999
    cfg = ops.make_cfg(**kw)
    m = host(ops.sim_gen_metrics(cfg, B, gparams=gp, seed=4, frame0=frame0))
    monkeypatch.setenv("OFDMGAN_SIM_IMPL", "general")
    g = host(ops.sim_gen_metrics(cfg, B, gparams=gp, seed=4, frame0=frame0))
    monkeypatch.delenv("OFDMGAN_SIM_IMPL")
    o = oracle.sim_gen_metrics(oracle.make_cfg(**kw), 0, B, gparams=gp, seed=4, frame0=frame0)
    assert np.array_equal(m[:, :2, 0], g[:, :2, 0]) and np.array_equal(m[:, :2, 0], o[:, :2, 0])
    assert m[:, 2:].sum() == 0                                    # no equaliser rows from this kernel
    for c in (1, 2, 3, 4, 7):
        assert_close(m[:, :2, c], g[:, :2, c], 1e-5, f"lean vs general metrics col {c} [{tag}]")
        assert_close(m[:, :2, c], o[:, :2, c], 5e-5, f"lean vs oracle metrics col {c} [{tag}]")


def test_lean_is_what_the_headline_calls_run_on(ops):
    """the dispatch: the benchmark configuration goes to the lean kernel, anything it does not build stays general"""
    import ctypes

    import ofdm_gan_sr_b200 as pkg
    f = pkg._lib.lib().ofdmgan_sim_impl_for
    head = ops.make_cfg(nonlinear=True, pa_saturation=0.8, normalize=1, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=1024)
    assert f(ctypes.byref(head), ops.GEN_F32, 0) == 1
    assert f(ctypes.byref(head), -1, 0) == 1
    assert f(ctypes.byref(head), ops.GEN_Q_SPEC, 0) == 0
    assert f(ctypes.byref(ops.make_cfg(equalizers=True)), ops.GEN_F32, 0) == 0
    assert f(ctypes.byref(ops.make_cfg(channel_type="rayleigh")), ops.GEN_F32, 0) == 0
    assert f(ctypes.byref(ops.make_cfg(symbol_source=1)), -1, 0) == 0
    assert f(ctypes.byref(head), ops.GEN_F32, 1) == 0                # caller-supplied time-domain frames / fading draws


# ------------------------------------------------------------------------------------------------ the fast-RNG workload (Philox4x32-7)
def test_philox7_workload_matches_its_oracle_and_is_a_different_stream(ops):
    """ofdmgan_chan_cfg.rng_rounds = 7: the separately named fast-RNG workload.  Same algorithm with seven rounds on both sides (the
    ten-round generator is the one pinned by Random123's known answers); frames and metrics against the oracle like the default."""
    import ofdm_gan_sr_b200 as pkg
    kw = dict(nonlinear=True, pa_saturation=0.8, normalize=1, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=64)
    c7, c10 = ops.make_cfg(rng_rounds=7, **kw), ops.make_cfg(**kw)
    d = ops.chan_draws(c7, 2048, seed=5, frame0=77)
    o = oracle.frame_draws(oracle.make_cfg(rng_rounds=7, **kw), 5, 77, 2048)
    for k in ("sym", "pn", "noise"):
        e = np.abs(host(d[k]) - o[k])
        assert e.mean() < 1e-6 and np.quantile(e, 0.999) < 5e-6 and e.max() < 2e-4, k
    assert np.abs(host(d["sym"]) - host(ops.chan_draws(c10, 2048, seed=5, frame0=77)["sym"])).max() > 0.5
    B = 7 * 64 * 5 + 3
    clean, noisy, _ = (host(t) for t in ops.chan_sim(c7, B, seed=5, frame0=77))
    oc, on, _ = oracle.chan_sim(oracle.make_cfg(rng_rounds=7, **kw), B, seed=5, frame0=77)
    assert_close(clean, oc, 5e-5, "philox-7 clean vs oracle")
    assert_close(noisy, on, 5e-5, "philox-7 noisy vs oracle")
    gp = (np.random.default_rng(1).standard_normal(258) * 0.3).astype(np.float32)
    m = host(ops.sim_gen_metrics(c7, B, gparams=gp, seed=5, frame0=77))
    om = oracle.sim_gen_metrics(oracle.make_cfg(rng_rounds=7, **kw), 0, B, gparams=gp, seed=5, frame0=77)
    assert np.array_equal(m[:, :2, 0], om[:, :2, 0])
    for c in (1, 3):
        assert_close(m[:, :2, c], om[:, :2, c], 5e-5, f"philox-7 metrics col {c}")
    m10 = host(ops.sim_gen_metrics(c10, B, gparams=gp, seed=5, frame0=77))
    assert not np.allclose(m[:, :2, 1], m10[:, :2, 1], rtol=1e-6)
    # built for the headline kernel only
    with pytest.raises(pkg.OfdmGanError):
        ops.sim_gen_metrics(ops.make_cfg(rng_rounds=7, equalizers=True, **kw), 64, gparams=gp)
    with pytest.raises(pkg.OfdmGanError):
        ops.chan_sim(ops.make_cfg(rng_rounds=7, symbol_source=1, **kw), 64)
    with pytest.raises(pkg.OfdmGanError):
        ops.chan_sim(ops.make_cfg(rng_rounds=5, **kw), 64)

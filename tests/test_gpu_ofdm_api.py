"""GPU parity of the utils/ofdm_utils.py call surface against fixtures recorded from the reference itself
(tests/golden/make_api_fixtures.py): QPSK map / hard decisions bit-exact, everything else <= 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, assert_close

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def api():
    import ofdm_gan_sr_b200 as pkg
    import ofdm_gan_sr_b200.utils as u
    assert torch.cuda.is_available()
    return pkg, u, dict(np.load(os.path.join(GOLDEN, "ref_api.npz")))


def c2(a):
    """complex array -> stacked (re, im) float64 for assert_close"""
    a = np.asarray(a)
    return np.stack([a.real, a.imag]).astype(np.float64)


def tx_of(x):
    return np.concatenate([x.real, x.imag], axis=1).astype(np.float32)


def test_qpsk_modulate_demodulate_bit_exact(api):
    pkg, u, r = api
    q = u.QAMModulator("QPSK")
    syms = q.modulate(r["bits"])                                  # numpy in -> numpy out; the odd trailing bit is dropped
    assert syms.dtype == np.complex128 and syms.shape == r["syms"].shape
    assert_close(c2(syms), c2(r["syms"]), 1e-7, "QPSK symbols")
    assert np.array_equal(np.sign(syms.real), np.sign(r["syms"].real)) and np.array_equal(np.sign(syms.imag), np.sign(r["syms"].imag))
    bits = q.demodulate(r["noisy_syms"])                          # includes exact ties: argmin -> lowest index
    assert np.array_equal(bits, r["demod_bits"])
    # torch in -> torch out, round trip at scale
    b = torch.randint(0, 2, (2 * 1_000_003,), device="cuda", dtype=torch.uint8)
    s = q.modulate(b)
    assert s.is_cuda and s.dtype == torch.complex64 and torch.equal(q.demodulate(s), b)
    with pytest.raises(ValueError):
        u.QAMModulator("PSK8")
    with pytest.raises(pkg.OfdmGanError):
        q.modulate(torch.zeros(8, dtype=torch.uint8))             # CPU tensor: no fallback


@pytest.mark.parametrize("tag,N,cp,sp,pv", [("o8", 8, 2, 4, 1 + 0j), ("o16", 16, 4, 8, 0.5 - 0.5j), ("o16np", 16, 0, 32, 1 + 0j)])
def test_ofdm_modulator_matches_reference(api, tag, N, cp, sp, pv):
    pkg, u, r = api
    m = u.OFDMModulator(N, cp, sp, pv)
    assert m.n_data_subcarriers == N - len(range(0, N, sp)) and m.samples_per_symbol == N + cp
    sig = m.modulate(r[tag + "_in"])
    assert sig.shape == r[tag + "_sig"].shape
    assert_close(c2(sig), c2(r[tag + "_sig"]), TOL, tag + " modulate")
    data, chan = m.demodulate(r[tag + "_rx"])
    assert data.shape == r[tag + "_data"].shape and chan.shape == r[tag + "_chan"].shape
    assert_close(c2(data), c2(r[tag + "_data"]), TOL, tag + " demodulate data")
    assert_close(c2(chan), c2(r[tag + "_chan"]), TOL, tag + " channel estimate")
    # modulate -> demodulate is the identity on the payload; QPSK decisions survive it
    q = u.QAMModulator("QPSK")
    bits = torch.randint(0, 2, (2 * 5 * m.n_data_subcarriers * 1000,), device="cuda", dtype=torch.uint8)
    back, _ = m.demodulate(m.modulate(q.modulate(bits)))
    if cp == 0:                                                   # reference quirk: cp_length 0 emits every symbol twice
        back = back.view(-1, 2, m.n_data_subcarriers)[:, 0].reshape(-1)
    assert torch.equal(q.demodulate(back), bits)
    with pytest.raises(pkg.OfdmGanError):
        u.OFDMModulator(128, 16, 8).modulate(r[tag + "_in"])      # valid upstream, not built here (8/16/32/64 are)


def test_memoryless_impairments_match_reference(api):
    pkg, u, r = api
    NL = u.NonLinearImpairments
    assert_close(c2(NL.apply_pa_rapp(r["x"], 0.8, 3.0)), c2(r["pa_08_3"]), TOL, "Rapp A=0.8 p=3")
    assert_close(c2(NL.apply_pa_rapp(r["x"], 1.0, 2.0)), c2(r["pa_10_2"]), TOL, "Rapp A=1 p=2")
    assert_close(c2(NL.apply_iq_imbalance(r["x"], 1.0, 5.0)), c2(r["iq_1_5"]), TOL, "IQ 1 dB 5 deg")
    assert_close(c2(NL.apply_iq_imbalance(r["x"], -0.5, -3.0)), c2(r["iq_m05_m3"]), TOL, "IQ -0.5 dB -3 deg")
    y = NL.apply_pa_rapp(r["x_long"], 0.7, 3.0)                   # any length for the memoryless stages
    assert y.shape == (37,)
    assert_close(c2(y), c2(r["pa_long"]), TOL, "Rapp, 37 samples")
    xt = torch.as_tensor(r["x"]).cuda()
    yt = NL.apply_pa_rapp(xt, 0.8, 3.0)
    assert yt.is_cuda and yt.dtype == torch.complex64 and yt.shape == xt.shape


def test_random_stages_match_reference_with_its_draws(api):
    """apply_phase_noise / apply_all / ChannelModel.apply replayed with the reference's recorded np.random draws."""
    pkg, u, r = api
    ops = pkg.ops
    B = r["x"].shape[0]
    tx = tx_of(r["x"])
    kw = dict(normalize=0, snr_mode=ops.SNR_NONE)
    _, y, _ = ops.chan_sim(ops.make_cfg(pa=False, iq=False, pn=True, **kw), B, tx=tx, pn=r["pn_draws"])
    got = y.cpu().numpy()
    assert_close(np.stack([got[:, 0], got[:, 1]]), c2(r["pn"]), TOL, "phase noise")
    _, y, _ = ops.chan_sim(ops.make_cfg(nonlinear=True, pa_saturation=0.8, **kw), B, tx=tx, pn=r["all_draws"])
    got = y.cpu().numpy()
    assert_close(np.stack([got[:, 0], got[:, 1]]), c2(r["all"]), TOL, "apply_all")
    cfg = ops.make_cfg(normalize=0, snr_mode=ops.SNR_GRID, snr_lo=12.5, snr_step=0.0, n_snr=1)
    _, y, s = ops.chan_sim(cfg, B, tx=tx, noise=r["awgn_draws"])
    got = y.cpu().numpy()
    assert_close(np.stack([got[:, 0], got[:, 1]]), c2(r["awgn"]), TOL, "AWGN")
    assert torch.all(s == 12.5)


def test_random_stages_surface_and_statistics(api):
    pkg, u, r = api
    NL, ops = u.NonLinearImpairments, pkg.ops
    x = torch.as_tensor(r["x"]).cuda().repeat(4000, 1)             # [96000, 16]
    y, info = u.ChannelModel("awgn").apply(x, 10.0, seed=3)
    assert y.shape == x.shape and info["type"] == "awgn" and info["snr_db"] == 10.0
    p_sig = (x.abs() ** 2).mean(dim=1)
    p_noise = ((y - x.to(torch.complex64)).abs() ** 2).mean(dim=1)
    assert abs(float((p_noise / p_sig).mean()) - 0.1) < 2e-3       # measured-power AWGN at 10 dB
    assert_close(info["noise_power"].cpu().numpy()[:24], r["awgn_noise_power"] * 10 ** (12.5 / 10) / 10, 2e-6, "noise power")
    y2, _ = u.ChannelModel("awgn").apply(x, 10.0, seed=3, frame0=0)
    y3, _ = u.ChannelModel("awgn").apply(x, 10.0, seed=3, frame0=0)
    assert torch.equal(y2, y3)                                     # counter-based: reproducible on request
    y4, _ = u.ChannelModel("awgn").apply(x, 10.0, seed=3)
    assert not torch.equal(y4, y)                                  # ... and fresh noise on successive calls otherwise
    z = NL.apply_phase_noise(x, -80, 1e6, seed=1)
    ang = torch.angle(z * x.to(torch.complex64).conj())           # accumulated phase: Wiener process with sigma 0.1 per step
    inc = torch.diff(ang, dim=1)
    assert abs(float(inc.std()) - 0.1) < 2e-3 and abs(float(ang[:, 0].std()) - 0.1) < 2e-3
    a = NL.apply_all(x[:24], pa_saturation=0.8, seed=5, frame0=0)
    b = NL.apply_all(x[:24], pa_saturation=0.8, seed=5, frame0=0)
    assert torch.equal(a, b)
    one = NL.apply_all(r["x"][0], pa_saturation=0.8)               # a single NumPy frame, like the reference's callers
    assert isinstance(one, np.ndarray) and one.shape == (16,) and one.dtype == np.complex128
    with pytest.raises(pkg.OfdmGanError):
        NL.apply_phase_noise(r["x_long"])                          # frame-coupled stage: 16-sample frames only
    with pytest.raises(ValueError):
        u.ChannelModel("bogus").apply(r["x"], 10.0)


# ------------------------------------------------------------------------------------------------ classical equalisers
def test_equalizers_match_reference(api):
    """ZF bit-identical to the reference's complex64 arithmetic; MMSE to 1e-6 (np.abs' SIMD hypot is not reproduced)."""
    pkg, u, _ = api
    r = dict(np.load(os.path.join(GOLDEN, "ref_eq.npz")))
    zf, met = u.ZeroForcingEqualizer().equalize_iq(r["noisy"], r["clean"])
    assert isinstance(zf, np.ndarray) and np.array_equal(zf, r["zf"])
    mm, _ = u.MMSEEqualizer().equalize_iq(torch.as_tensor(r["noisy"]).cuda(), torch.as_tensor(r["clean"]).cuda(),
                                          snr_db=torch.as_tensor(r["snr"]))
    assert mm.is_cuda
    assert_close(mm.cpu().numpy(), r["mmse"], 1e-6, "MMSE frames")
    one, m1 = u.MMSEEqualizer().equalize_iq(r["noisy"][5], r["clean"][5], snr_db=float(r["snr"][5]))
    assert one.shape == (2, 16) and isinstance(m1["mse"], float)
    assert_close(one, r["mmse"][5], 1e-6, "single frame")
    # per-trial metrics through the metric kernel, against compute_mse / compute_evm of the reference
    ops = pkg.ops
    bins = torch.as_tensor((r["snr"] / 5).astype(np.int32)).cuda()
    c = torch.as_tensor(r["clean"]).cuda()
    m = ops.frame_metrics(torch.as_tensor(zf).cuda(), c, bins, method=2, n_snr=7)
    m = ops.frame_metrics(mm, c, bins, method=3, n_snr=7, out=m).cpu().numpy()
    ref = r["metrics"].reshape(7, 100, 4)
    for method, col in ((2, 0), (3, 2)):
        assert np.all(m[:, method, 0] == 100)
        assert_close(m[:, method, 1], ref[:, :, col].sum(1), TOL, f"method {method} sum mse")
        assert_close(m[:, method, 3], ref[:, :, col + 1].sum(1), TOL, f"method {method} sum evm")
    with pytest.raises(pkg.OfdmGanError):
        u.ZeroForcingEqualizer().equalize_iq(r["noisy"])


def test_fused_equaliser_rows_vs_oracle_and_pieces(api):
    import oracle
    pkg, u, _ = api
    ops = pkg.ops
    gp = np.load(os.path.join(GOLDEN, "ref_fp32.npz"))["gparams"]
    kw = dict(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=512,
              equalizers=True)
    B = 7 * 512 * 6
    m = ops.sim_gen_metrics(ops.make_cfg(**kw), B, gparams=gp, seed=4).cpu().numpy()
    o = oracle.sim_gen_metrics(oracle.make_cfg(**kw), 0, B, gparams=gp, seed=4)
    assert np.array_equal(m[:, :, 0], o[:, :, 0])
    for c in (1, 3, 4):
        assert_close(m[:, :2, c], o[:, :2, c], 2e-5, f"GAN/NoEQ col {c}")
        assert_close(m[:, 3, c], o[:, 3, c], 2e-5, f"MMSE col {c}")
    # ZF's residual is the rounding noise of its inputs, and the fp32 simulator's frames differ from the float64 oracle's in
    # the last bit (and are scaled by a reciprocal multiply instead of a division): the rows agree statistically (mean EVM to
    # ~0.2 dB at -142 dB), not digit for digit.  Digit-for-digit agreement on identical inputs is asserted below.
    assert_close(m[:, 2, 3], o[:, 2, 3], 3e-3, "ZF sum evm")
    # exactness on identical inputs: the same frames through the separate entry points
    clean, noisy, snr = ops.chan_sim(ops.make_cfg(**kw), B, seed=4)
    bins = torch.as_tensor(((np.arange(B) // 512) % 7).astype(np.int32)).cuda()
    p = ops.frame_metrics(ops.equalize(noisy, clean, pkg._lib.METHOD_ZF), clean, bins, method=2, n_snr=7)
    p = ops.frame_metrics(ops.equalize(noisy, clean, pkg._lib.METHOD_MMSE, snr_db=snr), clean, bins, method=3, n_snr=7, out=p).cpu().numpy()
    for c in (1, 3, 4):
        assert_close(m[:, 2:, c], p[:, 2:, c], 1e-6, f"fused vs pieces col {c}")
    assert np.array_equal(ops.equalize(noisy, clean, pkg._lib.METHOD_ZF).cpu().numpy(),
                          oracle.equalize(noisy.cpu().numpy(), clean.cpu().numpy(), None, 2))
    # run_benchmark now reports the four data-parallel methods
    res = pkg.sweep.run_benchmark(gp, n_trials=1000, nonlinear=True, pa_saturation=0.8, seed=2)
    assert list(res) == ["GAN", "ZF", "MMSE", "NoEQ"] and res["ZF"][10.0]["evm"] < -130 < res["MMSE"][10.0]["evm"] < res["NoEQ"][10.0]["evm"]
    # integer generator x equaliser rows: built for the Gaussian source (tests/test_gpu_api.py), not for the QPSK sources
    ops.sim_gen_metrics(ops.make_cfg(**kw), 64, gen_kind=1, wrom=np.zeros(2048, np.int8), brom=np.zeros(64, np.int16))
    with pytest.raises(pkg.OfdmGanError):
        ops.sim_gen_metrics(ops.make_cfg(symbol_source=ops.SYM_QPSK, **kw), 64, gen_kind=1, wrom=np.zeros(2048, np.int8), brom=np.zeros(64, np.int16))


# ------------------------------------------------------------------------------------------------ remaining models
@pytest.fixture(scope="module")
def api2(api):
    return dict(np.load(os.path.join(GOLDEN, "ref_api2.npz")))


@pytest.mark.parametrize("tag,name", [("q16", "QAM16"), ("q64", "QAM64")])
def test_qam16_qam64_match_reference(api, api2, tag, name):
    pkg, u, _ = api
    q = u.QAMModulator(name)
    assert_close(c2(q.constellation), c2(api2[tag + "_const"]), 1e-12, name + " constellation")
    syms = q.modulate(api2[tag + "_bits"])
    assert syms.shape == api2[tag + "_syms"].shape
    assert_close(c2(syms), c2(api2[tag + "_syms"]), 1e-6, name + " symbols")
    # includes the symmetric tie at 0 and out-of-range points.  Points 5 and 6 of the fixture sit exactly half-way between two
    # asymmetric levels (2/sqrt(10)): the reference's decision there is whatever float64 rounding of |s-c|^2 happens to give -
    # not a property to reproduce - so they are left out of the comparison.
    keep = np.ones(300, bool)
    keep[5:7] = False
    got = q.demodulate(api2[tag + "_noisy"]).reshape(300, -1)
    assert np.array_equal(got[keep], api2[tag + "_demod"].reshape(300, -1)[keep])
    bits = torch.randint(0, 2, (q.bits_per_symbol * 500_000,), device="cuda", dtype=torch.uint8)
    assert torch.equal(q.demodulate(q.modulate(bits)), bits)


def test_saleh_dc_cfo_match_reference(api, api2):
    pkg, u, _ = api
    NL, r = u.NonLinearImpairments, api2
    assert_close(c2(NL.apply_pa_saleh(r["x"])), c2(r["saleh"]), TOL, "Saleh defaults")
    assert_close(c2(NL.apply_pa_saleh(r["x"], 1.5, 0.8, 2.0, 5.0)), c2(r["saleh2"]), TOL, "Saleh custom")
    assert_close(c2(NL.apply_dc_offset(r["x"], 0.01, 0.01)), c2(r["dc"]), TOL, "DC offset")
    assert_close(c2(NL.apply_dc_offset(r["x"], -0.05, 0.2)), c2(r["dc2"]), TOL, "DC offset custom")
    assert_close(c2(NL.apply_cfo(r["x"], 100, 1e6)), c2(r["cfo"]), TOL, "CFO 100 Hz")
    assert_close(c2(NL.apply_cfo(r["x"], 25000, 1e6)), c2(r["cfo2"]), TOL, "CFO 25 kHz")
    ops = pkg.ops
    cfg = ops.make_cfg(nonlinear=True, pa_saturation=0.8, dc_offset=(0.01, 0.01), cfo_hz=100, normalize=0, snr_mode=ops.SNR_NONE)
    _, y, _ = ops.chan_sim(cfg, 24, tx=tx_of(r["x"]), pn=r["all_dc_d"])
    got = y.cpu().numpy()
    assert_close(np.stack([got[:, 0], got[:, 1]]), c2(r["all_dc"]), TOL, "apply_all with DC + CFO")


@pytest.mark.parametrize("kind,kw,nf", [("rayleigh", {}, 2), ("rician", dict(k_factor=4.0), 3), ("multipath", {}, 6),
                                        ("multipath", dict(delays=[0, 3], powers=[2.0, 1.0]), 4)])
def test_fading_channels_match_reference_with_its_draws(api, api2, kind, kw, nf):
    pkg, u, _ = api
    ops, r = pkg.ops, api2
    tag = {"rayleigh": "ray", "rician": "ric"}.get(kind, "mp2" if kw else "mp")
    d = r[tag + "_d"]
    fade = np.zeros((24, 8))
    fade[:, :nf] = d[:, :nf]
    cfg = ops.make_cfg(normalize=0, snr_mode=ops.SNR_GRID, snr_lo=15.0, snr_step=0.0, n_snr=1, channel_type=kind,
                       rician_k=kw.get("k_factor", 3.0), delays=kw.get("delays", (0, 1, 2)), powers=kw.get("powers", (1.0, 0.5, 0.25)))
    _, y, _ = ops.chan_sim(cfg, 24, tx=tx_of(r["x"]), fade=fade, noise=d[:, nf:])
    got = y.cpu().numpy()
    assert_close(np.stack([got[:, 0], got[:, 1]]), c2(r[tag]), TOL, kind)


def test_fading_channel_surface_and_statistics(api, api2):
    pkg, u, _ = api
    x = torch.as_tensor(api2["x"]).cuda().repeat(4000, 1)
    y, info = u.ChannelModel("rayleigh").apply(x, 300.0, seed=2, frame0=0)                # 300 dB: fading only
    h = info["channel_response"]
    assert h.shape == (96000,) and info["type"] == "rayleigh"
    assert abs(float((h.abs() ** 2).mean()) - 1.0) < 0.02                               # h ~ CN(0, 1)
    assert_close(c2((y / x.to(torch.complex64)).cpu().numpy()[:, 0]), c2(h.cpu().numpy()), 1e-4, "y = h x")
    y, info = u.ChannelModel("rician").apply(x, 300.0, k_factor=10.0, seed=2, frame0=0)
    assert abs(float((info["channel_response"].abs() ** 2).mean()) - 1.0) < 0.02 and info["k_factor"] == 10.0
    assert float(info["channel_magnitude"].std()) < 0.35                                  # strong LOS: little spread
    y, info = u.ChannelModel("multipath").apply(x, 300.0, seed=2, frame0=0)
    assert info["channel_response"].shape == (96000, 3) and abs(sum(info["powers"]) - 1.0) < 1e-12
    one, info1 = u.ChannelModel("rayleigh").apply(api2["x"][0], 10.0)
    assert isinstance(one, np.ndarray) and one.shape == (16,) and isinstance(info1["channel_magnitude"], float)
    ds = u.SyntheticOFDMDataset(n_samples=256, channel_type="rayleigh", seed=1)
    b = ds.batch(0, 256)
    assert b["noisy"].shape == (256, 2, 16) and torch.isfinite(b["noisy"]).all()
    with pytest.raises(ValueError):
        u.SyntheticOFDMDataset(channel_type="bogus")

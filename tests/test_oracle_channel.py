"""Oracle pinning, channel simulator / metrics / RNG: the C restatement vs fixtures recorded from the imported
reference with its np.random draws captured.  CPU only."""
import numpy as np

import oracle
from conftest import assert_close


def _cfg_dataset(tag):
    kw = dict(awgn10=dict(), awgn=dict(),
              nl08=dict(nonlinear=True, pa_saturation=0.8),
              nl10=dict(nonlinear=True, pa_saturation=1.0, iq_imbalance_db=0.5, iq_phase_deg=-3.0,
                        phase_noise_dbchz=-85.0))[tag]
    return oracle.make_cfg(normalize=1, **kw)


def test_dataset_samples_match_reference(ref_channel):
    """SyntheticOFDMDataset.__getitem__ (utils/dataset.py:236-293) replayed from its recorded draws."""
    r = ref_channel
    for tag in ("awgn10", "awgn", "nl08", "nl10"):
        cfg = _cfg_dataset(tag)
        clean, noisy, snr = oracle.chan_sim(cfg, 64, sym=r[tag + "_sym"], pn=r[tag + "_pn"], snr_db=r[tag + "_snr"],
                                            noise=r[tag + "_noise"])
        assert_close(clean, r[tag + "_clean"], 2e-7, tag + " clean")
        assert_close(noisy, r[tag + "_noisy"], 2e-7, tag + " noisy")
        assert_close(snr, r[tag + "_snr"].astype(np.float32), 1e-7, tag + " snr")
        # joint normalisation: the larger of the two peaks is exactly 1
        peak = np.maximum(np.abs(clean).max(axis=(1, 2)), np.abs(noisy).max(axis=(1, 2)))
        assert np.all(peak == 1.0)


def test_benchmark_frames_and_metrics_match_reference(ref_channel):
    """benchmark_comparison.py:184-214: separate normalisation, fixed SNR grid, per-trial MSE / EVM(dB)."""
    r = ref_channel
    for tag, nl in (("bm_lin", False), ("bm_nl", True)):
        cfg = oracle.make_cfg(normalize=2, nonlinear=nl, pa_saturation=0.8 if nl else 1.0)
        n = len(r[tag + "_snr"])
        clean, noisy, _ = oracle.chan_sim(cfg, n, sym=r[tag + "_sym"], pn=r[tag + "_pn"], snr_db=r[tag + "_snr"],
                                          noise=r[tag + "_noise"])
        assert_close(clean, r[tag + "_clean"], 2e-7, tag + " clean")
        assert_close(noisy, r[tag + "_noisy"], 2e-7, tag + " noisy")
        bins = (r[tag + "_snr"] / 5).astype(np.int32)
        m_gan = oracle.frame_metrics(r[tag + "_gan"], r[tag + "_clean"], bins, method=0, n_snr=7)
        m_no = oracle.frame_metrics(r[tag + "_noisy"], r[tag + "_clean"], bins, method=1, n_snr=7)
        ref = r[tag + "_metrics"].reshape(7, 6, 4)                 # [snr][trial][gan mse, gan evm, noeq mse, noeq evm]
        for m, col, method in ((m_gan, 0, 0), (m_no, 2, 1)):
            assert np.all(m[:, method, 0] == 6)
            assert_close(m[:, method, 1], ref[:, :, col].sum(1), 1e-5, tag + " sum mse")
            assert_close(m[:, method, 2], (ref[:, :, col] ** 2).sum(1), 1e-5, tag + " sum mse^2")
            assert_close(m[:, method, 3], ref[:, :, col + 1].sum(1), 1e-5, tag + " sum evm")
            assert_close(m[:, method, 4], (ref[:, :, col + 1] ** 2).sum(1), 1e-5, tag + " sum evm^2")
        s = oracle.metrics_summary(m_gan)
        assert_close(s["evm"][:, 0], ref[:, :, 1].mean(1), 1e-5, "evm mean")
        assert_close(s["evm_std"][:, 0], ref[:, :, 1].std(1), 1e-4, "evm std")


def test_qpsk_ofdm_modulator_frames_match_reference(ref_channel):
    """QAMModulator('QPSK').modulate + OFDMModulator.modulate (ifft*N, pilots, CP), stream truncated to 16."""
    r = ref_channel
    for tag, (N, cp, sp) in (("q16", (16, 0, 8)), ("q8", (8, 2, 4)), ("q16cp", (16, 2, 16))):
        cfg = oracle.make_cfg(symbol_source=1, n_fft=N, cp_len=cp, pilot_spacing=sp, ifft_scale=1, normalize=0,
                              snr_mode=1, snr_lo=300.0, n_snr=1)     # 300 dB: noise below float32 resolution
        clean, noisy, _ = oracle.chan_sim(cfg, 32, bits=r[tag + "_words"])
        assert_close(clean, r[tag + "_frames"].astype(np.float32), 2e-7, tag + " frames")
        errs, nbits = oracle.qpsk_bit_errors(cfg, clean, r[tag + "_words"])
        assert errs == 0 and nbits == 32 * 2 * (16 // (N + cp)) * (N - len(range(0, N, sp)))
        flipped = r[tag + "_words"] ^ np.uint32(0x80000000)            # flip the first payload bit of every frame
        errs, _ = oracle.qpsk_bit_errors(cfg, clean, flipped)
        assert errs == (32 if 16 // (N + cp) else 0)                    # q16cp holds no complete symbol


def test_qpsk_table_and_tie_breaking(ref_channel):
    r = ref_channel
    tab = r["qpsk_table"]                                           # utils/ofdm_utils.py:105-109
    assert_close(tab, np.array([[1, 1, -1, -1], [1, -1, 1, -1]]) / np.sqrt(2), 1e-15, "table")
    s = r["qpsk_syms"]
    bits = np.stack([(s[0] < 0), (s[1] < 0)], axis=1).astype(np.int8).reshape(-1)   # MSB = Re<0, LSB = Im<0
    assert np.array_equal(bits, r["qpsk_bits"])                    # includes exact ties: argmin -> lowest index


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    L = oracle.lib()
    import ctypes
    out = (ctypes.c_uint32 * 4)()
    kat = [((0, 0), (0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff, 0xffffffff), (0xffffffff,) * 4, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for key, ctr, exp in kat:
        L.oracle_philox4x32_10(*[ctypes.c_uint32(v) for v in key + ctr], out)
        assert tuple(out) == exp


def test_philox_draw_statistics():
    cfg = oracle.make_cfg(nonlinear=True, snr_lo=0, snr_hi=30)
    d = oracle.frame_draws(cfg, seed=7, frame0=1000, B=4096)
    z = np.concatenate([d["sym"].ravel(), d["pn"].ravel(), d["noise"].ravel()])
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 1e-2
    assert abs(np.mean(z ** 4) - 3) < 0.1                           # kurtosis of a normal
    assert d["snr_db"].min() >= 0 and d["snr_db"].max() < 30 and abs(d["snr_db"].mean() - 15) < 0.5
    ones = np.unpackbits(d["bits"].view(np.uint8)).mean()
    assert abs(ones - 0.5) < 5e-3
    # counter-based: frame f's draws do not depend on where the batch starts
    d2 = oracle.frame_draws(cfg, seed=7, frame0=1003, B=4)
    assert np.array_equal(d2["sym"], d["sym"][3:7]) and np.array_equal(d2["bits"], d["bits"][3:7])


def test_snr_grid_and_fused_restatement(ref_fp32):
    cfg = oracle.make_cfg(nonlinear=True, pa_saturation=0.8, snr_mode=1, snr_lo=0, snr_step=5, n_snr=7,
                          frames_per_snr=100, normalize=2)
    m = oracle.sim_gen_metrics(cfg, 0, 1400, gparams=ref_fp32["gparams"], seed=3)
    assert np.all(m[:, :2, 0] == 200)                               # 1400 frames round-robin over 7 grid points
    s = oracle.metrics_summary(m)
    # NoEQ EVM improves (falls) monotonically with SNR until the non-linear floor
    assert np.all(np.diff(s["evm"][:4, 1]) < 0)
    # same frames through the unfused pieces
    clean, noisy, snr = oracle.chan_sim(cfg, 1400, seed=3)
    assert np.array_equal(np.unique(snr), np.arange(0, 35, 5, dtype=np.float32))
    bins = (np.arange(1400) // 100) % 7
    m2 = oracle.frame_metrics(oracle.gen_fwd_f32(noisy, ref_fp32["gparams"]), clean, bins.astype(np.int32), 0, 7)
    assert_close(m[:, 0, :5], m2[:, 0, :5], 1e-9, "fused vs unfused")


def test_quantize_tensor_semantics(ref_channel):
    """utils/quantization.py:73-161: half-to-even rounding, clamp, scale formula (numpy restatement)."""
    r = ref_channel
    t = r["qt_in"]
    q17 = np.clip(np.rint(t / np.float32(1.0 / 128)), -128, 127)
    assert np.array_equal(q17, r["qt_q17"])
    scale = np.float32(max(np.abs(t).max(), 1e-8) / 127)
    assert scale == r["qt_scale"]
    assert np.array_equal(np.clip(np.rint(t / scale), -128, 127), r["qt_q8"])

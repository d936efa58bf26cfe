"""TEST INFRASTRUCTURE - CPU oracle for the ofdm-gan-sr hot path (ctypes wrapper over oracle/_build/liboracle.so).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` leg may import this
package, and only as the checker / the timed CPU baseline.  Nothing under `ofdm-gan-sr_b200/` imports it; the
product path raises if its CUDA library is missing instead of falling back to this code.

C sources (each function cites the reference file:line it restates):
  fixed_point.c   Q1.7/Q8.8 integer generator, modes spec / rtl_literal        (rtl/ofdmGAN/generator_mini.v)
  fp32_models.c   MiniGenerator / MiniDiscriminator fwd+bwd, GP, critic & generator steps, Adam
  channel.c       channel simulator, metrics, Philox4x32-10 + Box-Muller, fused sim->G->metrics
rtl_cycle_emulator.py  cycle-level emulation of generator_mini.v (pins rtl_literal)

Pinning (see tests/test_oracle_*.py): 10 RTL known-answer frames from tb_generator_mini.vcd; float golden vectors
of verification_output/golden_vectors (Q8.8 truncation); fixtures recorded from the imported Python reference
(tests/golden/ref_fp32.npz, ref_channel.npz; generator script committed beside them).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

G_NP, D_NP = 258, 521
N_METHODS, METRIC_COLS = 4, 8


class ChanCfg(ctypes.Structure):
    """Mirror of `ofdmgan_chan_cfg` (include/ofdmgan.h)."""
    _fields_ = [
        ("symbol_source", ctypes.c_int32), ("n_fft", ctypes.c_int32), ("cp_len", ctypes.c_int32),
        ("pilot_spacing", ctypes.c_int32), ("pilot_re", ctypes.c_float), ("pilot_im", ctypes.c_float),
        ("ifft_scale", ctypes.c_int32), ("impair", ctypes.c_int32), ("pa_saturation", ctypes.c_float),
        ("pa_smoothness", ctypes.c_float), ("iq_gain", ctypes.c_float), ("iq_cos", ctypes.c_float),
        ("iq_sin", ctypes.c_float), ("pn_sigma", ctypes.c_float), ("snr_mode", ctypes.c_int32),
        ("snr_lo", ctypes.c_float), ("snr_hi", ctypes.c_float), ("snr_step", ctypes.c_float),
        ("n_snr", ctypes.c_int32), ("frames_per_snr", ctypes.c_int64), ("normalize", ctypes.c_int32),
        ("equalizers", ctypes.c_int32),
        ("channel_type", ctypes.c_int32), ("rician_k", ctypes.c_float), ("n_taps", ctypes.c_int32),
        ("tap_delay", ctypes.c_int32 * 4), ("tap_amp", ctypes.c_float * 4),
        ("saleh_alpha_a", ctypes.c_float), ("saleh_beta_a", ctypes.c_float), ("saleh_alpha_p", ctypes.c_float),
        ("saleh_beta_p", ctypes.c_float), ("dc_i", ctypes.c_float), ("dc_q", ctypes.c_float), ("cfo_step", ctypes.c_float),
        ("rng_rounds", ctypes.c_int32),
    ]


def make_cfg(symbol_source=0, n_fft=16, cp_len=0, pilot_spacing=0, pilot=1 + 0j, ifft_scale=0, nonlinear=False,
             pa=None, iq=None, pn=None, pa_saturation=1.0, pa_smoothness=3.0, iq_imbalance_db=1.0, iq_phase_deg=5.0,
             phase_noise_dbchz=-80.0, sample_rate=1e6, snr_mode=0, snr_lo=0.0, snr_hi=30.0, snr_step=5.0, n_snr=1,
             frames_per_snr=1, normalize=1, equalizers=False, rng_rounds=10):
    """Build a ChanCfg from the reference's user-facing parameters (SyntheticOFDMDataset.__init__,
    utils/dataset.py:195-206; NonLinearImpairments defaults, utils/ofdm_utils.py:394-521)."""
    pa = nonlinear if pa is None else pa
    iq = nonlinear if iq is None else iq
    pn = nonlinear if pn is None else pn
    c = ChanCfg()
    c.symbol_source, c.n_fft, c.cp_len, c.pilot_spacing = symbol_source, n_fft, cp_len, pilot_spacing
    c.pilot_re, c.pilot_im = float(np.real(pilot)), float(np.imag(pilot))
    c.ifft_scale = ifft_scale
    c.impair = (1 if pa else 0) | (2 if iq else 0) | (4 if pn else 0)
    c.pa_saturation, c.pa_smoothness = pa_saturation, pa_smoothness
    c.iq_gain = 10.0 ** (iq_imbalance_db / 20.0)
    phi = np.deg2rad(iq_phase_deg)
    c.iq_cos, c.iq_sin = float(np.cos(phi)), float(np.sin(phi))
    c.pn_sigma = float(np.sqrt(10.0 ** (phase_noise_dbchz / 10.0) * sample_rate))
    c.snr_mode, c.snr_lo, c.snr_hi, c.snr_step = snr_mode, snr_lo, snr_hi, snr_step
    c.n_snr, c.frames_per_snr, c.normalize = n_snr, frames_per_snr, normalize
    c.equalizers = 1 if equalizers else 0
    c.rng_rounds = int(rng_rounds)
    return c


def build(force=False):
    """Compile liboracle.so with the Makefile in this directory (gcc)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(_LIB_PATH)
            for s in ("fixed_point.c", "fp32_models.c", "channel.c", "Makefile")):
        subprocess.run(["make", "-C", _HERE, "-B"] if force else ["make", "-C", _HERE], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        try:
            build()                      # no-op when up to date; needs gcc + make
        except (OSError, subprocess.CalledProcessError):
            if not os.path.exists(_LIB_PATH):
                raise
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_snr_of_frame.restype = ctypes.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- fixed point
def rom_arrays(weights, biases):
    """{addr: value} dicts (or arrays) -> (int8[2048], int16[64]) in weight_rom.v layout."""
    W = np.zeros(2048, np.int8)
    Bq = np.zeros(64, np.int16)
    if isinstance(weights, dict):
        for k, v in weights.items():
            W[int(k)] = v
    else:
        W[:len(weights)] = weights
    if isinstance(biases, dict):
        for k, v in biases.items():
            Bq[int(k)] = v
    else:
        Bq[:len(biases)] = biases
    return W, Bq


def gen_fwd_q(x, wrom, brom, mode):
    x = np.ascontiguousarray(x, dtype=np.int16).reshape(-1, 2, 16)
    W = np.ascontiguousarray(wrom, dtype=np.int8)
    Bq = np.ascontiguousarray(brom, dtype=np.int16)
    assert W.size == 2048 and Bq.size == 64
    y = np.empty_like(x)
    rc = lib().oracle_gen_fwd_q(_p(x), _p(W), _p(Bq), _p(y), ctypes.c_int64(x.shape[0]), int(mode))
    assert rc == 0
    return y


def disc_fwd_q(cand, cond, wrom, brom, mode):
    """Fixed-point critic score per frame (int16 Q8.8): mode 0 spec, 1 rtl_literal steady state, 2 first frame after reset."""
    cand = np.ascontiguousarray(cand, dtype=np.int16).reshape(-1, 2, 16)
    cond = np.ascontiguousarray(cond, dtype=np.int16).reshape(-1, 2, 16)
    W = np.ascontiguousarray(wrom, dtype=np.int8)
    Bq = np.ascontiguousarray(brom, dtype=np.int16)
    assert W.size == 2048 and Bq.size == 64 and cand.shape == cond.shape
    score = np.empty(cand.shape[0], dtype=np.int16)
    rc = lib().oracle_disc_fwd_q(_p(cand), _p(cond), _p(W), _p(Bq), _p(score), ctypes.c_int64(cand.shape[0]), int(mode))
    assert rc == 0
    return score


def set_threads(n=None):
    """Give the C restatement n OpenMP threads (default: every core this process may run on).  torchrun exports OMP_NUM_THREADS=1
    to its workers; the CPU baseline legs of bench.py call this so that they still use the whole host.  Returns the count used."""
    n = int(n or len(os.sched_getaffinity(0)))
    return int(lib().oracle_set_threads(n))


def digest_i16(y):
    y = np.ascontiguousarray(y, dtype=np.int16)
    s, x = ctypes.c_uint64(0), ctypes.c_uint64(0)
    lib().oracle_digest_i16(_p(y), ctypes.c_int64(y.size), ctypes.byref(s), ctypes.byref(x))
    return s.value, x.value


def quantize_q88(x):
    """(x*256).astype(int16): truncation toward zero, proof/verification.py:297-298."""
    return (np.asarray(x, dtype=np.float32) * np.float32(256.0)).astype(np.int16)


# ---------------------------------------------------------------- fp32 models
def gen_fwd_f32(x, gp, slope=0.2):
    x = _f32(x).reshape(-1, 2, 16)
    y = np.empty_like(x)
    lib().oracle_gen_fwd_f32(_p(x), _p(_f32(gp)), _p(y), ctypes.c_int64(x.shape[0]), ctypes.c_float(slope))
    return y


def gen_bwd_f32(x, gp, dy, slope=0.2, need_dx=True):
    x = _f32(x).reshape(-1, 2, 16)
    dy = _f32(dy).reshape(-1, 2, 16)
    dx = np.empty_like(x) if need_dx else None
    dparams = np.empty(G_NP, np.float32)
    lib().oracle_gen_bwd_f32(_p(x), _p(_f32(gp)), _p(dy), _p(dx), _p(dparams), ctypes.c_int64(x.shape[0]),
                             ctypes.c_float(slope))
    return dx, dparams


def disc_fwd_f32(cand, cond, dp, slope=0.2):
    cand = _f32(cand).reshape(-1, 2, 16)
    cond = _f32(cond).reshape(-1, 2, 16)
    s = np.empty(cand.shape[0], np.float32)
    lib().oracle_disc_fwd_f32(_p(cand), _p(cond), _p(_f32(dp)), _p(s), ctypes.c_int64(cand.shape[0]), ctypes.c_float(slope))
    return s


def disc_bwd_f32(cand, cond, dp, g, slope=0.2):
    cand = _f32(cand).reshape(-1, 2, 16)
    cond = _f32(cond).reshape(-1, 2, 16)
    g = _f32(g).reshape(-1)
    dcand, dcond = np.empty_like(cand), np.empty_like(cond)
    grads = np.empty(D_NP, np.float32)
    lib().oracle_disc_bwd_f32(_p(cand), _p(cond), _p(_f32(dp)), _p(g), ctypes.c_float(slope), _p(dcand), _p(dcond),
                              _p(grads), ctypes.c_int64(cand.shape[0]))
    return dcand, dcond, grads


def gradient_penalty(real, fake, cond, alpha, dp, slope=0.2):
    real, fake, cond = (_f32(a).reshape(-1, 2, 16) for a in (real, fake, cond))
    alpha = _f32(alpha).reshape(-1)
    gp = ctypes.c_float(0)
    grads = np.empty(D_NP, np.float32)
    norms = np.empty(real.shape[0], np.float32)
    lib().oracle_gradient_penalty(_p(real), _p(fake), _p(cond), _p(alpha), _p(_f32(dp)), ctypes.c_float(slope),
                                  ctypes.byref(gp), _p(grads), _p(norms), ctypes.c_int64(real.shape[0]))
    return gp.value, grads, norms


def critic_step(clean, noisy, fake, alpha, dp, gp_weight=10.0, slope=0.2):
    clean, noisy, fake = (_f32(a).reshape(-1, 2, 16) for a in (clean, noisy, fake))
    alpha = _f32(alpha).reshape(-1)
    grads = np.empty(D_NP, np.float32)
    stats = np.empty(5, np.float32)
    lib().oracle_critic_step(_p(clean), _p(noisy), _p(fake), _p(alpha), _p(_f32(dp)), ctypes.c_float(gp_weight),
                             ctypes.c_float(slope), _p(grads), _p(stats), ctypes.c_int64(clean.shape[0]))
    return grads, stats


def gen_step(clean, noisy, dp, gp, adv_w=1.0, rec_w=100.0, slope=0.2):
    clean, noisy = (_f32(a).reshape(-1, 2, 16) for a in (clean, noisy))
    grads = np.empty(G_NP, np.float32)
    stats = np.empty(3, np.float32)
    fake = np.empty_like(clean)
    lib().oracle_gen_step(_p(clean), _p(noisy), _p(_f32(dp)), _p(_f32(gp)), ctypes.c_float(adv_w), ctypes.c_float(rec_w),
                          ctypes.c_float(slope), _p(grads), _p(stats), _p(fake), ctypes.c_int64(clean.shape[0]))
    return grads, stats, fake


def adam(p, m, v, g, lr, b1, b2, eps, step, grad_scale=1.0):
    p, m, v = (np.array(a, dtype=np.float32, copy=True) for a in (p, m, v))
    g = _f32(g)
    lib().oracle_adam(_p(p), _p(m), _p(v), _p(g), int(p.size), ctypes.c_double(lr), ctypes.c_double(b1), ctypes.c_double(b2),
                      ctypes.c_double(eps), int(step), ctypes.c_float(grad_scale))
    return p, m, v


# ---------------------------------------------------------------- channel / rng / metrics
def philox_blocks(seed, ctr0, c2, c3, n):
    out = np.empty((n, 4), np.uint32)
    lib().oracle_philox_blocks(ctypes.c_uint64(seed), ctypes.c_uint64(ctr0), ctypes.c_uint32(c2), ctypes.c_uint32(c3),
                               _p(out), ctypes.c_int64(n))
    return out


def frame_draws(cfg, seed, frame0, B):
    """The Philox draws frames frame0..frame0+B-1 consume: dict(sym[B,32], bits[B], pn[B,16], snr_db[B], noise[B,32])."""
    sym = np.empty((B, 32)); pn = np.empty((B, 16)); noise = np.empty((B, 32)); snr = np.empty(B)
    bits = np.empty(B, np.uint32)
    L = lib()
    for b in range(B):
        bw, sv = ctypes.c_uint32(0), ctypes.c_double(0)
        L.oracle_frame_draws(ctypes.byref(cfg), ctypes.c_uint64(seed), ctypes.c_uint64(frame0 + b),
                             ctypes.c_void_p(sym[b].ctypes.data), ctypes.byref(bw), ctypes.c_void_p(pn[b].ctypes.data),
                             ctypes.byref(sv), ctypes.c_void_p(noise[b].ctypes.data))
        bits[b], snr[b] = bw.value, sv.value
    return dict(sym=sym, bits=bits, pn=pn, snr_db=snr, noise=noise)


def chan_sim(cfg, B, seed=0, frame0=0, sym=None, bits=None, pn=None, snr_db=None, noise=None):
    """-> clean[B,2,16] f32, noisy[B,2,16] f32, snr[B] f32.  Any of the draw arrays may be injected (float64)."""
    sym, pn, snr_db, noise = _f64(sym), _f64(pn), _f64(snr_db), _f64(noise)
    bits = None if bits is None else np.ascontiguousarray(bits, dtype=np.uint32)
    clean = np.empty((B, 2, 16), np.float32)
    noisy = np.empty((B, 2, 16), np.float32)
    snr = np.empty(B, np.float32)
    rc = lib().oracle_chan_sim(ctypes.byref(cfg), _p(sym), _p(bits), _p(pn), _p(snr_db), _p(noise), ctypes.c_uint64(seed),
                               ctypes.c_uint64(frame0), _p(clean), _p(noisy), _p(snr), ctypes.c_int64(B))
    assert rc == 0
    return clean, noisy, snr


def frame_metrics(est, ref, bins=None, method=0, n_snr=1):
    est, ref = (_f32(a).reshape(-1, 2, 16) for a in (est, ref))
    bins = None if bins is None else np.ascontiguousarray(bins, dtype=np.int32)
    m = np.zeros((n_snr, N_METHODS, METRIC_COLS))
    rc = lib().oracle_frame_metrics(_p(est), _p(ref), _p(bins), int(method), int(n_snr), ctypes.c_int64(est.shape[0]), _p(m))
    assert rc == 0
    return m


def sim_gen_metrics(cfg, gen_kind, B, gparams=None, wrom=None, brom=None, slope=0.2, seed=0, frame0=0):
    n_snr = cfg.n_snr if cfg.snr_mode == 1 else 1
    m = np.zeros((n_snr, N_METHODS, METRIC_COLS))
    gparams = _f32(gparams)
    wrom = None if wrom is None else np.ascontiguousarray(wrom, dtype=np.int8)
    brom = None if brom is None else np.ascontiguousarray(brom, dtype=np.int16)
    rc = lib().oracle_sim_gen_metrics(ctypes.byref(cfg), int(gen_kind), _p(gparams), _p(wrom), _p(brom), ctypes.c_float(slope),
                                      ctypes.c_uint64(seed), ctypes.c_uint64(frame0), ctypes.c_int64(B), _p(m))
    assert rc == 0
    return m


def equalize(noisy, clean, snr_db=None, method=2):
    """ZF (method 2) / MMSE (method 3) with the genie channel estimate -> equalised frames [B,2,16] float32."""
    noisy, clean = (_f32(a).reshape(-1, 2, 16) for a in (noisy, clean))
    snr = None if snr_db is None else _f32(np.broadcast_to(np.asarray(snr_db, dtype=np.float32), (noisy.shape[0],)))
    est = np.empty_like(noisy)
    rc = lib().oracle_equalize(_p(noisy), _p(clean), _p(snr), int(method), _p(est), ctypes.c_int64(noisy.shape[0]))
    assert rc == 0
    return est


def qpsk_bit_errors(cfg, frames, bits):
    frames = _f32(frames).reshape(-1, 2, 16)
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    e, n = ctypes.c_int64(0), ctypes.c_int64(0)
    lib().oracle_qpsk_bit_errors(ctypes.byref(cfg), _p(frames), _p(bits), ctypes.c_int64(frames.shape[0]), ctypes.byref(e), ctypes.byref(n))
    return e.value, n.value


def metrics_summary(m):
    """accumulator rows -> dict(mean/std of mse and evm_dB, ber) with np.mean/np.std (population) semantics,
    benchmark_comparison.py:253-259."""
    m = np.asarray(m, dtype=np.float64)
    n = np.maximum(m[..., 0], 1.0)
    mse, evm = m[..., 1] / n, m[..., 3] / n
    return dict(n=m[..., 0], mse=mse, mse_std=np.sqrt(np.maximum(m[..., 2] / n - mse ** 2, 0.0)), evm=evm,
                evm_std=np.sqrt(np.maximum(m[..., 4] / n - evm ** 2, 0.0)),
                ber=m[..., 5] / np.maximum(m[..., 6], 1.0))

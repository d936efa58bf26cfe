"""Functional layer over the C ABI (include/ofdmgan.h): torch CUDA tensors in, torch CUDA tensors out, every call
asynchronous on torch's current stream.  No CPU path - a CPU tensor raises OfdmGanError."""
import ctypes
import math
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import (CRITIC_OUT, D_NPARAMS, G_NPARAMS, GEN_F32, GEN_OUT, GEN_Q_RTL, GEN_Q_SPEC, METRIC_COLS, N_METHODS,
                   ChanCfg, ChanRand, OfdmGanError, check, dptr, frames, stream_ptr)

SYM_GAUSSIAN, SYM_QPSK = 0, 1
SCALE_SQRT_N, SCALE_N = 0, 1
IMPAIR_PA, IMPAIR_IQ, IMPAIR_PN, IMPAIR_SALEH, IMPAIR_DC, IMPAIR_CFO = 1, 2, 4, 8, 16, 32
CHAN_AWGN, CHAN_RAYLEIGH, CHAN_RICIAN, CHAN_MULTIPATH = 0, 1, 2, 3
CHANNEL_TYPES = {"awgn": 0, "rayleigh": 1, "rician": 2, "multipath": 3}
SNR_UNIFORM, SNR_GRID, SNR_NONE = 0, 1, 2
NORM_NONE, NORM_JOINT, NORM_SEPARATE = 0, 1, 2


def make_cfg(symbol_source=SYM_GAUSSIAN, n_fft=16, cp_len=0, pilot_spacing=0, pilot=1 + 0j, ifft_scale=SCALE_SQRT_N,
             nonlinear=False, pa=None, iq=None, pn=None, pa_saturation=1.0, pa_smoothness=3.0, iq_imbalance_db=1.0,
             iq_phase_deg=5.0, phase_noise_dbchz=-80.0, sample_rate=1e6, snr_mode=SNR_UNIFORM, snr_lo=0.0, snr_hi=30.0,
             snr_step=5.0, n_snr=1, frames_per_snr=1, normalize=NORM_JOINT, equalizers=False, channel_type=CHAN_AWGN,
             rician_k=3.0, delays=(0, 1, 2), powers=(1.0, 0.5, 0.25), saleh=None, dc_offset=None, cfo_hz=None, rng_rounds=10):
    """ChanCfg from the reference's user-facing parameters: SyntheticOFDMDataset.__init__ (utils/dataset.py:195-206),
    NonLinearImpairments defaults (utils/ofdm_utils.py:394-521), run_benchmark's SNR grid
    (benchmark_comparison.py:179-182)."""
    pa = nonlinear if pa is None else pa
    iq = nonlinear if iq is None else iq
    pn = nonlinear if pn is None else pn
    c = ChanCfg()
    c.symbol_source, c.n_fft, c.cp_len, c.pilot_spacing = symbol_source, n_fft, cp_len, pilot_spacing
    c.pilot_re, c.pilot_im = float(np.real(pilot)), float(np.imag(pilot))
    c.ifft_scale = ifft_scale
    c.impair = (IMPAIR_PA if pa else 0) | (IMPAIR_IQ if iq else 0) | (IMPAIR_PN if pn else 0)
    c.pa_saturation, c.pa_smoothness = pa_saturation, pa_smoothness
    c.iq_gain = 10.0 ** (iq_imbalance_db / 20.0)
    phi = math.radians(iq_phase_deg)
    c.iq_cos, c.iq_sin = math.cos(phi), math.sin(phi)
    c.pn_sigma = math.sqrt(10.0 ** (phase_noise_dbchz / 10.0) * sample_rate)
    c.snr_mode, c.snr_lo, c.snr_hi, c.snr_step = snr_mode, snr_lo, snr_hi, snr_step
    c.n_snr, c.frames_per_snr, c.normalize = n_snr, frames_per_snr, normalize
    c.equalizers = 1 if equalizers else 0        # also fill the ZF / MMSE rows of the fused sweep
    # fading channels (ChannelModel, utils/ofdm_utils.py:710-832) and the remaining impairments (:424-455, :524-568)
    c.channel_type = CHANNEL_TYPES[channel_type] if isinstance(channel_type, str) else int(channel_type)
    c.rician_k = float(rician_k)
    if len(delays) != len(powers) or len(delays) > 4:
        raise OfdmGanError("multipath: up to 4 taps, one power per delay")
    c.n_taps = len(delays)
    tot = float(sum(powers))
    for t, (d, p) in enumerate(zip(delays, powers)):
        c.tap_delay[t], c.tap_amp[t] = int(d), math.sqrt(p / tot)
    if saleh is not None:                          # (alpha_a, beta_a, alpha_p, beta_p)
        c.impair |= IMPAIR_SALEH
        c.saleh_alpha_a, c.saleh_beta_a, c.saleh_alpha_p, c.saleh_beta_p = (float(v) for v in saleh)
    if dc_offset is not None:                      # (dc_i, dc_q) relative to sqrt(mean |x|^2)
        c.impair |= IMPAIR_DC
        c.dc_i, c.dc_q = float(dc_offset[0]), float(dc_offset[1])
    if cfo_hz is not None:
        c.impair |= IMPAIR_CFO
        c.cfo_step = 2.0 * math.pi * float(cfo_hz) / float(sample_rate)
    c.rng_rounds = int(rng_rounds)         # 10: Philox4x32-10 (default, the parity workload); 7: the "fast RNG" workload
    return c


def _params(t, n, name):
    """flat fp32 parameter vector: CUDA tensor (device pointer) or CPU tensor / ndarray (host pointer; the library
    accepts either).  Returns (keepalive, c_void_p)."""
    if isinstance(t, torch.Tensor):
        t = t.detach()
        if t.numel() != n:
            raise OfdmGanError(f"{name}: expected {n} values, got {t.numel()}")
        t = t.to(torch.float32).contiguous().view(-1)
        return t, ctypes.c_void_p(t.data_ptr())
    a = np.ascontiguousarray(t, dtype=np.float32).reshape(-1)
    if a.size != n:
        raise OfdmGanError(f"{name}: expected {n} values, got {a.size}")
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _roms(wrom, brom):
    W = np.ascontiguousarray(wrom, dtype=np.int8).reshape(-1)
    Bq = np.ascontiguousarray(brom, dtype=np.int16).reshape(-1)
    if W.size != 2048 or Bq.size != 64:
        raise OfdmGanError("ROMs must be int8[2048] and int16[64] (weight_rom.v layout)")
    return W, Bq


def flatten_params(module):
    """parameters() order, flattened (the packing of include/ofdmgan.h)."""
    return torch.cat([p.detach().reshape(-1) for p in module.parameters()]).to(torch.float32).contiguous()


_FLAT_CACHE = {}


def flat_cached(params):
    """The flat fp32 vector of a parameter list, rebuilt only when a parameter changed (torch bumps a tensor's `_version` on every
    in-place write, e.g. optimizer.step()): the module path of the fused forward / backward then launches no `cat` kernel per call.
    An entry belongs to the tensor OBJECTS it was built from (held weakly): a new module whose parameters were allocated at the
    addresses of a freed one, with the same version counters, must not see the old module's weights."""
    params = tuple(params)
    key = tuple((p.data_ptr(), p._version) for p in params)
    ident = tuple(k[0] for k in key)
    hit = _FLAT_CACHE.get(ident)
    if hit is not None and hit[0] == key and len(hit[2]) == len(params) and all(r() is p for r, p in zip(hit[2], params)):
        return hit[1]
    flat = torch.cat([p.detach().reshape(-1) for p in params]).to(torch.float32).contiguous()
    if len(_FLAT_CACHE) > 16:
        _FLAT_CACHE.clear()
    _FLAT_CACHE[ident] = (key, flat, tuple(weakref.ref(p) for p in params))
    return flat


# ---------------------------------------------------------------------------------------------- generator
def gen_fwd_f32(x, gparams, slope=0.2):
    x = frames(x)
    y = torch.empty_like(x)
    keep, gp = _params(gparams, G_NPARAMS, "gparams")
    check(_lib.lib().ofdmgan_gen_fwd_f32(dptr(x), gp, dptr(y), x.shape[0], slope, stream_ptr(x.device)))
    return y


def gen_bwd_f32(x, gparams, dy, slope=0.2, need_dx=True):
    x, dy = frames(x), frames(dy)
    dx = torch.empty_like(x) if need_dx else None
    dparams = torch.empty(G_NPARAMS, dtype=torch.float32, device=x.device)
    keep, gp = _params(gparams, G_NPARAMS, "gparams")
    check(_lib.lib().ofdmgan_gen_bwd_f32(dptr(x), gp, dptr(dy), dptr(dx), dptr(dparams), x.shape[0], slope,
                                         stream_ptr(x.device)))
    return dx, dparams


def gen_fwd_q(x_q88, wrom, brom, mode=GEN_Q_SPEC, want_digest=False):
    x = frames(x_q88, torch.int16)
    y = torch.empty_like(x)
    W, Bq = _roms(wrom, brom)
    digest = torch.zeros(2, dtype=torch.int64, device=x.device) if want_digest else None
    check(_lib.lib().ofdmgan_gen_fwd_q(dptr(x), W.ctypes.data_as(ctypes.c_void_p), Bq.ctypes.data_as(ctypes.c_void_p),
                                       dptr(y), x.shape[0], mode, dptr(digest), stream_ptr(x.device)))
    if want_digest:
        d = digest.cpu().numpy().view(np.uint64)
        return y, (int(d[0]), int(d[1]))
    return y


def disc_fwd_q(cand_q88, cond_q88, wrom, brom, mode=GEN_Q_SPEC):
    """Integer critic score per frame ([B] int16, Q8.8): rtl/ofdmGAN/discriminator_mini.v, mode as for gen_fwd_q."""
    a, c = frames(cand_q88, torch.int16), frames(cond_q88, torch.int16)
    if a.shape != c.shape or a.device != c.device:
        raise OfdmGanError("candidate and condition must have the same shape and device")
    W, Bq = _roms(wrom, brom)
    score = torch.empty(a.shape[0], dtype=torch.int16, device=a.device)
    check(_lib.lib().ofdmgan_disc_fwd_q(dptr(a), dptr(c), W.ctypes.data_as(ctypes.c_void_p), Bq.ctypes.data_as(ctypes.c_void_p),
                                        dptr(score), a.shape[0], mode, stream_ptr(a.device)))
    return score


def compute_scale(x2d, n_bits):
    """x2d: [C, inner] float32 CUDA -> [C] scales (utils/quantization.py:73-112)."""
    x2d = x2d.to(torch.float32).contiguous()
    scale = torch.empty(x2d.shape[0], dtype=torch.float32, device=x2d.device)
    check(_lib.lib().ofdmgan_compute_scale(dptr(x2d), x2d.shape[0], x2d.shape[1], n_bits, dptr(scale), stream_ptr(x2d.device)))
    return scale


def quantize_tensor(x2d, scale, n_bits):
    x2d, scale = x2d.to(torch.float32).contiguous(), scale.to(torch.float32).contiguous().view(-1)
    q = torch.empty_like(x2d)
    check(_lib.lib().ofdmgan_quantize_tensor(dptr(x2d), x2d.shape[0], x2d.shape[1], dptr(scale), n_bits, dptr(q), stream_ptr(x2d.device)))
    return q


def dequantize_tensor(q2d, scale):
    q2d, scale = q2d.to(torch.float32).contiguous(), scale.to(torch.float32).contiguous().view(-1)
    x = torch.empty_like(q2d)
    check(_lib.lib().ofdmgan_dequantize_tensor(dptr(q2d), q2d.shape[0], q2d.shape[1], dptr(scale), dptr(x), stream_ptr(q2d.device)))
    return x


def quantize_q88(x):
    if not x.is_cuda:
        raise OfdmGanError("quantize_q88 needs a CUDA tensor")
    x = x.to(torch.float32).contiguous()
    q = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    check(_lib.lib().ofdmgan_quantize_q88(dptr(x), dptr(q), x.numel(), stream_ptr(x.device)))
    return q


def dequantize_q88(q):
    if not q.is_cuda:
        raise OfdmGanError("dequantize_q88 needs a CUDA tensor")
    q = q.to(torch.int16).contiguous()
    x = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    check(_lib.lib().ofdmgan_dequantize_q88(dptr(q), dptr(x), q.numel(), stream_ptr(q.device)))
    return x


# ---------------------------------------------------------------------------------------------- channel
def _dev(device):
    device = torch.device("cuda" if device is None else device)
    if device.type != "cuda":
        raise OfdmGanError("libofdmgan has no CPU path: a CUDA device is required")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def chan_sim(cfg, B, seed=0, frame0=0, device=None, sym=None, bits=None, pn=None, snr_db=None, noise=None, tx=None, fade=None,
             tx_gain=None, want_clean=True, want_noisy=True, want_snr=True):
    """frames frame0..frame0+B-1 of SyntheticOFDMDataset / run_benchmark -> (clean, noisy, snr) CUDA tensors.
    Any of the draw tensors may be injected (host-generated randomness in the reference's draw order)."""
    device = _dev(device)
    with torch.cuda.device(device):
        clean = torch.empty(B, 2, 16, dtype=torch.float32, device=device) if want_clean else None
        noisy = torch.empty(B, 2, 16, dtype=torch.float32, device=device) if want_noisy else None
        snr = torch.empty(B, dtype=torch.float32, device=device) if want_snr else None
        rand = None
        keep = []
        if tx_gain is not None and tx is None:
            raise OfdmGanError("tx_gain scales an injected time-domain frame: pass tx as well")
        if any(a is not None for a in (sym, bits, pn, snr_db, noise, tx, fade)):
            def f32(a, shape):
                if a is None:
                    return None
                t = torch.as_tensor(a).to(device=device, dtype=torch.float32).contiguous().view(*shape)
                keep.append(t)
                return t.data_ptr()
            rand = ChanRand()
            rand.sym = f32(sym, (B, 32))
            rand.pn = f32(pn, (B, 16))
            rand.snr_db = f32(snr_db, (B,))
            rand.noise = f32(noise, (B, 32))
            rand.tx = f32(tx, (B, 32))
            rand.fade = f32(fade, (B, 8))
            rand.tx_gain = f32(tx_gain, (B,))
            if bits is not None:
                tb = torch.as_tensor(np.ascontiguousarray(bits, dtype=np.uint32).view(np.int32)).to(device).contiguous()
                keep.append(tb)
                rand.bits = tb.data_ptr()
        check(_lib.lib().ofdmgan_chan_sim(ctypes.byref(cfg), ctypes.byref(rand) if rand is not None else None, seed, frame0,
                                          dptr(clean), dptr(noisy), dptr(snr), B, stream_ptr(device)))
        if keep:
            torch.cuda.current_stream(device).synchronize()     # the injected buffers must outlive the launch
    return clean, noisy, snr


def chan_draws(cfg, B, seed=0, frame0=0, device=None):
    device = _dev(device)
    with torch.cuda.device(device):
        sym = torch.empty(B, 32, dtype=torch.float32, device=device)
        pn = torch.empty(B, 16, dtype=torch.float32, device=device)
        noise = torch.empty(B, 32, dtype=torch.float32, device=device)
        snr = torch.empty(B, dtype=torch.float32, device=device)
        bits = torch.empty(B, dtype=torch.int32, device=device)
        check(_lib.lib().ofdmgan_chan_draws(ctypes.byref(cfg), seed, frame0, dptr(sym), dptr(bits), dptr(pn), dptr(snr),
                                            dptr(noise), B, stream_ptr(device)))
    return dict(sym=sym, bits=bits, pn=pn, snr_db=snr, noise=noise)


def chan_fade_draws(cfg, B, seed=0, frame0=0, device=None):
    """[B,8] fading draws of frames frame0.. in the ofdmgan_chan_rand.fade layout."""
    device = _dev(device)
    with torch.cuda.device(device):
        fade = torch.empty(B, 8, dtype=torch.float32, device=device)
        check(_lib.lib().ofdmgan_chan_fade_draws(ctypes.byref(cfg), seed, frame0, dptr(fade), B, stream_ptr(device)))
    return fade


def philox_blocks(seed, ctr0, c2, c3, n, device=None):
    device = _dev(device)
    with torch.cuda.device(device):
        out = torch.empty(n, 4, dtype=torch.int32, device=device)
        check(_lib.lib().ofdmgan_philox_blocks(seed, ctr0, c2, c3, dptr(out), n, stream_ptr(device)))
    return out


def n_snr_of(cfg):
    return cfg.n_snr if cfg.snr_mode == SNR_GRID else 1


def sim_gen_metrics(cfg, B, gen_kind=GEN_F32, gparams=None, wrom=None, brom=None, slope=0.2, seed=0, frame0=0, device=None,
                    out=None):
    """Fused simulate -> reconstruct -> metrics for frames frame0..frame0+B-1; ADDS into `out`
    (double [n_snr][N_METHODS][METRIC_COLS] on the device) and returns it."""
    device = _dev(device)
    with torch.cuda.device(device):
        if out is None:
            out = torch.zeros(n_snr_of(cfg), N_METHODS, METRIC_COLS, dtype=torch.float64, device=device)
        keep, gp, W, Bq = None, None, None, None
        if gen_kind == GEN_F32:
            keep, gp = _params(gparams, G_NPARAMS, "gparams")
        else:
            W, Bq = _roms(wrom, brom)
        check(_lib.lib().ofdmgan_sim_gen_metrics(
            ctypes.byref(cfg), gen_kind, gp, None if W is None else W.ctypes.data_as(ctypes.c_void_p),
            None if Bq is None else Bq.ctypes.data_as(ctypes.c_void_p), slope, seed, frame0, B, dptr(out), stream_ptr(device)))
    return out


def sim_gen_metrics_host(cfg, B, gen_kind=GEN_F32, gparams=None, wrom=None, brom=None, slope=0.2, seed=0, frame0=0,
                         device=None, out=None):
    """Same through host buffers (weights in, metric table out, synchronous): the end-to-end call."""
    device = _dev(device)
    if out is None:
        out = np.zeros((n_snr_of(cfg), N_METHODS, METRIC_COLS), dtype=np.float64)
    gp = W = Bq = None
    if gen_kind == GEN_F32:
        gp = np.ascontiguousarray(gparams.detach().cpu().numpy() if isinstance(gparams, torch.Tensor) else gparams,
                                  dtype=np.float32).reshape(-1)
        if gp.size != G_NPARAMS:
            raise OfdmGanError("gparams: expected 258 values")
    else:
        W, Bq = _roms(wrom, brom)
    vp = ctypes.c_void_p
    with torch.cuda.device(device):
        check(_lib.lib().ofdmgan_sim_gen_metrics_host(
            ctypes.byref(cfg), gen_kind, None if gp is None else gp.ctypes.data_as(vp),
            None if W is None else W.ctypes.data_as(vp), None if Bq is None else Bq.ctypes.data_as(vp), slope, seed, frame0, B,
            out.ctypes.data_as(vp), stream_ptr(device)))
    return out


def frame_metrics(est, ref, bins=None, method=0, n_snr=1, out=None):
    est, ref = frames(est), frames(ref)
    if out is None:
        out = torch.zeros(n_snr, N_METHODS, METRIC_COLS, dtype=torch.float64, device=est.device)
    if bins is not None:
        bins = bins.to(device=est.device, dtype=torch.int32).contiguous()
    check(_lib.lib().ofdmgan_frame_metrics(dptr(est), dptr(ref), dptr(bins), method, n_snr, est.shape[0], dptr(out),
                                           stream_ptr(est.device)))
    return out


def equalize(noisy, clean, method, snr_db=None):
    """Genie-aided ZF (METHOD_ZF) / MMSE (METHOD_MMSE) on [B,2,16] frames -> equalised frames
    (ZeroForcingEqualizer / MMSEEqualizer.equalize_iq, utils/classical_equalizers.py:88-126,203-230)."""
    noisy, clean = frames(noisy), frames(clean)
    est = torch.empty_like(noisy)
    snr = None
    if snr_db is not None:
        snr = torch.as_tensor(snr_db, dtype=torch.float32, device=noisy.device).reshape(-1)
        snr = snr.expand(noisy.shape[0]).contiguous() if snr.numel() == 1 else snr.contiguous()
    check(_lib.lib().ofdmgan_equalize(dptr(noisy), dptr(clean), dptr(snr), method, dptr(est), noisy.shape[0], stream_ptr(noisy.device)))
    return est


def metrics_summary(m):
    """accumulator rows -> mean/std of MSE and EVM(dB) with np.mean / np.std (population) semantics
    (benchmark_comparison.py:253-259) and BER."""
    m = np.asarray(m.cpu() if isinstance(m, torch.Tensor) else m, dtype=np.float64)
    n = np.maximum(m[..., 0], 1.0)
    mse, evm = m[..., 1] / n, m[..., 3] / n
    return dict(n=m[..., 0], mse=mse, mse_std=np.sqrt(np.maximum(m[..., 2] / n - mse ** 2, 0.0)), evm=evm,
                evm_std=np.sqrt(np.maximum(m[..., 4] / n - evm ** 2, 0.0)), ber=m[..., 5] / np.maximum(m[..., 6], 1.0))


# ---------------------------------------------------------------------------------------------- modulators
def _c64(t, device=None):
    """complex64 contiguous CUDA tensor (view_as_real gives the interleaved float layout the C ABI takes)."""
    t = torch.as_tensor(t)
    if device is None and not t.is_cuda:
        raise OfdmGanError("expected a CUDA tensor: libofdmgan has no CPU path")
    return t.to(device=device if device is not None else t.device, dtype=torch.complex64).contiguous()


def unpackbits(bytes_):
    """uint8 CUDA tensor -> uint8 0/1 tensor, 8 per byte, MSB first (np.unpackbits)."""
    b = bytes_.to(torch.uint8).contiguous().view(-1)
    if not b.is_cuda:
        raise OfdmGanError("expected a CUDA tensor: libofdmgan has no CPU path")
    bits = torch.empty(b.numel() * 8, dtype=torch.uint8, device=b.device)
    check(_lib.lib().ofdmgan_unpackbits(dptr(b), b.numel(), dptr(bits), stream_ptr(b.device)))
    return bits


def packbits(bits):
    """0/1 uint8 CUDA tensor (a multiple of 8 long) -> bytes, MSB first (np.packbits)."""
    b = bits.to(torch.uint8).contiguous().view(-1)
    if not b.is_cuda or b.numel() % 8:
        raise OfdmGanError("packbits needs a CUDA tensor whose length is a multiple of 8")
    out = torch.empty(b.numel() // 8, dtype=torch.uint8, device=b.device)
    check(_lib.lib().ofdmgan_packbits(dptr(b), out.numel(), dptr(out), stream_ptr(b.device)))
    return out


def qam_modulate(bits, bits_per_symbol):
    """bits: uint8 CUDA tensor of 0/1 (MSB first) -> complex64 symbols; bits_per_symbol 2 (QPSK), 4 (QAM16) or 6 (QAM64)."""
    if not bits.is_cuda:
        raise OfdmGanError("expected a CUDA tensor: libofdmgan has no CPU path")
    bits = bits.to(torch.uint8).contiguous().view(-1)
    n = bits.numel() // bits_per_symbol
    sym = torch.empty(n, dtype=torch.complex64, device=bits.device)
    check(_lib.lib().ofdmgan_qam_modulate(dptr(bits), ctypes.c_void_p(sym.data_ptr()), n, bits_per_symbol, stream_ptr(bits.device)))
    return sym


def qam_demodulate(symbols, bits_per_symbol):
    sym = _c64(symbols).view(-1)
    bits = torch.empty(bits_per_symbol * sym.numel(), dtype=torch.uint8, device=sym.device)
    check(_lib.lib().ofdmgan_qam_demodulate(ctypes.c_void_p(sym.data_ptr()), dptr(bits), sym.numel(), bits_per_symbol, stream_ptr(sym.device)))
    return bits


def qpsk_modulate(bits):
    """bits: uint8 CUDA tensor of 0/1, length 2n (MSB first) -> complex64[n] (QAMModulator('QPSK').modulate)."""
    if not bits.is_cuda:
        raise OfdmGanError("expected a CUDA tensor: libofdmgan has no CPU path")
    bits = bits.to(torch.uint8).contiguous().view(-1)
    n = bits.numel() // 2
    sym = torch.empty(n, dtype=torch.complex64, device=bits.device)
    check(_lib.lib().ofdmgan_qpsk_modulate(dptr(bits), ctypes.c_void_p(sym.data_ptr()), n, stream_ptr(bits.device)))
    return sym


def qpsk_demodulate(symbols):
    sym = _c64(symbols).view(-1)
    bits = torch.empty(2 * sym.numel(), dtype=torch.uint8, device=sym.device)
    check(_lib.lib().ofdmgan_qpsk_demodulate(ctypes.c_void_p(sym.data_ptr()), dptr(bits), sym.numel(), stream_ptr(sym.device)))
    return bits


def _n_pilots(n_fft, spacing):
    return (n_fft + spacing - 1) // spacing if spacing > 0 else 0


def ofdm_modulate(symbols, n_fft, cp_len, pilot_spacing, pilot=1 + 0j):
    sym = _c64(symbols).view(-1)
    n_data = n_fft - _n_pilots(n_fft, pilot_spacing)
    n_ofdm = -(-sym.numel() // max(n_data, 1))
    cp_eff = cp_len if cp_len > 0 else n_fft                    # the reference's [-0:] slice duplicates the whole symbol
    out = torch.empty(n_ofdm * (n_fft + cp_eff), dtype=torch.complex64, device=sym.device)
    check(_lib.lib().ofdmgan_ofdm_modulate(ctypes.c_void_p(sym.data_ptr()), sym.numel(), n_fft, cp_len, pilot_spacing, float(np.real(pilot)),
                                           float(np.imag(pilot)), ctypes.c_void_p(out.data_ptr()), stream_ptr(sym.device)))
    return out


def ofdm_demodulate(signal, n_fft, cp_len, pilot_spacing, pilot=1 + 0j):
    sig = _c64(signal).view(-1)
    n_ofdm = sig.numel() // (n_fft + cp_len)
    n_pil = _n_pilots(n_fft, pilot_spacing)
    data = torch.empty(n_ofdm * (n_fft - n_pil), dtype=torch.complex64, device=sig.device)
    chan = torch.empty(n_ofdm, n_pil, dtype=torch.complex64, device=sig.device)
    check(_lib.lib().ofdmgan_ofdm_demodulate(ctypes.c_void_p(sig.data_ptr()), n_ofdm, n_fft, cp_len, pilot_spacing, float(np.real(pilot)),
                                             float(np.imag(pilot)), ctypes.c_void_p(data.data_ptr()),
                                             ctypes.c_void_p(chan.data_ptr()) if n_pil else None, stream_ptr(sig.device)))
    return data, chan


# ---------------------------------------------------------------------------------------------- critic
def disc_fwd_f32(cand, cond, dparams, slope=0.2):
    cand, cond = frames(cand), frames(cond)
    score = torch.empty(cand.shape[0], dtype=torch.float32, device=cand.device)
    keep, dp = _params(dparams, D_NPARAMS, "dparams")
    check(_lib.lib().ofdmgan_disc_fwd_f32(dptr(cand), dptr(cond), dp, dptr(score), cand.shape[0], slope, stream_ptr(cand.device)))
    return score


def disc_bwd_f32(cand, cond, dparams, g, slope=0.2, need_dcand=True, need_dcond=True, need_dparams=True):
    cand, cond = frames(cand), frames(cond)
    g = g.to(torch.float32).contiguous().view(-1)
    dcand = torch.empty_like(cand) if need_dcand else None
    dcond = torch.empty_like(cond) if need_dcond else None
    grads = torch.empty(D_NPARAMS, dtype=torch.float32, device=cand.device) if need_dparams else None
    keep, dp = _params(dparams, D_NPARAMS, "dparams")
    check(_lib.lib().ofdmgan_disc_bwd_f32(dptr(cand), dptr(cond), dp, dptr(g), dptr(dcand), dptr(dcond), dptr(grads),
                                          cand.shape[0], slope, stream_ptr(cand.device)))
    return dcand, dcond, grads


def gradient_penalty(real, fake, cond, dparams, alpha=None, seed=0, sample0=0, alpha_iter=0, slope=0.2, need_dparams=True):
    real, fake, cond = frames(real), frames(fake), frames(cond)
    if alpha is not None:
        alpha = alpha.to(torch.float32).contiguous().view(-1)
    gp = torch.empty(1, dtype=torch.float32, device=real.device)
    grads = torch.empty(D_NPARAMS, dtype=torch.float32, device=real.device) if need_dparams else None
    keep, dp = _params(dparams, D_NPARAMS, "dparams")
    check(_lib.lib().ofdmgan_gradient_penalty(dptr(real), dptr(fake), dptr(cond), dptr(alpha), seed, sample0, alpha_iter, dp,
                                              dptr(gp), dptr(grads), real.shape[0], slope, stream_ptr(real.device)))
    return gp, grads


def critic_step(clean, noisy, fake, dparams, alpha=None, seed=0, sample0=0, alpha_iter=0, gp_weight=10.0, slope=0.2,
                b_global=None, out=None, alpha_iter_dev=None):
    """-> out[528] = grad[521] (local sum / B_global), stats[5] (d_loss, wasserstein, gp, d_real, d_fake), 2 pad.
    alpha_iter_dev (int32 CUDA tensor, 1 element): take the Philox alpha counter from device memory (graph-replayable)."""
    clean, noisy, fake = frames(clean), frames(noisy), frames(fake)
    if alpha_iter_dev is not None:
        if alpha is not None:
            raise OfdmGanError("alpha_iter_dev draws alpha from Philox; an injected alpha cannot be combined with it")
        if out is None:
            out = torch.empty(CRITIC_OUT, dtype=torch.float32, device=clean.device)
        B = clean.shape[0]
        keep, dp = _params(dparams, D_NPARAMS, "dparams")
        check(_lib.lib().ofdmgan_critic_step_ctr(dptr(clean), dptr(noisy), dptr(fake), seed, sample0, dptr(alpha_iter_dev), dp, gp_weight,
                                                 slope, B, B if b_global is None else b_global, dptr(out), stream_ptr(clean.device)))
        return out
    if alpha is not None:
        alpha = alpha.to(torch.float32).contiguous().view(-1)
    if out is None:
        out = torch.empty(CRITIC_OUT, dtype=torch.float32, device=clean.device)
    B = clean.shape[0]
    keep, dp = _params(dparams, D_NPARAMS, "dparams")
    check(_lib.lib().ofdmgan_critic_step(dptr(clean), dptr(noisy), dptr(fake), dptr(alpha), seed, sample0, alpha_iter, dp,
                                         gp_weight, slope, B, B if b_global is None else b_global, dptr(out),
                                         stream_ptr(clean.device)))
    return out


def critic_train(clean, noisy, fake, dparams, m, v, step_dev, lr, beta1, beta2, eps, seed=0, sample0=0, gp_weight=10.0, slope=0.2, out=None,
                 image_is_current=False, comm=None, b_global=None):
    """One whole critic iteration (loss, backward, gradient sum over the ranks of `comm` if given, Adam in place on dparams / m / v,
    weight-image refresh): two launches.  step_dev: int32 CUDA tensor (1 element) = optimiser steps taken so far; doubles as the
    Philox alpha counter."""
    clean, noisy, fake = frames(clean), frames(noisy), frames(fake)
    if out is None:
        out = torch.empty(CRITIC_OUT, dtype=torch.float32, device=clean.device)
    check(_lib.lib().ofdmgan_critic_train_ctr(dptr(clean), dptr(noisy), dptr(fake), seed, sample0, dptr(step_dev), dptr(dparams), dptr(m),
                                              dptr(v), lr, beta1, beta2, eps, gp_weight, slope, clean.shape[0],
                                              clean.shape[0] if b_global is None else b_global, dptr(out),
                                              1 if image_is_current else 0, comm._h if comm is not None else None,
                                              stream_ptr(clean.device)))
    return out


def gen_step(clean, noisy, dparams, gparams, adv_weight=1.0, rec_weight=100.0, slope=0.2, b_global=None, out=None,
             fake_out=None, fake=None):
    """-> out[264] = grad[258] (local sum / B_global), stats[3] (g_loss, adv, rec), 3 pad.
    fake: G(noisy) from an earlier gen_fwd_f32 with the same gparams (a training iteration has it for its critic updates): the step
    then skips its first forward pass (ofdmgan_gen_step_fake)."""
    clean, noisy = frames(clean), frames(noisy)
    if out is None:
        out = torch.empty(GEN_OUT, dtype=torch.float32, device=clean.device)
    B = clean.shape[0]
    keepd, dp = _params(dparams, D_NPARAMS, "dparams")
    keepg, gp = _params(gparams, G_NPARAMS, "gparams")
    if fake is not None:
        fake = frames(fake)
        if fake.shape[0] != B or fake_out is not None:
            raise OfdmGanError("gen_step: fake must hold one frame per sample, and excludes fake_out")
        check(_lib.lib().ofdmgan_gen_step_fake(dptr(clean), dptr(noisy), dptr(fake), dp, gp, adv_weight, rec_weight, slope, B,
                                               B if b_global is None else b_global, dptr(out), stream_ptr(clean.device)))
        return out
    check(_lib.lib().ofdmgan_gen_step(dptr(clean), dptr(noisy), dp, gp, adv_weight, rec_weight, slope, B,
                                      B if b_global is None else b_global, dptr(out), dptr(fake_out), stream_ptr(clean.device)))
    return out


def gen_train(clean, noisy, fake, gparams, m, v, step_dev, lr, beta1, beta2, eps, dparams, adv_weight=1.0, rec_weight=100.0, slope=0.2,
              out=None, d_image_staged=False, g_image_staged=False, comm=None, b_global=None):
    """One whole generator update (loss, backward, gradient sum over the ranks of `comm` if given, Adam in place on gparams / m / v):
    two launches.  fake = gen_fwd_f32(noisy, gparams).  *_image_staged: see ofdmgan_gen_train_ctr - only right after critic_train on
    `dparams` / gen_fwd_f32 with this `gparams` tensor, with nothing else on the library in between."""
    clean, noisy, fake = frames(clean), frames(noisy), frames(fake)
    if out is None:
        out = torch.empty(GEN_OUT, dtype=torch.float32, device=clean.device)
    check(_lib.lib().ofdmgan_gen_train_ctr(dptr(clean), dptr(noisy), dptr(fake), dptr(step_dev), dptr(gparams), dptr(m), dptr(v), lr, beta1,
                                           beta2, eps, dptr(dparams), adv_weight, rec_weight, slope, clean.shape[0],
                                           clean.shape[0] if b_global is None else b_global, dptr(out), 1 if d_image_staged else 0,
                                           1 if g_image_staged else 0, comm._h if comm is not None else None, stream_ptr(clean.device)))
    return out


def adam(p, m, v, g, lr, beta1, beta2, eps, step, grad_scale=1.0, step_dev=None):
    """In-place fused Adam on flat fp32 CUDA vectors (torch.optim.Adam semantics, train.py:114-127).
    step_dev (int32 CUDA tensor, 1 element): the step count lives on the device - t = step_dev + 1 is used and stored back."""
    n = p.numel()
    if step_dev is not None:
        check(_lib.lib().ofdmgan_adam_ctr(dptr(p), dptr(m), dptr(v), dptr(g), n, lr, beta1, beta2, eps, dptr(step_dev), grad_scale,
                                          stream_ptr(p.device)))
        return
    check(_lib.lib().ofdmgan_adam(dptr(p), dptr(m), dptr(v), dptr(g), n, lr, beta1, beta2, eps, step, grad_scale,
                                  stream_ptr(p.device)))


class PeerComm:
    """Exchange blocks of the data-parallel ranks of ONE node, mapped into each other through CUDA IPC (include/ofdmgan.h,
    "data-parallel exchange fused with the optimiser").  Needs an initialised torch.distributed group with one process per GPU;
    the 64-byte handles travel through one all_gather on that group."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.device = _dev(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._h = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            check(_lib.lib().ofdmgan_comm_create(self.rank, self.world, ctypes.byref(self._h), ctypes.cast(handle, ctypes.c_void_p)))
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
            every = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine, group=group)
            blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in every)
            check(_lib.lib().ofdmgan_comm_connect(self._h, ctypes.c_char_p(blob)))
        dist.barrier(group)                                      # nobody sends before everybody has mapped everybody
        self.group = group

    def allreduce_adam(self, g, p=None, m=None, v=None, lr=0.0, beta1=0.0, beta2=0.0, eps=0.0, step=1, grad_scale=1.0, step_dev=None):
        """g <- sum over ranks of g (fixed rank order), then Adam on p / m / v with the first p.numel() entries of the sum.
        step_dev (int32 CUDA tensor, 1 element): Adam's step count lives on the device (graph-replayable)."""
        n_params = 0 if p is None else p.numel()
        if step_dev is not None:
            check(_lib.lib().ofdmgan_allreduce_adam_ctr(self._h, dptr(g), g.numel(), dptr(p), dptr(m), dptr(v), n_params, lr, beta1, beta2,
                                                        eps, dptr(step_dev), grad_scale, stream_ptr(g.device)))
            return
        check(_lib.lib().ofdmgan_allreduce_adam(self._h, dptr(g), g.numel(), dptr(p), dptr(m), dptr(v), n_params, lr, beta1, beta2, eps,
                                                step, grad_scale, stream_ptr(g.device)))

    def check(self):
        """Synchronise and raise if a peer ever failed to arrive."""
        check(_lib.lib().ofdmgan_comm_check(self._h, stream_ptr(self.device)))

    def close(self):
        if self._h:
            import torch.distributed as dist
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)                             # no peer may still be writing into a block that is about to go
            _lib.lib().ofdmgan_comm_destroy(self._h)
            self._h = ctypes.c_void_p()


def ffma_peak(iters=4096, device=None):
    device = _dev(device)
    out = ctypes.c_double(0)
    with torch.cuda.device(device):
        check(_lib.lib().ofdmgan_ffma_peak(iters, ctypes.byref(out), stream_ptr(device)))
    return out.value

"""ctypes binding of libofdmgan.so (include/ofdmgan.h).  There is no CPU fallback: if the CUDA library is missing or a
call fails, this raises."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFDMGAN_LIB") or os.path.join(_HERE, "lib", "libofdmgan.so")
_lib = None

G_NPARAMS, D_NPARAMS = 258, 521
CRITIC_OUT, GEN_OUT = 528, 264
N_METHODS, METRIC_COLS, MAX_SNR_BINS = 4, 8, 16
METHOD_GAN, METHOD_NOEQ, METHOD_ZF, METHOD_MMSE = 0, 1, 2, 3
GEN_F32, GEN_Q_SPEC, GEN_Q_RTL = 0, 1, 2

c_f = ctypes.c_float
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_p = ctypes.c_void_p


class OfdmGanError(RuntimeError):
    pass


class ChanCfg(ctypes.Structure):
    """`ofdmgan_chan_cfg` (include/ofdmgan.h)."""
    _fields_ = [
        ("symbol_source", ctypes.c_int32), ("n_fft", ctypes.c_int32), ("cp_len", ctypes.c_int32),
        ("pilot_spacing", ctypes.c_int32), ("pilot_re", c_f), ("pilot_im", c_f),
        ("ifft_scale", ctypes.c_int32), ("impair", ctypes.c_int32), ("pa_saturation", c_f),
        ("pa_smoothness", c_f), ("iq_gain", c_f), ("iq_cos", c_f), ("iq_sin", c_f), ("pn_sigma", c_f),
        ("snr_mode", ctypes.c_int32), ("snr_lo", c_f), ("snr_hi", c_f), ("snr_step", c_f),
        ("n_snr", ctypes.c_int32), ("frames_per_snr", c_i64), ("normalize", ctypes.c_int32),
        ("equalizers", ctypes.c_int32),
        ("channel_type", ctypes.c_int32), ("rician_k", ctypes.c_float), ("n_taps", ctypes.c_int32),
        ("tap_delay", ctypes.c_int32 * 4), ("tap_amp", ctypes.c_float * 4),
        ("saleh_alpha_a", ctypes.c_float), ("saleh_beta_a", ctypes.c_float), ("saleh_alpha_p", ctypes.c_float),
        ("saleh_beta_p", ctypes.c_float), ("dc_i", ctypes.c_float), ("dc_q", ctypes.c_float), ("cfo_step", ctypes.c_float),
        ("rng_rounds", ctypes.c_int32),
    ]


class ChanRand(ctypes.Structure):
    """`ofdmgan_chan_rand`: device pointers to host-generated draws (any may be NULL)."""
    _fields_ = [("sym", c_p), ("bits", c_p), ("pn", c_p), ("snr_db", c_p), ("noise", c_p), ("tx", c_p), ("fade", c_p), ("tx_gain", c_p)]


_SIGNATURES = {
    "ofdmgan_abi_version": (ctypes.c_int, []),
    "ofdmgan_device_sms": (ctypes.c_int, []),
    "ofdmgan_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "ofdmgan_gen_fwd_f32": (ctypes.c_int, [c_p, c_p, c_p, c_i64, c_f, c_p]),
    "ofdmgan_gen_bwd_f32": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_f, c_p]),
    "ofdmgan_gen_fwd_q": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_i64, ctypes.c_int, c_p, c_p]),
    "ofdmgan_disc_fwd_q": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, ctypes.c_int, c_p]),
    "ofdmgan_critic_step_ctr": (ctypes.c_int, [c_p, c_p, c_p, c_u64, c_u64, c_p, c_p, ctypes.c_float, ctypes.c_float, c_i64, c_i64, c_p, c_p]),
    "ofdmgan_critic_train_ctr": (ctypes.c_int, [c_p, c_p, c_p, c_u64, c_u64, c_p, c_p, c_p, c_p, ctypes.c_double, ctypes.c_double,
                                                ctypes.c_double, ctypes.c_double, ctypes.c_float, ctypes.c_float, c_i64, c_i64, c_p, ctypes.c_int, c_p, c_p]),
    "ofdmgan_adam_ctr": (ctypes.c_int, [c_p, c_p, c_p, c_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_double, c_p, ctypes.c_float, c_p]),
    "ofdmgan_comm_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), c_p]),
    "ofdmgan_comm_connect": (ctypes.c_int, [c_p, c_p]),
    "ofdmgan_comm_destroy": (ctypes.c_int, [c_p]),
    "ofdmgan_comm_check": (ctypes.c_int, [c_p, c_p]),
    "ofdmgan_allreduce_adam": (ctypes.c_int, [c_p, c_p, ctypes.c_int, c_p, c_p, c_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                              ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_float, c_p]),
    "ofdmgan_allreduce_adam_ctr": (ctypes.c_int, [c_p, c_p, ctypes.c_int, c_p, c_p, c_p, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                  ctypes.c_double, ctypes.c_double, c_p, ctypes.c_float, c_p]),
    "ofdmgan_compute_scale": (ctypes.c_int, [c_p, c_i64, c_i64, ctypes.c_int, c_p, c_p]),
    "ofdmgan_quantize_tensor": (ctypes.c_int, [c_p, c_i64, c_i64, c_p, ctypes.c_int, c_p, c_p]),
    "ofdmgan_dequantize_tensor": (ctypes.c_int, [c_p, c_i64, c_i64, c_p, c_p, c_p]),
    "ofdmgan_quantize_q88": (ctypes.c_int, [c_p, c_p, c_i64, c_p]),
    "ofdmgan_dequantize_q88": (ctypes.c_int, [c_p, c_p, c_i64, c_p]),
    "ofdmgan_chan_sim": (ctypes.c_int, [c_p, c_p, c_u64, c_u64, c_p, c_p, c_p, c_i64, c_p]),
    "ofdmgan_chan_draws": (ctypes.c_int, [c_p, c_u64, c_u64, c_p, c_p, c_p, c_p, c_p, c_i64, c_p]),
    "ofdmgan_philox_blocks": (ctypes.c_int, [c_u64, c_u64, ctypes.c_uint32, ctypes.c_uint32, c_p, c_i64, c_p]),
    "ofdmgan_chan_fade_draws": (ctypes.c_int, [c_p, c_u64, c_u64, c_p, c_i64, c_p]),
    "ofdmgan_unpackbits": (ctypes.c_int, [c_p, c_i64, c_p, c_p]),
    "ofdmgan_packbits": (ctypes.c_int, [c_p, c_i64, c_p, c_p]),
    "ofdmgan_qam_modulate": (ctypes.c_int, [c_p, c_p, c_i64, ctypes.c_int, c_p]),
    "ofdmgan_qam_demodulate": (ctypes.c_int, [c_p, c_p, c_i64, ctypes.c_int, c_p]),
    "ofdmgan_qpsk_modulate": (ctypes.c_int, [c_p, c_p, c_i64, c_p]),
    "ofdmgan_qpsk_demodulate": (ctypes.c_int, [c_p, c_p, c_i64, c_p]),
    "ofdmgan_ofdm_modulate": (ctypes.c_int, [c_p, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f, c_f, c_p, c_p]),
    "ofdmgan_ofdm_demodulate": (ctypes.c_int, [c_p, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f, c_f, c_p, c_p, c_p]),
    "ofdmgan_sim_gen_metrics": (ctypes.c_int, [c_p, ctypes.c_int, c_p, c_p, c_p, c_f, c_u64, c_u64, c_i64, c_p, c_p]),
    "ofdmgan_sim_gen_metrics_host": (ctypes.c_int, [c_p, ctypes.c_int, c_p, c_p, c_p, c_f, c_u64, c_u64, c_i64, c_p, c_p]),
    "ofdmgan_sim_impl_for": (ctypes.c_int, [c_p, ctypes.c_int, ctypes.c_int]),
    "ofdmgan_equalize": (ctypes.c_int, [c_p, c_p, c_p, ctypes.c_int, c_p, c_i64, c_p]),
    "ofdmgan_frame_metrics": (ctypes.c_int, [c_p, c_p, c_p, ctypes.c_int, ctypes.c_int, c_i64, c_p, c_p]),
    "ofdmgan_disc_fwd_f32": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_i64, c_f, c_p]),
    "ofdmgan_disc_bwd_f32": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_f, c_p]),
    "ofdmgan_gradient_penalty": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_u64, c_u64, ctypes.c_uint32, c_p, c_p, c_p, c_i64, c_f, c_p]),
    "ofdmgan_critic_step": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_u64, c_u64, ctypes.c_uint32, c_p, c_f, c_f, c_i64, c_i64, c_p, c_p]),
    "ofdmgan_gen_step": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_f, c_f, c_f, c_i64, c_i64, c_p, c_p, c_p]),
    "ofdmgan_gen_step_fake": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_f, c_i64, c_i64, c_p, c_p]),
    "ofdmgan_gen_train_ctr": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                             ctypes.c_double, c_p, c_f, c_f, c_f, c_i64, c_i64, c_p, ctypes.c_int, ctypes.c_int, c_p, c_p]),
    "ofdmgan_adam": (ctypes.c_int, [c_p, c_p, c_p, c_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_double, ctypes.c_int, c_f, c_p]),
    "ofdmgan_ffma_peak": (ctypes.c_int, [ctypes.c_int, c_p, c_p]),
}
EXPORTS = tuple(_SIGNATURES)


def lib():
    """Load libofdmgan.so.  Raises OfdmGanError if it has not been built (python ofdm-gan-sr_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OfdmGanError(f"{LIB_PATH} is missing: build it with `python ofdm-gan-sr_b200/build.py` "
                               "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        if L.ofdmgan_abi_version() != 16:
            raise OfdmGanError("libofdmgan ABI version mismatch")
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().ofdmgan_error_string(rc)
        raise OfdmGanError(f"libofdmgan error {rc}: {msg.decode() if msg else '?'}")


def stream_ptr(device=None):
    return c_p(torch.cuda.current_stream(device).cuda_stream)


def dptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise OfdmGanError("expected a CUDA tensor: libofdmgan has no CPU path")
    if not t.is_contiguous():
        raise OfdmGanError("expected a contiguous tensor")
    return c_p(t.data_ptr())


def frames(t, dtype=torch.float32):
    """[B,2,16] contiguous CUDA tensor of `dtype` (copying only if needed)."""
    if t.dim() != 3 or t.shape[1] != 2 or t.shape[2] != 16:
        raise OfdmGanError(f"expected frames of shape [B,2,16], got {tuple(t.shape)}")
    if not t.is_cuda:
        raise OfdmGanError("frames must live on a CUDA device: libofdmgan has no CPU path")
    return t.to(dtype).contiguous()

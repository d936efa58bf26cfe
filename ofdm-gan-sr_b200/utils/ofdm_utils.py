"""QAMModulator / OFDMModulator / NonLinearImpairments / ChannelModel with the reference's interface
(utils/ofdm_utils.py:90-832), computed by libofdmgan.

The reference works on one NumPy signal per call from the process-global np.random state.  Here every call is a batched
GPU launch:
  * a CUDA torch tensor in -> a CUDA torch tensor out (complex64; bits as uint8) - the fast path;
  * a NumPy array in -> a NumPy array out (complex128 / int), for callers that were written against the reference.
    The data makes a round trip through the device; there is still no CPU implementation behind it.
Signals handled by the channel stages are batches of 16-sample frames ([16] or [B,16]), the frame length of the
reference's training and benchmark paths; the memoryless stages (Rapp PA, IQ imbalance) accept any length.
Randomness comes from Philox4x32-10(seed, frame index): pass `seed=` / `frame0=` for reproducible draws; the class-level
counter otherwise advances so that successive calls see fresh noise, like successive np.random calls.
"""
import itertools
import math
from typing import Any, Dict, Tuple

import numpy as np
import torch

from .. import ops
from .._lib import OfdmGanError

_frame_counter = itertools.count()


def _to_dev(x, dtype):
    """-> (CUDA tensor, was_numpy)"""
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise OfdmGanError("expected a CUDA tensor (or a NumPy array): libofdmgan has no CPU path")
        return x.to(dtype), False
    return torch.as_tensor(np.ascontiguousarray(x)).cuda().to(dtype), True


def _back(t, was_numpy, np_dtype):
    return t.cpu().numpy().astype(np_dtype) if was_numpy else t


class QAMModulator:
    """utils/ofdm_utils.py:90-222: QPSK (the scheme of config/config.yaml:14), QAM16, QAM64."""

    CONSTELLATIONS = {"QPSK": {"bits_per_symbol": 2}, "QAM16": {"bits_per_symbol": 4}, "QAM64": {"bits_per_symbol": 6}}

    def __init__(self, modulation: str = "QAM16"):
        self.modulation = modulation.upper()
        if self.modulation not in self.CONSTELLATIONS:
            raise ValueError(f"Unsupported modulation: {modulation}")
        self.bits_per_symbol = self.CONSTELLATIONS[self.modulation]["bits_per_symbol"]
        if self.modulation == "QPSK":
            self.constellation = np.array([1 + 1j, 1 - 1j, -1 + 1j, -1 - 1j]) / np.sqrt(2)
        else:                                                    # utils/ofdm_utils.py:150-161
            sq = int(np.sqrt(2 ** self.bits_per_symbol))
            I, Q = np.meshgrid(np.arange(-sq + 1, sq, 2), np.arange(-sq + 1, sq, 2))
            self.constellation = (I + 1j * Q).flatten() / np.sqrt({4: 10, 6: 42}[self.bits_per_symbol])

    def modulate(self, bits):
        """bits (0/1, MSB first; trailing bits that do not fill a symbol are dropped) -> complex symbols."""
        b, was_np = _to_dev(bits, torch.uint8)
        b = b.reshape(-1)
        n = b.numel() // self.bits_per_symbol
        return _back(ops.qam_modulate(b[:n * self.bits_per_symbol], self.bits_per_symbol), was_np, np.complex128)

    def demodulate(self, symbols):
        """hard decisions: nearest constellation point, ties to the lowest index -> bits (flattened, MSB first)."""
        s, was_np = _to_dev(symbols, torch.complex64)
        return _back(ops.qam_demodulate(s, self.bits_per_symbol), was_np, np.int64)


class OFDMModulator:
    """utils/ofdm_utils.py:229-371.  n_subcarriers in {8, 16, 32, 64} are built (config/config.yaml:11 uses 8, the class default is 64)."""

    def __init__(self, n_subcarriers: int = 64, cp_length: int = 16, pilot_spacing: int = 8, pilot_value: complex = 1 + 0j):
        self.n_subcarriers, self.cp_length, self.pilot_spacing, self.pilot_value = n_subcarriers, cp_length, pilot_spacing, pilot_value
        self.pilot_indices = np.arange(0, n_subcarriers, pilot_spacing)
        self.data_indices = np.array([i for i in range(n_subcarriers) if i not in self.pilot_indices])
        self.n_data_subcarriers = len(self.data_indices)
        self.samples_per_symbol = n_subcarriers + cp_length

    def modulate(self, qam_symbols):
        s, was_np = _to_dev(qam_symbols, torch.complex64)
        return _back(ops.ofdm_modulate(s, self.n_subcarriers, self.cp_length, self.pilot_spacing, self.pilot_value), was_np, np.complex128)

    def demodulate(self, ofdm_signal):
        s, was_np = _to_dev(ofdm_signal, torch.complex64)
        data, chan = ops.ofdm_demodulate(s, self.n_subcarriers, self.cp_length, self.pilot_spacing, self.pilot_value)
        return _back(data, was_np, np.complex128), _back(chan, was_np, np.complex128)


def _frames_of(signal, allow_any_length):
    """complex signal -> (float32 [B,32] tx layout Re[16] | Im[16], restore(frames32 -> same kind/shape as the input))"""
    t, was_np = _to_dev(signal, torch.complex64)
    shape = t.shape
    flat = t.reshape(-1)
    L = shape[-1] if t.dim() else 1
    if L != 16 and not allow_any_length:
        raise OfdmGanError("this stage couples the samples of a frame: signals must be [16] or [B,16] (the reference's frame length)")
    n = flat.numel()
    pad = (-n) % 16
    if pad:
        flat = torch.cat([flat, torch.zeros(pad, dtype=flat.dtype, device=flat.device)])
    fr = flat.view(-1, 16)
    tx = torch.cat([fr.real, fr.imag], dim=1).contiguous()

    def restore(frames):                                         # frames: [B,2,16] float32
        c = torch.complex(frames[:, 0, :], frames[:, 1, :]).reshape(-1)[:n].reshape(shape)
        return _back(c, was_np, np.complex128)

    return tx, restore


def _run(tx, seed, frame0, **cfg_kw):
    if frame0 is None:
        frame0 = next(_frame_counter) * (1 << 32)
    cfg = ops.make_cfg(normalize=ops.NORM_NONE, **cfg_kw)
    clean, noisy, snr = ops.chan_sim(cfg, tx.shape[0], seed=seed, frame0=frame0, device=tx.device, tx=tx, want_clean=False, want_snr=False)
    return noisy


class NonLinearImpairments:
    """utils/ofdm_utils.py:378-605: Rapp and Saleh PA, IQ imbalance, Wiener phase noise, DC offset, CFO and the apply_all chain."""

    @staticmethod
    def apply_pa_rapp(signal, saturation_level: float = 1.0, smoothness: float = 3.0):
        tx, restore = _frames_of(signal, True)
        return restore(_run(tx, 0, 0, pa=True, iq=False, pn=False, pa_saturation=saturation_level, pa_smoothness=smoothness,
                            snr_mode=ops.SNR_NONE))

    @staticmethod
    def apply_pa_saleh(signal, alpha_a: float = 2.1587, beta_a: float = 1.1517, alpha_p: float = 4.0033, beta_p: float = 9.1040):
        """Saleh AM/AM + AM/PM model (utils/ofdm_utils.py:424-455)."""
        tx, restore = _frames_of(signal, True)
        return restore(_run(tx, 0, 0, pa=False, iq=False, pn=False, saleh=(alpha_a, beta_a, alpha_p, beta_p), snr_mode=ops.SNR_NONE))

    @staticmethod
    def apply_dc_offset(signal, dc_offset_i: float = 0.01, dc_offset_q: float = 0.01):
        """DC offset relative to the frame's RMS amplitude (utils/ofdm_utils.py:524-543)."""
        tx, restore = _frames_of(signal, False)
        return restore(_run(tx, 0, 0, pa=False, iq=False, pn=False, dc_offset=(dc_offset_i, dc_offset_q), snr_mode=ops.SNR_NONE))

    @staticmethod
    def apply_cfo(signal, cfo_hz: float = 100, sample_rate: float = 1e6):
        """Carrier frequency offset: sample n rotated by 2 pi cfo n / fs (utils/ofdm_utils.py:546-568)."""
        tx, restore = _frames_of(signal, False)
        return restore(_run(tx, 0, 0, pa=False, iq=False, pn=False, cfo_hz=cfo_hz, sample_rate=sample_rate, snr_mode=ops.SNR_NONE))

    @staticmethod
    def apply_iq_imbalance(signal, amplitude_imbalance_db: float = 1.0, phase_imbalance_deg: float = 5.0):
        tx, restore = _frames_of(signal, True)
        return restore(_run(tx, 0, 0, pa=False, iq=True, pn=False, iq_imbalance_db=amplitude_imbalance_db, iq_phase_deg=phase_imbalance_deg,
                            snr_mode=ops.SNR_NONE))

    @staticmethod
    def apply_phase_noise(signal, phase_noise_power_dbchz: float = -80, sample_rate: float = 1e6, seed: int = 0, frame0=None):
        tx, restore = _frames_of(signal, False)
        return restore(_run(tx, seed, frame0, pa=False, iq=False, pn=True, phase_noise_dbchz=phase_noise_power_dbchz, sample_rate=sample_rate,
                            snr_mode=ops.SNR_NONE))

    @staticmethod
    def apply_all(signal, pa_enabled: bool = True, pa_saturation: float = 1.0, iq_imbalance_enabled: bool = True, iq_amplitude_db: float = 1.0,
                  iq_phase_deg: float = 5.0, phase_noise_enabled: bool = True, phase_noise_dbchz: float = -80, dc_offset_enabled: bool = False,
                  cfo_enabled: bool = False, seed: int = 0, frame0=None):
        tx, restore = _frames_of(signal, not (phase_noise_enabled or dc_offset_enabled or cfo_enabled))
        return restore(_run(tx, seed, frame0, pa=pa_enabled, iq=iq_imbalance_enabled, pn=phase_noise_enabled, pa_saturation=pa_saturation,
                            iq_imbalance_db=iq_amplitude_db, iq_phase_deg=iq_phase_deg, phase_noise_dbchz=phase_noise_dbchz,
                            dc_offset=(0.01, 0.01) if dc_offset_enabled else None, cfo_hz=100 if cfo_enabled else None,
                            snr_mode=ops.SNR_NONE))


class ChannelModel:
    """utils/ofdm_utils.py:612-832: 'awgn' (per-frame measured signal power), 'rayleigh', 'rician' (k_factor=), 'multipath'
    (delays=, powers=; np.convolve(x, h, 'same') with independent Rayleigh taps).  The fading coefficient(s) are per frame."""

    def __init__(self, channel_type: str = "awgn"):
        self.channel_type = channel_type.lower()

    def apply(self, signal, snr_db: float, seed: int = 0, frame0=None, **kwargs) -> Tuple[Any, Dict[str, Any]]:
        if self.channel_type not in ops.CHANNEL_TYPES:
            raise ValueError(f"Unknown channel type: {self.channel_type}")
        tx, restore = _frames_of(signal, False)
        if frame0 is None:
            frame0 = next(_frame_counter) * (1 << 32)
        k = float(kwargs.get("k_factor", 3.0))
        delays, powers = list(kwargs.get("delays", [0, 1, 2])), list(kwargs.get("powers", [1.0, 0.5, 0.25]))
        cfg = ops.make_cfg(normalize=ops.NORM_NONE, pa=False, iq=False, pn=False, snr_mode=ops.SNR_GRID, snr_lo=float(snr_db), snr_step=0.0,
                           n_snr=1, channel_type=self.channel_type, rician_k=k, delays=delays, powers=powers)
        B = tx.shape[0]
        _, noisy, _ = ops.chan_sim(cfg, B, seed=seed, frame0=frame0, device=tx.device, tx=tx, want_clean=False, want_snr=False)
        info: Dict[str, Any] = {"type": self.channel_type, "snr_db": snr_db}
        x = torch.complex(tx[:, :16], tx[:, 16:])
        if self.channel_type == "awgn":
            faded, info["channel_response"] = x, np.array([1.0])
        else:
            f = ops.chan_fade_draws(cfg, B, seed=seed, frame0=frame0, device=tx.device)
            if self.channel_type == "rayleigh":
                h = torch.complex(f[:, 0], f[:, 1]) / math.sqrt(2.0)
            elif self.channel_type == "rician":
                h = math.sqrt(k / (k + 1)) * torch.polar(torch.ones_like(f[:, 0]), f[:, 0]) + \
                    math.sqrt(1 / (k + 1)) * torch.complex(f[:, 1], f[:, 2]) / math.sqrt(2.0)
                info["k_factor"] = k
            else:
                pw = np.asarray(powers, dtype=np.float64) / np.sum(powers)
                h = torch.zeros(B, max(delays) + 1, dtype=torch.complex64, device=tx.device)
                for t, (d, p) in enumerate(zip(delays, pw)):
                    h[:, d] = math.sqrt(p) * torch.complex(f[:, 2 * t], f[:, 2 * t + 1]) / math.sqrt(2.0)
                info["delays"], info["powers"] = delays, pw.tolist()
            if self.channel_type != "multipath":
                faded = h[:, None] * x
                info["channel_magnitude"] = h.abs() if B > 1 else float(h.abs()[0])
            else:
                faded = None
            info["channel_response"] = h if B > 1 else h[0]
        if faded is not None:
            power = (faded.real ** 2 + faded.imag ** 2).mean(dim=1)
            noise_power = power / (10.0 ** (float(snr_db) / 10.0))
            info["noise_power"] = noise_power if B > 1 else float(noise_power[0])
        return restore(noisy), info


class ImageOFDMConverter:
    """utils/ofdm_utils.py:839-1024: image -> bits -> QAM -> OFDM -> [2, L] float32 I/Q, and back.

    Bit (un)packing, the constellation map and the IFFT/FFT all run on the GPU (`images_to_ofdm` converts a whole list of
    images with one launch per stage); decoding image files stays with the caller (PIL), as in the reference."""

    def __init__(self, modulation: str = "QAM16", n_subcarriers: int = 64, cp_length: int = 16, frame_length: int = 1024):
        self.qam = QAMModulator(modulation)
        self.ofdm = OFDMModulator(n_subcarriers, cp_length)
        self.frame_length = frame_length
        self.modulation = modulation

    @staticmethod
    def _gray(image: np.ndarray) -> np.ndarray:
        image = np.asarray(image)
        if image.ndim == 3:                                      # 0.299 R + 0.587 G + 0.114 B, truncated (ofdm_utils.py:901-903)
            image = np.dot(image[..., :3], [0.299, 0.587, 0.114]).astype(np.uint8)
        return image

    def _signal(self, pixels_dev: torch.Tensor) -> torch.Tensor:
        """uint8 pixels (CUDA, flat) -> complex64 OFDM signal padded / truncated to frame_length"""
        bits = ops.unpackbits(pixels_dev)                                        # np.unpackbits: MSB first
        sig = self.ofdm.modulate(self.qam.modulate(bits))
        out = torch.zeros(self.frame_length, dtype=torch.complex64, device=sig.device)
        n = min(self.frame_length, sig.numel())
        out[:n] = sig[:n]
        return out

    def image_to_ofdm(self, image: np.ndarray, normalize: bool = True):
        image = self._gray(image)
        pixels = torch.as_tensor(np.ascontiguousarray(image.reshape(-1).astype(np.uint8))).cuda()
        sig = self._signal(pixels)
        iq = torch.stack([sig.real, sig.imag], dim=0)
        peak = iq.abs().max()
        max_val = float(peak) if normalize else 1.0
        if normalize and max_val > 0:
            iq = iq / peak                                       # tensor / tensor: a true division, the peak maps to exactly 1
        n_bits = pixels.numel() * 8
        metadata = {"original_shape": image.shape, "n_pixels": pixels.numel(), "n_bits": n_bits,
                    "n_qam_symbols": n_bits // self.qam.bits_per_symbol, "signal_length": self.frame_length,
                    "normalization_factor": max_val}
        return iq.cpu().numpy().astype(np.float32), metadata

    def images_to_ofdm(self, images, normalize: bool = True):
        """list of images -> ([n, 2, L] float32 CUDA tensor, [n] normalisation factors, list of metadata)"""
        sigs, metas = [], []
        for im in images:
            im = self._gray(im)
            sig = self._signal(torch.as_tensor(np.ascontiguousarray(im.reshape(-1).astype(np.uint8))).cuda())
            sigs.append(torch.stack([sig.real, sig.imag], dim=0))
            metas.append({"original_shape": im.shape, "n_pixels": im.size, "n_bits": im.size * 8,
                          "n_qam_symbols": im.size * 8 // self.qam.bits_per_symbol, "signal_length": self.frame_length})
        iq = torch.stack(sigs) if sigs else torch.zeros(0, 2, self.frame_length, device="cuda")
        factor = iq.abs().amax(dim=(1, 2)) if normalize else torch.ones(iq.shape[0], device=iq.device)
        if normalize:
            iq = iq / torch.where(factor > 0, factor, torch.ones_like(factor))[:, None, None]
        for m, f in zip(metas, factor.tolist()):
            m["normalization_factor"] = f
        return iq, factor, metas

    def ofdm_to_image(self, iq_signal, original_shape, denormalize_factor: float = 1.0) -> np.ndarray:
        iq, _ = _to_dev(iq_signal, torch.float32)
        iq = iq * float(denormalize_factor)
        sig = torch.complex(iq[0], iq[1])
        n_sym = sig.numel() // self.ofdm.samples_per_symbol
        data, _ = self.ofdm.demodulate(sig[:n_sym * self.ofdm.samples_per_symbol])
        bits = self.qam.demodulate(data).to(torch.uint8)
        n_pixels = int(np.prod(original_shape))
        need = n_pixels * 8
        if bits.numel() >= need:
            bits = bits[:need]
        else:
            bits = torch.cat([bits, torch.zeros(need - bits.numel(), dtype=torch.uint8, device=bits.device)])
        pixels = ops.packbits(bits)                                              # np.packbits
        return pixels.cpu().numpy().reshape(original_shape)

    def to_tensor(self, iq_signal: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.asarray(iq_signal)).unsqueeze(0).float()

    def from_tensor(self, tensor: torch.Tensor) -> np.ndarray:
        if tensor.dim() == 3:
            tensor = tensor[0]
        return tensor.detach().cpu().numpy()

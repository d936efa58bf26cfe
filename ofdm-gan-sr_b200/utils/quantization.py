"""Weight / activation quantisation with the reference's interface (utils/quantization.py:44-161) plus the Q1.7 / Q8.8
ROM export that feeds the integer generator kernel (rtl/ofdmGAN/weight_rom.v layout).

These functions touch a few hundred weights once per export: they are host-side tensor arithmetic, not part of the
per-frame hot path.  The per-frame integer arithmetic lives in libofdmgan (ofdmgan_gen_fwd_q, ofdmgan_quantize_q88).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops


class QuantizationConfig:
    """utils/quantization.py:44-70."""

    def __init__(self, weight_bits: int = 8, activation_bits: int = 16, accumulator_bits: int = 32, per_channel: bool = True):
        self.weight_bits, self.activation_bits, self.accumulator_bits = weight_bits, activation_bits, accumulator_bits
        self.per_channel = per_channel
        self.weight_max, self.weight_min = 2 ** (weight_bits - 1) - 1, -(2 ** (weight_bits - 1))
        self.activation_max, self.activation_min = 2 ** (activation_bits - 1) - 1, -(2 ** (activation_bits - 1))


def compute_scale(tensor: torch.Tensor, n_bits: int, per_channel: bool = False, channel_dim: int = 0) -> torch.Tensor:
    """scale = max(|x|, 1e-8) / (2^(n-1) - 1), per tensor or per channel (utils/quantization.py:73-112)."""
    qmax = 2 ** (n_bits - 1) - 1
    if per_channel:
        dims = [d for d in range(tensor.dim()) if d != channel_dim]
        amax = tensor.abs().amax(dim=dims, keepdim=True)
    else:
        amax = tensor.abs().max()
    return torch.clamp(amax, min=1e-8) / qmax


def quantize_tensor(tensor: torch.Tensor, scale: torch.Tensor, n_bits: int) -> torch.Tensor:
    """clamp(round_half_even(x / scale), -2^(n-1), 2^(n-1)-1), returned as float (utils/quantization.py:115-141)."""
    return torch.clamp(torch.round(tensor / scale), -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1)


def dequantize_tensor(quantized: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    return quantized * scale


class FakeQuantize(nn.Module):
    """Quantise-dequantise with a straight-through gradient (utils/quantization.py:164-205)."""

    def __init__(self, n_bits: int = 8, per_channel: bool = True, channel_dim: int = 0):
        super().__init__()
        self.n_bits, self.per_channel, self.channel_dim, self.momentum = n_bits, per_channel, channel_dim, 0.1
        self.register_buffer("scale", torch.tensor(1.0))
        self.register_buffer("running_max", torch.tensor(0.0))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            with torch.no_grad():
                self.running_max = (1 - self.momentum) * self.running_max + self.momentum * x.abs().max()
                self.scale = compute_scale(x, self.n_bits, self.per_channel, self.channel_dim)
        dq = dequantize_tensor(quantize_tensor(x, self.scale, self.n_bits), self.scale)
        return x + (dq - x).detach()


# ---- ROMs for the integer generator --------------------------------------------------------------------------------
# rtl/ofdmGAN/generator_mini.v:70-79: weight bases enc1 0, bneck 24, dec1 120, out 216 ([oc][ic][k]; the output conv is
# 1x1 = 8 weights); bias bases 0 / 4 / 12 / 16.
W_BASE = {"enc1": 0, "bottleneck": 24, "dec1": 120, "out_conv": 216}
B_BASE = {"enc1": 0, "bottleneck": 4, "dec1": 12, "out_conv": 16}


def export_q_roms(generator: nn.Module):
    """MiniGenerator -> (int8[2048] Q1.7 weight ROM, int16[64] Q8.8 bias ROM).

    Weights: quantize_tensor(w, scale=1/128, n_bits=8) (fixed Q1.7, README.md:241); biases: truncation toward zero
    to Q8.8 like every other float -> Q8.8 conversion the reference pins (proof/verification.py:297-298).  The RTL's
    output convolution is 1x1, so only the centre taps of out_conv are exported (utils/export_mini_weights.py:127-136)."""
    sd = {k: v.detach().cpu().float() for k, v in generator.state_dict().items()}
    wrom, brom = np.zeros(2048, np.int8), np.zeros(64, np.int16)
    q17 = lambda w: quantize_tensor(w, torch.tensor(1.0 / 128.0), 8).numpy().astype(np.int8).reshape(-1)
    q88 = lambda b: np.trunc(b.numpy().astype(np.float32) * np.float32(256.0)).astype(np.int16)
    for name, key in (("enc1", "enc1.conv"), ("bottleneck", "bottleneck.conv"), ("dec1", "dec1.conv")):
        w = q17(sd[key + ".weight"])
        wrom[W_BASE[name]:W_BASE[name] + w.size] = w
        b = q88(sd[key + ".bias"])
        brom[B_BASE[name]:B_BASE[name] + b.size] = b
    w = q17(sd["out_conv.weight"][:, :, 1])
    wrom[W_BASE["out_conv"]:W_BASE["out_conv"] + w.size] = w
    b = q88(sd["out_conv.bias"])
    brom[B_BASE["out_conv"]:B_BASE["out_conv"] + b.size] = b
    return wrom, brom


def float_to_q88(x: torch.Tensor) -> torch.Tensor:
    """(x*256).astype(int16), truncation toward zero, on the device (ofdmgan_quantize_q88)."""
    return ops.quantize_q88(x)


def q88_to_float(q: torch.Tensor) -> torch.Tensor:
    return ops.dequantize_q88(q)

"""Weight / activation quantisation with the reference's interface (utils/quantization.py:44-161) plus the Q1.7 / Q8.8
ROM export that feeds the integer generator kernel (rtl/ofdmGAN/weight_rom.v layout).

compute_scale / quantize_tensor / dequantize_tensor run in libofdmgan (ofdmgan_compute_scale, ofdmgan_quantize_tensor,
ofdmgan_dequantize_tensor) for CUDA tensors.  The file exporters at the bottom write a few hundred bytes from module state
that train.py keeps on the host at that point (train.py:530-531): host-side file output, the same tensor expressions as
upstream.  The per-frame integer arithmetic lives in ofdmgan_gen_fwd_q / ofdmgan_disc_fwd_q / ofdmgan_quantize_q88.
"""
import binascii
import json
from pathlib import Path
from typing import Any, Dict

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
    if tensor.is_cuda:
        if per_channel:
            t = tensor.movedim(channel_dim, 0)
            shape = [1] * tensor.dim()
            shape[channel_dim] = tensor.shape[channel_dim]
            return ops.compute_scale(t.reshape(t.shape[0], -1), n_bits).view(shape)
        return ops.compute_scale(tensor.reshape(1, -1), n_bits).view(())
    if per_channel:
        dims = [d for d in range(tensor.dim()) if d != channel_dim]
        amax = tensor.abs().amax(dim=dims, keepdim=True)
    else:
        amax = tensor.abs().max()
    return torch.clamp(amax, min=1e-8) / qmax


def quantize_tensor(tensor: torch.Tensor, scale: torch.Tensor, n_bits: int) -> torch.Tensor:
    """clamp(round_half_even(x / scale), -2^(n-1), 2^(n-1)-1), returned as float (utils/quantization.py:115-141)."""
    if tensor.is_cuda:
        return _per_channel_call(ops.quantize_tensor, tensor, scale, n_bits)
    return torch.clamp(torch.round(tensor / scale), -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1)


def dequantize_tensor(quantized: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    if quantized.is_cuda:
        return _per_channel_call(ops.dequantize_tensor, quantized, scale)
    return quantized * scale


def _per_channel_call(fn, tensor, scale, *extra):
    """Run a [C, inner] kernel for a scale that is one number or a keepdim per-channel tensor (what compute_scale returns)."""
    scale = torch.as_tensor(scale, dtype=torch.float32, device=tensor.device)
    if scale.numel() == 1:
        return fn(tensor.reshape(1, -1), scale.reshape(1), *extra).view(tensor.shape)
    dims = [d for d in range(scale.dim()) if scale.shape[d] != 1]
    if scale.dim() != tensor.dim() or len(dims) != 1 or scale.shape[dims[0]] != tensor.shape[dims[0]]:
        raise ops.OfdmGanError("scale must be a scalar or a keepdim per-channel tensor (the shapes compute_scale returns)")
    d = dims[0]
    t = tensor.movedim(d, 0)
    out = fn(t.reshape(t.shape[0], -1), scale.reshape(-1), *extra)
    return out.view(t.shape).movedim(0, d).contiguous()


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


# ---- weight export (utils/quantization.py:259-453): int8 weights + float32 scales / biases + CRC32 + metadata.json ------------
def compute_crc32(data: bytes) -> str:
    return f"{binascii.crc32(data) & 0xffffffff:08x}"


def compute_layer_crc(tensor: torch.Tensor) -> str:
    """CRC32 of a tensor's bytes (utils/quantization.py:443-453)."""
    return compute_crc32(tensor.detach().cpu().numpy().tobytes())


def _export_layer(name: str, layer: nn.Module, out: Path, config: QuantizationConfig) -> Dict[str, Any]:
    stem = name.replace(".", "_")
    weight = layer.weight.detach().cpu()
    scale = compute_scale(weight, config.weight_bits, config.per_channel, channel_dim=0)
    w_int8 = quantize_tensor(weight, scale, config.weight_bits).to(torch.int8).numpy().flatten()
    w_int8.tofile(out / f"{stem}_weights.bin")
    scale.squeeze().numpy().astype(np.float32).tofile(out / f"{stem}_scale.bin")
    bias_info = None
    if layer.bias is not None:
        layer.bias.detach().cpu().numpy().astype(np.float32).tofile(out / f"{stem}_bias.bin")
        bias_info = {"file": f"{stem}_bias.bin", "shape": list(layer.bias.shape)}
    info = {"type": "Conv1d" if isinstance(layer, nn.Conv1d) else "Linear", "weight_file": f"{stem}_weights.bin",
            "scale_file": f"{stem}_scale.bin", "bias": bias_info, "weight_shape": list(weight.shape)}
    if isinstance(layer, nn.Conv1d):
        info.update(kernel_size=layer.kernel_size[0], stride=layer.stride[0], padding=layer.padding[0], in_channels=layer.in_channels,
                    out_channels=layer.out_channels)
    else:
        info.update(in_features=layer.in_features, out_features=layer.out_features)
    info["crc32"] = compute_crc32(w_int8.tobytes())
    return info


def export_weights_fpga(model: nn.Module, output_dir: str, config: QuantizationConfig = None) -> Dict[str, Any]:
    """Same files, bytes and metadata.json as the reference's exporter (train.py:530-531 calls it after training): per layer
    `<name>_weights.bin` (int8, per-channel or per-tensor symmetric scale), `<name>_scale.bin`, `<name>_bias.bin` (float32), the
    CRC32 of the weight bytes.  Host-side file output of a few hundred bytes: not part of the device path."""
    config = config or QuantizationConfig()
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    metadata = {"config": {"weight_bits": config.weight_bits, "activation_bits": config.activation_bits, "per_channel": config.per_channel},
                "layers": {}}
    for name, module in model.named_modules():
        if isinstance(module, (nn.Conv1d, nn.Linear)):
            metadata["layers"][name] = _export_layer(name, module, out, config)
    with open(out / "metadata.json", "w") as f:
        json.dump(metadata, f, indent=2)
    return metadata


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

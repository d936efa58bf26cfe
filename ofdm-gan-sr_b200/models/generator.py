"""MiniGenerator with the reference's interface (models/generator.py:83-250), computed by libofdmgan kernel (2).

The module only owns parameters (same names and shapes as the reference: enc1.conv.weight [4,2,3] ... out_conv.bias [2],
258 in total, Xavier-uniform weights and zero biases); forward / backward are one fused launch each through a
torch.autograd.Function.  Integer inference (kernel (3)) is exposed as `forward_q88`.
"""
from typing import List

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from .. import ops
from .._lib import G_NPARAMS, OfdmGanError


class _GeneratorFunction(torch.autograd.Function):
    """y = G(x; theta): ofdmgan_gen_fwd_f32 / ofdmgan_gen_bwd_f32 (replaces the 11 ATen launches of
    models/generator.py:191-206 and their autograd graph)."""

    @staticmethod
    def forward(ctx, x, slope, *params):
        flat = ops.flat_cached(params)
        xc = ops.frames(x)
        y = ops.gen_fwd_f32(xc, flat, slope)
        ctx.save_for_backward(xc, flat)
        ctx.slope = slope
        ctx.shapes = [p.shape for p in params]
        ctx.need_dx = x.requires_grad
        return y

    @staticmethod
    @once_differentiable                     # first-order only: a double backward through this node must fail loudly, not drop terms
    def backward(ctx, dy):
        x, flat = ctx.saved_tensors
        dx, dflat = ops.gen_bwd_f32(x, flat, dy.contiguous(), ctx.slope, need_dx=ctx.need_dx)
        grads, off = [], 0
        for s in ctx.shapes:
            n = s.numel()
            grads.append(dflat[off:off + n].view(s))
            off += n
        return (dx, None, *grads)


class ConvBlock(nn.Module):
    """Parameter holder for Conv1d -> LeakyReLU (models/generator.py:37-80); keeps the `.conv.weight/.bias` names."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1, padding: int = 1,
                 leaky_slope: float = 0.2):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=True)
        self.activation = nn.LeakyReLU(negative_slope=leaky_slope)
        self.in_channels, self.out_channels, self.kernel_size, self.stride = in_channels, out_channels, kernel_size, stride

    def get_params_count(self) -> int:
        return self.kernel_size * self.in_channels * self.out_channels + self.out_channels

    def get_macs(self, output_length: int) -> int:
        return self.kernel_size * self.in_channels * self.out_channels * output_length


class MiniGenerator(nn.Module):
    """1-D U-Net 2 -> 4 -> 8 -> 4 -> 2 on 16-sample frames with an additive skip (models/generator.py:83-233)."""

    def __init__(self, input_channels: int = 2, output_channels: int = 2, frame_length: int = 16, leaky_slope: float = 0.2):
        super().__init__()
        self.input_channels, self.output_channels, self.frame_length = input_channels, output_channels, frame_length
        self.leaky_slope = leaky_slope
        self.enc1 = ConvBlock(input_channels, 4, 3, 2, 1, leaky_slope)
        self.bottleneck = ConvBlock(4, 8, 3, 2, 1, leaky_slope)
        self.dec1 = ConvBlock(8, 4, 3, 1, 1, leaky_slope)
        self.out_conv = nn.Conv1d(4, output_channels, 3, stride=1, padding=1, bias=True)
        for m in self.modules():                                   # models/generator.py:172-178
            if isinstance(m, nn.Conv1d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def _check(self):
        if (self.input_channels, self.output_channels, self.frame_length) != (2, 2, 16):
            raise OfdmGanError("libofdmgan builds the 2x16 MiniGenerator only (the configuration the reference trains and "
                               "puts in RTL); other channel counts / frame lengths are not supported")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._check()
        if not x.is_cuda:
            raise OfdmGanError("MiniGenerator.forward: input must be a CUDA tensor (libofdmgan has no CPU path)")
        return _GeneratorFunction.apply(x, self.leaky_slope, *self.parameters())

    @torch.no_grad()
    def forward_q88(self, x_q88: torch.Tensor, wrom, brom, mode: str = "spec") -> torch.Tensor:
        """Q1.7-weight / Q8.8-activation integer inference on int16 frames (rtl/ofdmGAN/generator_mini.v:326-649).
        mode 'spec' = RTL primitives on the textbook dataflow, 'rtl_literal' = what the committed RTL computes."""
        return ops.gen_fwd_q(x_q88, wrom, brom, ops.GEN_Q_SPEC if mode == "spec" else ops.GEN_Q_RTL)

    def get_layer_info(self) -> List[dict]:
        return [
            {"name": "enc1", "in_ch": 2, "out_ch": 4, "stride": 2, "length": 8},
            {"name": "bottleneck", "in_ch": 4, "out_ch": 8, "stride": 2, "length": 4},
            {"name": "upsample1", "scale": 2, "length": 8},
            {"name": "dec1", "in_ch": 8, "out_ch": 4, "stride": 1, "length": 8},
            {"name": "skip_add", "channels": 4, "length": 8},
            {"name": "upsample2", "scale": 2, "length": 16},
            {"name": "out_conv", "in_ch": 4, "out_ch": 2, "stride": 1, "length": 16},
            {"name": "tanh", "length": 16},
        ]

    def count_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def estimate_macs(self) -> int:
        return 192 + 384 + 768 + 384

    def flat_parameters(self) -> torch.Tensor:
        """258 floats in the packing of include/ofdmgan.h (= parameters() order)."""
        v = ops.flatten_params(self)
        assert v.numel() == G_NPARAMS
        return v


UNetGenerator = MiniGenerator


def create_generator(config: dict = None) -> MiniGenerator:
    config = config or {}
    return MiniGenerator(input_channels=config.get("input_channels", 2), output_channels=config.get("output_channels", 2),
                         frame_length=config.get("frame_length", 16), leaky_slope=config.get("leaky_slope", 0.2))

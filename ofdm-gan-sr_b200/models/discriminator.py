"""MiniDiscriminator and compute_gradient_penalty with the reference's interface (models/discriminator.py:42-250),
computed by libofdmgan kernel (4).

The critic is piecewise linear, so the gradient penalty's "double backward" has a closed form; it is
evaluated by one fused launch instead of building and differentiating an autograd graph of the backward pass.
"""
import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from .. import ops
from .._lib import D_NPARAMS, OfdmGanError


def _split(flat, shapes):
    out, off = [], 0
    for s in shapes:
        n = s.numel()
        out.append(flat[off:off + n].view(s))
        off += n
    return out


class _CriticFunction(torch.autograd.Function):
    """score = D(candidate, condition; theta): ofdmgan_disc_fwd_f32 / ofdmgan_disc_bwd_f32 (replaces the 7 ATen
    launches of models/discriminator.py:136-150 and their autograd graph)."""

    @staticmethod
    def forward(ctx, candidate, condition, slope, *params):
        flat = ops.flat_cached(params)
        cand, cond = ops.frames(candidate), ops.frames(condition)
        score = ops.disc_fwd_f32(cand, cond, flat, slope)
        ctx.save_for_backward(cand, cond, flat)
        ctx.slope = slope
        ctx.shapes = [p.shape for p in params]
        ctx.need = (candidate.requires_grad, condition.requires_grad, any(p.requires_grad for p in params))
        return score.view(-1, 1)

    @staticmethod
    @once_differentiable                     # first-order only: a double backward through this node must fail loudly, not drop terms
    def backward(ctx, g):
        cand, cond, flat = ctx.saved_tensors
        dcand, dcond, dflat = ops.disc_bwd_f32(cand, cond, flat, g.reshape(-1), ctx.slope, need_dcand=ctx.need[0],
                                               need_dcond=ctx.need[1], need_dparams=ctx.need[2])
        grads = _split(dflat, ctx.shapes) if dflat is not None else [None] * len(ctx.shapes)
        return (dcand, dcond, None, *grads)


class _GradientPenaltyFunction(torch.autograd.Function):
    """gp = mean_b (||d D(x_hat_b, c_b) / d x_hat_b||_2 - 1)^2 and d gp / d theta in closed form
    (ofdmgan_gradient_penalty).  Differentiable w.r.t. the critic's parameters, which is what train.py:237-253 needs;
    the samples themselves get no gradient (the reference detaches `fake` and never asks for d/d real)."""

    @staticmethod
    def forward(ctx, real, fake, cond, alpha, slope, *params):
        flat = ops.flat_cached(params)
        need = any(p.requires_grad for p in params)
        gp, grads = ops.gradient_penalty(ops.frames(real), ops.frames(fake), ops.frames(cond), flat, alpha=alpha, slope=slope,
                                         need_dparams=need)
        ctx.shapes = [p.shape for p in params]
        if need:
            ctx.save_for_backward(grads)
        ctx.has = need
        return gp.reshape(())

    @staticmethod
    @once_differentiable                     # first-order only: a double backward through this node must fail loudly, not drop terms
    def backward(ctx, g):
        if not ctx.has:
            return (None,) * (5 + len(ctx.shapes))
        (grads,) = ctx.saved_tensors
        return (None, None, None, None, None, *_split(grads * g, ctx.shapes))


class MiniDiscriminator(nn.Module):
    """Conditional critic cat(candidate, condition) -> Conv(4->8,s2) -> LReLU -> Conv(8->16,s2) -> LReLU -> sum over
    time -> Linear(16->1); 521 parameters (models/discriminator.py:42-164)."""

    def __init__(self, input_channels: int = 4, frame_length: int = 16, leaky_slope: float = 0.2):
        super().__init__()
        self.input_channels, self.frame_length, self.leaky_slope = input_channels, frame_length, leaky_slope
        self.conv1 = nn.Conv1d(input_channels, 8, 3, stride=2, padding=1, bias=True)
        self.conv2 = nn.Conv1d(8, 16, 3, stride=2, padding=1, bias=True)
        self.lrelu = nn.LeakyReLU(negative_slope=leaky_slope)
        self.dense = nn.Linear(16, 1)
        for m in self.modules():                                   # models/discriminator.py:104-110
            if isinstance(m, (nn.Conv1d, nn.Linear)):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def _check(self, *tensors):
        if (self.input_channels, self.frame_length) != (4, 16):
            raise OfdmGanError("libofdmgan builds the 4x16 MiniDiscriminator only; other shapes are not supported")
        for t in tensors:
            if not t.is_cuda:
                raise OfdmGanError("MiniDiscriminator: inputs must be CUDA tensors (libofdmgan has no CPU path)")

    def forward(self, candidate: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        self._check(candidate, condition)
        return _CriticFunction.apply(candidate, condition, self.leaky_slope, *self.parameters())

    def count_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def estimate_macs(self) -> int:
        return 768 + 1536 + 64 + 16

    def flat_parameters(self) -> torch.Tensor:
        v = ops.flatten_params(self)
        assert v.numel() == D_NPARAMS
        return v


Discriminator = MiniDiscriminator
ConditionalDiscriminator = MiniDiscriminator


def compute_gradient_penalty(discriminator: MiniDiscriminator, real_samples: torch.Tensor, fake_samples: torch.Tensor,
                             condition: torch.Tensor, device: torch.device = None) -> torch.Tensor:
    """WGAN-GP penalty, same signature and RNG use as models/discriminator.py:172-236: alpha ~ torch.rand(B,1,1) from
    torch's default generator on `device`, x_hat = alpha*real + (1-alpha)*fake, mean((||grad||-1)^2)."""
    if device is None:
        device = real_samples.device
    discriminator._check(real_samples, fake_samples, condition)
    alpha = torch.rand(real_samples.size(0), 1, 1, device=device)
    return _GradientPenaltyFunction.apply(real_samples.detach(), fake_samples.detach(), condition.detach(), alpha.view(-1),
                                          discriminator.leaky_slope, *discriminator.parameters())


def create_discriminator(config: dict = None) -> MiniDiscriminator:
    config = config or {}
    return MiniDiscriminator(input_channels=config.get("input_channels", 4), frame_length=config.get("frame_length", 16),
                             leaky_slope=config.get("leaky_slope", 0.2))

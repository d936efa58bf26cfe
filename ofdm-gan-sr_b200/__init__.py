"""ofdm_gan_sr_b200 - B200-native (sm_100a) hot path of orpheus016/ofdm-gan-sr.

Hand-written CUDA kernels behind the C ABI of include/ofdmgan.h (lib/libofdmgan.so), mirrored upward through the
reference's own Python call surface:

    models.MiniGenerator / MiniDiscriminator / compute_gradient_penalty      (models/generator.py, discriminator.py)
    utils.ofdm_utils  QAMModulator, OFDMModulator, NonLinearImpairments, ChannelModel, ImageOFDMConverter
    utils.dataset     SyntheticOFDMDataset, OFDMDataset (+ batched GPU frame source)
    utils.quantization  compute_scale / quantize_tensor / dequantize_tensor / FakeQuantize / Q-ROM export
    train_step.CWGANGPStep   the 5-critic + 1-generator step of train.py:327-344, data-parallel over NCCL
    sweep.run_benchmark      the SNR x trial loop of benchmark_comparison.py:179-250, sharded by frame index

There is no CPU fallback anywhere in this package: without the built library or without a CUDA device the compute
entry points raise `OfdmGanError`.
"""
from . import _lib, ops, sweep
from ._lib import OfdmGanError

__all__ = ["ops", "OfdmGanError", "_lib"]

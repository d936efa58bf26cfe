"""Mirrors of the reference's `utils` package for the hot path (utils/__init__.py:6-37)."""
from .ofdm_utils import ChannelModel, ImageOFDMConverter, NonLinearImpairments, OFDMModulator, QAMModulator
from .classical_equalizers import MMSEEqualizer, ZeroForcingEqualizer
from .dataset import GPUBatchLoader, OFDMDataset, SyntheticOFDMDataset, create_dataloader, generate_test_samples
from .quantization import (FakeQuantize, QuantizationConfig, compute_layer_crc, compute_scale, dequantize_tensor, export_q_roms,
                           export_weights_fpga, float_to_q88,
                           q88_to_float, quantize_tensor)

__all__ = ["ZeroForcingEqualizer", "MMSEEqualizer", "QAMModulator", "OFDMModulator", "ImageOFDMConverter", "OFDMDataset", "NonLinearImpairments", "ChannelModel", "SyntheticOFDMDataset", "GPUBatchLoader", "create_dataloader", "generate_test_samples", "QuantizationConfig", "compute_scale", "quantize_tensor",
           "dequantize_tensor", "FakeQuantize", "export_weights_fpga", "compute_layer_crc", "export_q_roms", "float_to_q88", "q88_to_float"]

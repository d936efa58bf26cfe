#!/usr/bin/env python
"""A few graph-replayed CWGAN-GP steps at 65,536 frames (the command profiled under ncu for the per-node launch list)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
import ofdm_gan_sr_b200 as pkg  # noqa: E402
from ofdm_gan_sr_b200.train_step import CWGANGPStep  # noqa: E402

gp, dp = bench.seed_params()
cfg = pkg.ops.make_cfg(normalize=1, snr_lo=0.0, snr_hi=30.0)
clean, noisy, _ = pkg.ops.chan_sim(cfg, 65536, seed=0)
tr = CWGANGPStep(gp, dp, graph=True)
for _ in range(4):
    tr.step(clean, noisy)
torch.cuda.synchronize()
print(tr.stats()["d_loss"])

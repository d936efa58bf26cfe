#!/usr/bin/env python
"""Run a few CWGAN-GP steps at BASELINE config 3 size (65,536 frames) - the command profiled under ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ofdm_gan_sr_b200 as pkg  # noqa: E402
from ofdm_gan_sr_b200.train_step import CWGANGPStep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
gp, dp = bench.seed_params()
cfg = pkg.ops.make_cfg(normalize=1, snr_lo=0.0, snr_hi=30.0)
clean, noisy, _ = pkg.ops.chan_sim(cfg, 65536, seed=0)
tr = CWGANGPStep(gp, dp)
for _ in range(steps):
    tr.step(clean, noisy)
torch.cuda.synchronize()
print(tr.stats()["d_loss"])

#!/usr/bin/env python
"""CWGAN-GP step at 65,536 frames as bench.py times it (one CUDA graph per iteration), 5 x 200 steps: min / median ms per step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
import ofdm_gan_sr_b200 as pkg  # noqa: E402
from ofdm_gan_sr_b200.train_step import CWGANGPStep  # noqa: E402

if os.environ.get('NO_GEN_TRAIN'):
    del pkg.ops.gen_train                  # A/B: generator update as gen_step + separate Adam launch
gp, dp = bench.seed_params()
cfg = pkg.ops.make_cfg(normalize=1, snr_lo=0.0, snr_hi=30.0)
clean, noisy, _ = pkg.ops.chan_sim(cfg, 65536, seed=0)
tr = CWGANGPStep(gp, dp, graph=True, reuse_fake=os.environ.get('OWN_FWD') is None)
for _ in range(20):
    tr.step(clean, noisy)
ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        tr.step(clean, noisy)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1) / 200)
print("train ms/step min %.4f median %.4f  d_loss %.6f" % (min(ms), float(np.median(ms)), tr.stats()["d_loss"]))

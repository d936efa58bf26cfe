import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import oracle
import ofdm_gan_sr_b200 as pkg
ops = pkg.ops
rng = np.random.default_rng(0)
gp = (rng.standard_normal(258) * 0.3).astype(np.float32)
for eq in (False, True):
    kw = dict(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=2000)
    cfg = ops.make_cfg(equalizers=eq, **kw)
    B = 14000
    m = ops.sim_gen_metrics(cfg, B, gparams=gp, seed=1).cpu().numpy()
    o = oracle.sim_gen_metrics(oracle.make_cfg(equalizers=eq, **kw), 0, B, gparams=gp, seed=1)
    print("eq", eq)
    for meth in range(4 if eq else 2):
        print(" method", meth, "n", m[:, meth, 0], "evm mean gpu", np.round(m[:, meth, 3] / np.maximum(m[:, meth, 0], 1), 3))
        print("                      evm mean orc", np.round(o[:, meth, 3] / np.maximum(o[:, meth, 0], 1), 3))

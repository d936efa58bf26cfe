import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
import ofdm_gan_sr_b200 as pkg
ops = pkg.ops
cfg = ops.make_cfg(nonlinear=True, snr_lo=0.0, snr_hi=30.0)
ocfg = oracle.make_cfg(nonlinear=True, snr_lo=0.0, snr_hi=30.0)
for seed, f0 in ((7, (1 << 33) + 5), (1, 0), (3, 12345)):
    d = ops.chan_draws(cfg, 65536, seed=seed, frame0=f0)
    o = oracle.frame_draws(ocfg, seed, f0, 65536)
    for k in ("sym", "pn", "noise"):
        g = d[k].cpu().numpy()
        e = np.abs(g - o[k])
        i = np.unravel_index(np.argmax(e), e.shape)
        print(seed, k, "max abs err %.3e at value %.5f (oracle %.5f), mean abs err %.2e, 99.99pct %.2e" % (e.max(), g[i], o[k][i], e.mean(), np.quantile(e, 0.9999)))

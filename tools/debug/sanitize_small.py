"""A small pass over the entry points touched this round, for compute-sanitizer (memcheck / racecheck): ragged sizes, both simulator
kernels, by-value and constant-image weights, the dataset path, the metrics path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ofdm_gan_sr_b200 as pkg
ops = pkg.ops
rng = np.random.default_rng(0)
gp = (rng.standard_normal(258) * 0.3).astype(np.float32)
gp_d = torch.as_tensor(gp).cuda()
for kw in (dict(), dict(nonlinear=True, pa_saturation=0.8), dict(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, n_snr=7, frames_per_snr=5),
           dict(nonlinear=True, pa_saturation=0.8, rng_rounds=7, snr_mode=1, n_snr=16, snr_step=2.0, frames_per_snr=3)):
    cfg = ops.make_cfg(**kw)
    for B in (1, 31, 33, 1000, 4099):
        c, n, s = ops.chan_sim(cfg, B, seed=1, frame0=12345)
        m1 = ops.sim_gen_metrics(cfg, B, gparams=gp, seed=1)          # by-value image
        m2 = ops.sim_gen_metrics(cfg, B, gparams=gp_d, seed=1)        # constant image
        assert torch.equal(m1, m2)
cfg = ops.make_cfg(nonlinear=True, pa_saturation=0.8, equalizers=True, channel_type="multipath", normalize=2, snr_mode=1, n_snr=7, frames_per_snr=9)
ops.sim_gen_metrics(cfg, 777, gparams=gp, seed=2)
sym = rng.standard_normal((100, 32)); noise = rng.standard_normal((100, 32)); pn = rng.standard_normal((100, 16))
ops.chan_sim(ops.make_cfg(nonlinear=True), 100, sym=sym, noise=noise, pn=pn, snr_db=np.full(100, 10.0))
x = torch.randn(1000, 2, 16, device="cuda")
ops.gen_fwd_f32(x, gp); ops.gen_fwd_f32(x, gp_d)
W = rng.integers(-128, 128, 2048).astype(np.int8); Bq = np.zeros(64, np.int16)
ops.gen_fwd_q(ops.quantize_q88(x), W, Bq, mode=ops.GEN_Q_SPEC)
torch.cuda.synchronize()
print("sanitize pass ok")

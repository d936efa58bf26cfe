"""Is the critic kernel losing time to its uneven last wave?  Per-sample time at 65,536 samples (1,536 work items on 592 CTA
slots: loads 4 / 3 / 2 units) against 75,776 = 592 x 128 samples (every CTA gets exactly one penalty and two score items)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ofdm_gan_sr_b200 as pkg
ops = pkg.ops
rng = np.random.default_rng(0)
gp = (rng.standard_normal(258) * 0.3).astype(np.float32)
dp = torch.as_tensor((rng.standard_normal(521) * 0.2).astype(np.float32)).cuda()
for B in (65536, 75776, 592 * 128 * 2, 60000, 592 * 96):
    clean, noisy, _ = ops.chan_sim(ops.make_cfg(normalize=1), B, seed=1)
    fake = ops.gen_fwd_f32(noisy, gp)
    out = torch.zeros(528, device="cuda")
    for _ in range(5):
        ops.critic_step(clean, noisy, fake, dp, seed=1, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        ops.critic_step(clean, noisy, fake, dp, seed=1, alpha_iter=i, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"B={B:7d}  {us:8.2f} us per critic_step (kernel + finalize)  {us / B * 1e3:7.3f} ns per sample")

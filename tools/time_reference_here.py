#!/usr/bin/env python3
"""Time the UNMODIFIED reference's own CPU path for the hot path.  Build container only: imports
/root/reference, which does not exist on the GPU box, so nothing under tests/ or bench.py uses this.  Writes
profiles/r1_reference_cpu_container.json (the host is this container's CPU share, not the GPU box's)."""
import json
import os
import statistics
import sys
import time
import types

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
import benchmark_comparison as bc  # noqa: E402
from models import MiniDiscriminator, MiniGenerator, compute_gradient_penalty  # noqa: E402
from utils.dataset import SyntheticOFDMDataset  # noqa: E402


def med(fn, reps=5):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return statistics.median(t)


out = {"host": {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "torch_threads": torch.get_num_threads(),
                "torch": torch.__version__, "numpy": np.__version__}}
np.random.seed(0)
torch.manual_seed(0)
for tag, kw in (("awgn_10dB", dict(snr_range=(10, 10))), ("nonlinear", dict(nonlinear=True, pa_saturation=0.8))):
    ds = SyntheticOFDMDataset(n_samples=2000, **kw)
    out[f"dataset_getitem_frames_per_s_{tag}"] = 2000 / med(lambda: [ds[i] for i in range(2000)], 3)
G, D = MiniGenerator(), MiniDiscriminator()
ds = SyntheticOFDMDataset(n_samples=6400, snr_range=(10, 10))
loader = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False, num_workers=0)


def c1():
    with torch.no_grad():
        for b in loader:
            G(b["noisy"])


out["c1_dataloader_b64_plus_generator_frames_per_s"] = 6400 / med(c1, 3)
for B in (64, 65536):
    x = torch.randn(B, 2, 16)
    with torch.no_grad():
        out[f"generator_forward_frames_per_s_B{B}"] = B / med(lambda: G(x), 5)
opt_g = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.0, 0.9))
opt_d = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.0, 0.9))


def train_step(clean, noisy):
    for _ in range(5):                                         # train.py:201-261
        with torch.no_grad():
            fake = G(noisy)
        opt_d.zero_grad()
        gp = compute_gradient_penalty(D, clean, fake, noisy, device="cpu")
        d_loss = D(fake, noisy).mean() - D(clean, noisy).mean() + 10.0 * gp
        d_loss.backward()
        opt_d.step()
    opt_g.zero_grad()                                          # train.py:263-305
    fake = G(noisy)
    g_loss = -D(fake, noisy).mean() + 100.0 * torch.nn.functional.l1_loss(fake, clean)
    g_loss.backward()
    opt_g.step()


for B in (64, 65536):
    clean, noisy = torch.randn(B, 2, 16).clamp(-1, 1), torch.randn(B, 2, 16).clamp(-1, 1)
    out[f"train_step_samples_per_s_B{B}"] = B / med(lambda: train_step(clean, noisy), 3)
t0 = time.perf_counter()
bc.run_benchmark(G, n_trials=100, nonlinear=True, pa_saturation=0.8) if "pa_saturation" in bc.run_benchmark.__code__.co_varnames else \
    bc.run_benchmark(G, n_trials=100, nonlinear=True)
out["run_benchmark_trials_per_s_n100_x7snr"] = 700 / (time.perf_counter() - t0)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r1_reference_cpu_container.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))

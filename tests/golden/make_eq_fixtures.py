#!/usr/bin/env python3
"""Record ZF / MMSE equaliser fixtures from the UNMODIFIED reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_eq_fixtures.py   ->  tests/golden/ref_eq.npz

Frames are produced exactly as benchmark_comparison.run_benchmark does (:184-197, non-linear scenario), then passed through
ZeroForcingEqualizer.equalize_iq / MMSEEqualizer.equalize_iq (:218-226) and compute_mse / compute_evm (:137-146)."""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
import benchmark_comparison as bc  # noqa: E402
from utils.classical_equalizers import MMSEEqualizer, ZeroForcingEqualizer  # noqa: E402

np.random.seed(11)
noisy, clean, snrs, zf, mm, met = [], [], [], [], [], []
for snr in (0, 5, 10, 15, 20, 25, 30):
    for trial in range(100):
        c = bc.generate_test_signal(16, "ofdm")
        n, _ = bc.apply_channel_and_impairments(c, snr, "awgn", trial % 2 == 1, 0.8)
        ci, ni = bc.complex_to_iq(c), bc.complex_to_iq(n)
        nn, _ = bc.normalize_iq(ni)
        cn, _ = bc.normalize_iq(ci)
        z, _ = ZeroForcingEqualizer().equalize_iq(nn, cn)
        m, _ = MMSEEqualizer().equalize_iq(nn, cn, snr_db=snr)
        noisy.append(nn); clean.append(cn); snrs.append(snr); zf.append(z); mm.append(m)
        met.append([bc.compute_mse(z, cn), bc.compute_evm(z, cn), bc.compute_mse(m, cn), bc.compute_evm(m, cn)])
out = dict(noisy=np.array(noisy, np.float32), clean=np.array(clean, np.float32), snr=np.array(snrs, np.float64),
           zf=np.array(zf, np.float32), mmse=np.array(mm, np.float32), metrics=np.array(met, np.float64))
np.savez_compressed(os.path.join(HERE, "ref_eq.npz"), **out)
m = out["metrics"]
print({k: v.shape for k, v in out.items()}, "zf exact-zero trials:", int((m[:, 0] == 0).sum()), "of", len(m), "zf evm range", m[:, 1].min(), m[:, 1].max())
print("numpy", np.__version__, "mmse evm mean", m[:, 3].mean())

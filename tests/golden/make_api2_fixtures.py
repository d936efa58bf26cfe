#!/usr/bin/env python3
"""Record fixtures for the remaining utils/ofdm_utils.py models from the UNMODIFIED reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_api2_fixtures.py   ->  tests/golden/ref_api2.npz

QAMModulator QAM16 / QAM64 (modulate, demodulate incl. exact ties), NonLinearImpairments.apply_pa_saleh / apply_dc_offset /
apply_cfo / apply_all with DC + CFO enabled, ChannelModel rayleigh / rician / multipath, each with its recorded np.random draws."""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from utils.ofdm_utils import ChannelModel, NonLinearImpairments, QAMModulator  # noqa: E402

rng = np.random.RandomState(77)
out = {}
B = 24
x = (rng.randn(B, 16) + 1j * rng.randn(B, 16)) * (0.3 + 0.9 * rng.rand(B, 1))
out["x"] = x
NL = NonLinearImpairments
out["saleh"] = np.stack([NL.apply_pa_saleh(r) for r in x])
out["saleh2"] = np.stack([NL.apply_pa_saleh(r, 1.5, 0.8, 2.0, 5.0) for r in x])
out["dc"] = np.stack([NL.apply_dc_offset(r, 0.01, 0.01) for r in x])
out["dc2"] = np.stack([NL.apply_dc_offset(r, -0.05, 0.2) for r in x])
out["cfo"] = np.stack([NL.apply_cfo(r, 100, 1e6) for r in x])
out["cfo2"] = np.stack([NL.apply_cfo(r, 25000, 1e6) for r in x])

real = {"randn": np.random.randn, "uniform": np.random.uniform}


def recording(fn):
    log = []

    def randn(*shape):
        v = real["randn"](*shape)
        log.append(np.atleast_1d(np.asarray(v, dtype=np.float64)).reshape(-1))
        return v

    def uniform(lo=0.0, hi=1.0, size=None):
        v = real["uniform"](lo, hi, size)
        log.append(np.atleast_1d(np.asarray(v, dtype=np.float64)).reshape(-1))
        return v
    np.random.randn, np.random.uniform = randn, uniform
    try:
        r = fn()
    finally:
        np.random.randn, np.random.uniform = real["randn"], real["uniform"]
    return r, np.concatenate(log)


np.random.seed(5)
acc = {k: [] for k in ("all_dc", "all_dc_d", "ray", "ray_d", "ray_h", "ric", "ric_d", "ric_h", "mp", "mp_d", "mp_h", "mp2", "mp2_d", "mp2_h")}
for r in x:
    y, d = recording(lambda: NL.apply_all(r, pa_saturation=0.8, dc_offset_enabled=True, cfo_enabled=True))
    acc["all_dc"].append(y); acc["all_dc_d"].append(d)                                   # 16 phase-noise normals
    (y, info), d = recording(lambda: ChannelModel("rayleigh").apply(r, 15.0))
    acc["ray"].append(y); acc["ray_d"].append(d); acc["ray_h"].append(info["channel_response"])   # 2 + 32 draws
    (y, info), d = recording(lambda: ChannelModel("rician").apply(r, 15.0, k_factor=4.0))
    acc["ric"].append(y); acc["ric_d"].append(d); acc["ric_h"].append(info["channel_response"])   # 3 + 32
    (y, info), d = recording(lambda: ChannelModel("multipath").apply(r, 15.0))
    acc["mp"].append(y); acc["mp_d"].append(d); acc["mp_h"].append(info["channel_response"])      # 6 + 32
    (y, info), d = recording(lambda: ChannelModel("multipath").apply(r, 15.0, delays=[0, 3], powers=[2.0, 1.0]))
    acc["mp2"].append(y); acc["mp2_d"].append(d); acc["mp2_h"].append(info["channel_response"])   # 4 + 32
out.update({k: np.stack(v) for k, v in acc.items()})

for tag, name in (("q16", "QAM16"), ("q64", "QAM64")):
    q = QAMModulator(name)
    bits = rng.randint(0, 2, q.bits_per_symbol * 300 + 3)
    syms = q.modulate(bits)
    noisy = syms + 0.25 * (rng.randn(300) + 1j * rng.randn(300))
    lv = q.constellation.real.max()
    noisy[:10] = [0, 0.5 * lv + 0j, -1e-12 + 0j, 5 + 5j, -5 - 5j, 2 * lv / 3 + 0j, 0 + 2j * lv / 3, q.constellation[3] * 1.0,
                  (q.constellation[0] + q.constellation[1]) / 2, (q.constellation[0] + q.constellation[int(np.sqrt(len(q.constellation)))]) / 2]
    out.update({tag + "_bits": bits, tag + "_syms": syms, tag + "_noisy": noisy, tag + "_demod": q.demodulate(noisy), tag + "_const": q.constellation})
np.savez_compressed(os.path.join(HERE, "ref_api2.npz"), **out)
print({k: v.shape for k, v in out.items()})

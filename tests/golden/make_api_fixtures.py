#!/usr/bin/env python3
"""Record fixtures for the utils/ofdm_utils.py call surface from the UNMODIFIED reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_api_fixtures.py   ->  tests/golden/ref_api.npz

QAMModulator('QPSK') modulate / demodulate, OFDMModulator modulate / demodulate, NonLinearImpairments.apply_pa_rapp /
apply_iq_imbalance / apply_phase_noise / apply_all and ChannelModel('awgn').apply on 16-sample frames, with the np.random
draws each call consumed recorded beside its output."""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from utils.ofdm_utils import ChannelModel, NonLinearImpairments, OFDMModulator, QAMModulator  # noqa: E402

rng = np.random.RandomState(2024)
out = {}
B = 24
x = (rng.randn(B, 16) + 1j * rng.randn(B, 16)) * (0.3 + 0.9 * rng.rand(B, 1))
out["x"] = x
out["pa_08_3"] = np.stack([NonLinearImpairments.apply_pa_rapp(r, 0.8, 3.0) for r in x])
out["pa_10_2"] = np.stack([NonLinearImpairments.apply_pa_rapp(r, 1.0, 2.0) for r in x])
out["iq_1_5"] = np.stack([NonLinearImpairments.apply_iq_imbalance(r, 1.0, 5.0) for r in x])
out["iq_m05_m3"] = np.stack([NonLinearImpairments.apply_iq_imbalance(r, -0.5, -3.0) for r in x])
long = (rng.randn(37) + 1j * rng.randn(37))
out["x_long"], out["pa_long"] = long, NonLinearImpairments.apply_pa_rapp(long, 0.7, 3.0)

real_randn = np.random.randn


def recording(fn, n_draws):
    """run fn with np.random.randn recorded; returns (result, draws[n_draws])"""
    log = []

    def randn(*shape):
        v = real_randn(*shape)
        log.append(np.asarray(v, dtype=np.float64).reshape(-1))
        return v
    np.random.randn = randn
    try:
        r = fn()
    finally:
        np.random.randn = real_randn
    d = np.concatenate(log) if log else np.zeros(0)
    assert d.size == n_draws, (d.size, n_draws)
    return r, d


np.random.seed(7)
pn, pn_d, al, al_d, aw, aw_d, aw_np = [], [], [], [], [], [], []
for r in x:
    y, d = recording(lambda: NonLinearImpairments.apply_phase_noise(r, -80, 1e6), 16)
    pn.append(y); pn_d.append(d)
    y, d = recording(lambda: NonLinearImpairments.apply_all(r, pa_saturation=0.8, dc_offset_enabled=False, cfo_enabled=False), 16)
    al.append(y); al_d.append(d)
    (y, info), d = recording(lambda: ChannelModel("awgn").apply(r, 12.5), 32)
    aw.append(y); aw_d.append(d); aw_np.append(info["noise_power"])
out.update(pn=np.stack(pn), pn_draws=np.stack(pn_d), all=np.stack(al), all_draws=np.stack(al_d), awgn=np.stack(aw),
           awgn_draws=np.stack(aw_d), awgn_noise_power=np.array(aw_np))

q = QAMModulator("QPSK")
bits = rng.randint(0, 2, 2 * 203 + 1)
out["bits"], out["syms"] = bits, q.modulate(bits)
noisy = out["syms"] + 0.5 * (rng.randn(203) + 1j * rng.randn(203))
noisy[:8] = [0, 1e-9, -1e-9j, 0.3, -0.3j, 0.7 + 0j, 0 - 0.7j, -0.0 + 0.0j]      # exact ties
out["noisy_syms"], out["demod_bits"] = noisy, q.demodulate(noisy)
for tag, (N, cp, sp, pv) in {"o8": (8, 2, 4, 1 + 0j), "o16": (16, 4, 8, 0.5 - 0.5j), "o16np": (16, 0, 32, 1 + 0j)}.items():
    m = OFDMModulator(N, cp, sp, pv)
    s = out["syms"][:5 * m.n_data_subcarriers - 3]                 # not a multiple: exercises the zero padding
    sig = m.modulate(s)
    rx = sig * (0.8 + 0.3j) + 0.01 * (rng.randn(sig.size) + 1j * rng.randn(sig.size))
    data, chan = m.demodulate(rx)
    out.update({tag + "_in": s, tag + "_sig": sig, tag + "_rx": rx, tag + "_data": data, tag + "_chan": chan})
np.savez_compressed(os.path.join(HERE, "ref_api.npz"), **out)
print({k: v.shape for k, v in out.items()})

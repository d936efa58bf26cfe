"""ZeroForcingEqualizer / MMSEEqualizer with the reference's interface (utils/classical_equalizers.py:33-230) for the way the
benchmark uses them: `equalize_iq(noisy_iq, clean_iq[, snr_db])` with the genie channel estimate H = Y / (X + eps).

Batched: [B,2,16] (or one [2,16] frame); CUDA tensors in -> CUDA tensors out, NumPy in -> NumPy out.  The arithmetic is the
complex64 arithmetic NumPy 2 gives the reference (ZF frames are bit-identical to it).  The decision-feedback, LMS and RLS
equalisers are serial per-sample recursions - not data-parallel, out of scope."""
from typing import Dict, Tuple

import numpy as np
import torch

from .. import ops
from .._lib import METHOD_MMSE, METHOD_ZF, OfdmGanError


def _prep(iq):
    was_np = not isinstance(iq, torch.Tensor)
    t = torch.as_tensor(np.ascontiguousarray(iq)).cuda() if was_np else iq
    if not t.is_cuda:
        raise OfdmGanError("expected a CUDA tensor (or a NumPy array): libofdmgan has no CPU path")
    single = t.dim() == 2
    return (t.unsqueeze(0) if single else t).float(), was_np, single


def _finish(est, noisy, clean, was_np, single) -> Tuple[object, Dict[str, object]]:
    mse = ((est - clean) ** 2).mean(dim=(1, 2))
    gain = 10 * torch.log10((noisy ** 2).mean(dim=(1, 2)) / (mse + 1e-10))
    if single:
        est, metrics = est[0], {"mse": float(mse[0]), "snr_improvement_db": float(gain[0])}
    else:
        metrics = {"mse": mse, "snr_improvement_db": gain}
    return (est.cpu().numpy() if was_np else est), metrics


class ZeroForcingEqualizer:
    """X_hat = Y / (H + eps), H = Y / (X + eps)   (utils/classical_equalizers.py:33-126)."""

    def __init__(self, n_subcarriers: int = 64):
        self.n_subcarriers, self.channel_estimate = n_subcarriers, None

    def equalize_iq(self, noisy_iq, clean_iq=None):
        if clean_iq is None:
            raise OfdmGanError("the genie-aided form (clean_iq given) is the one the benchmark uses and the one built here")
        noisy, was_np, single = _prep(noisy_iq)
        clean, _, _ = _prep(clean_iq)
        return _finish(ops.equalize(noisy, clean, METHOD_ZF), noisy, clean, was_np, single)


class MMSEEqualizer:
    """X_hat = conj(H) / (|H|^2 + 1/SNR) * Y   (utils/classical_equalizers.py:129-230)."""

    def __init__(self, n_subcarriers: int = 64):
        self.n_subcarriers, self.channel_estimate, self.noise_variance = n_subcarriers, None, None

    def equalize_iq(self, noisy_iq, clean_iq=None, snr_db=20.0):
        if clean_iq is None:
            raise OfdmGanError("the genie-aided form (clean_iq given) is the one the benchmark uses and the one built here")
        noisy, was_np, single = _prep(noisy_iq)
        clean, _, _ = _prep(clean_iq)
        return _finish(ops.equalize(noisy, clean, METHOD_MMSE, snr_db=snr_db), noisy, clean, was_np, single)

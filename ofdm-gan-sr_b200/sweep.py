"""The SNR x trial evaluation loop of the reference (benchmark_comparison.py:149-261) as sharded launches of the fused
simulate -> reconstruct -> metrics kernel.

Frames are independent and a frame is a pure function of (seed, global frame index), so the sweep shards by contiguous
global frame ranges with no data-path exchange; the only collective is one all-reduce (sum) of the
[n_snr][4 methods][8 columns] double accumulator at the end (< 4 KB) - SURVEY.md 8(e).  The result is bit-identical in
its counts and equal to rounding in its sums for any number of ranks.
"""
from typing import Dict, List

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from ._lib import GEN_F32, METHOD_GAN, METHOD_MMSE, METHOD_NOEQ, METHOD_ZF, OfdmGanError

METHOD_ROWS = {"GAN": METHOD_GAN, "ZF": METHOD_ZF, "MMSE": METHOD_MMSE, "NoEQ": METHOD_NOEQ}


def shard_range(total: int, rank: int, world: int):
    """Contiguous global frame range [lo, hi) of `rank` (balanced to within one frame)."""
    return total * rank // world, total * (rank + 1) // world


def run_sweep(cfg, total_frames: int, gen_kind: int = GEN_F32, gparams=None, wrom=None, brom=None, slope: float = 0.2,
              seed: int = 0, frame0: int = 0, group=None, device=None, backend=ops, chunk: int = 1 << 26) -> np.ndarray:
    """Accumulator table [n_snr][N_METHODS][METRIC_COLS] (float64, host) over frames frame0..frame0+total_frames-1,
    summed over all ranks of `group` (every rank returns the same table)."""
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    lo, hi = shard_range(total_frames, rank, world)
    out = None
    for start in range(lo, hi, chunk):
        n = min(chunk, hi - start)
        out = backend.sim_gen_metrics(cfg, n, gen_kind=gen_kind, gparams=gparams, wrom=wrom, brom=brom, slope=slope, seed=seed,
                                      frame0=frame0 + start, device=device, out=out)
    if out is None:                                              # empty shard: still take part in the reduce
        out = backend.sim_gen_metrics(cfg, 0, gen_kind=gen_kind, gparams=gparams, wrom=wrom, brom=brom, slope=slope, seed=seed,
                                      frame0=frame0, device=device, out=None)
    if distributed:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out.cpu().numpy() if isinstance(out, torch.Tensor) else np.asarray(out)


def _is_arithmetic(v):
    if len(v) < 2:
        return True
    d = np.diff(np.asarray(v, dtype=np.float64))
    return bool(np.all(np.abs(d - d[0]) < 1e-9))


def run_benchmark(generator, n_trials: int = 100, frame_length: int = 16, snr_values: List[float] = (0, 5, 10, 15, 20, 25, 30),
                  channel_type: str = "awgn", nonlinear: bool = False, pa_saturation: float = 1.0, device=None, seed: int = 0,
                  group=None, backend=ops) -> Dict[str, Dict[float, Dict[str, float]]]:
    """Signature and return structure of benchmark_comparison.run_benchmark (method -> snr -> {'mse','mse_std','evm',
    'evm_std'}) for the methods computed on the GPU (GAN, ZF, MMSE, NoEQ - the data-parallel ones; DFE / LMS / RLS are serial
    adaptive filters and stay with the reference); the trial loop (SNR outer, trial inner, separate
    normalisation of noisy and clean, per-trial MSE / EVM in dB, mean and population std over trials) runs in one
    fused launch per rank.  `generator` is a MiniGenerator (its flat parameters are used) or a flat 258-vector."""
    if frame_length != 16:
        raise OfdmGanError("run_benchmark: only frame_length 16 is built (benchmark_comparison.py:355-472 default)")
    if channel_type not in ops.CHANNEL_TYPES:
        raise ValueError(f"Unknown channel type: {channel_type}")             # utils/ofdm_utils.py:673
    snr_values = [float(s) for s in snr_values]
    gparams = ops.flatten_params(generator) if isinstance(generator, torch.nn.Module) else generator
    slope = getattr(generator, "leaky_slope", 0.2)
    common = dict(nonlinear=nonlinear, pa_saturation=pa_saturation if nonlinear else 1.0, normalize=ops.NORM_SEPARATE,
                  snr_mode=ops.SNR_GRID, frames_per_snr=n_trials, equalizers=True, channel_type=channel_type)
    tables = []
    if _is_arithmetic(snr_values):
        step = snr_values[1] - snr_values[0] if len(snr_values) > 1 else 0.0
        cfg = ops.make_cfg(snr_lo=snr_values[0], snr_step=step, n_snr=len(snr_values), **common)
        t = run_sweep(cfg, n_trials * len(snr_values), GEN_F32, gparams=gparams, slope=slope, seed=seed, group=group, device=device,
                      backend=backend)
        tables = [t[i] for i in range(len(snr_values))]
    else:                                                        # arbitrary SNR list: one single-bin launch per value
        for i, s in enumerate(snr_values):
            cfg = ops.make_cfg(snr_lo=s, snr_step=0.0, n_snr=1, **common)
            t = run_sweep(cfg, n_trials, GEN_F32, gparams=gparams, slope=slope, seed=seed, frame0=i * n_trials, group=group,
                          device=device, backend=backend)
            tables.append(t[0])
    res = {m: {} for m in METHOD_ROWS}
    for snr, t in zip(snr_values, tables):
        s = ops.metrics_summary(t)
        for m, row in METHOD_ROWS.items():
            res[m][snr] = {"mse": float(s["mse"][row]), "mse_std": float(s["mse_std"][row]), "evm": float(s["evm"][row]),
                           "evm_std": float(s["evm_std"][row])}
    return res

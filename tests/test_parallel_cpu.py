"""Host-side logic of the multi-GPU paths on CPU: world_size-2 `gloo` process groups with the CUDA kernels replaced by
the oracle (tests may do that; the product never does).  Checks that frame sharding + the final all-reduce of the sweep,
and batch sharding + per-optimizer-step all-reduce of the trainer, reproduce the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from conftest import GOLDEN


class OracleBackend:
    """Stand-in for ofdm_gan_sr_b200.ops with the same call signatures, computing on the CPU oracle."""

    @staticmethod
    def sim_gen_metrics(cfg, B, gen_kind=0, gparams=None, wrom=None, brom=None, slope=0.2, seed=0, frame0=0, device=None, out=None):
        ocfg = oracle.ChanCfg.from_buffer_copy(bytes(cfg))
        n_snr = ocfg.n_snr if ocfg.snr_mode == 1 else 1
        if out is None:
            out = torch.zeros(n_snr, 4, 8, dtype=torch.float64)
        if B > 0:
            gp = None if gparams is None else np.asarray(gparams, dtype=np.float32)
            out += torch.from_numpy(oracle.sim_gen_metrics(ocfg, gen_kind, B, gparams=gp, wrom=wrom, brom=brom, slope=slope, seed=seed,
                                                           frame0=frame0))
        return out

    @staticmethod
    def gen_fwd_f32(x, gparams, slope=0.2):
        return torch.from_numpy(oracle.gen_fwd_f32(x.numpy(), gparams.numpy(), slope))

    @staticmethod
    def critic_step(clean, noisy, fake, dparams, alpha=None, seed=0, sample0=0, alpha_iter=0, gp_weight=10.0, slope=0.2,
                    b_global=None, out=None):
        B = clean.shape[0]
        if alpha is None:            # the documented stream: u = (x0 >> 8) * 2^-24 of Philox block (sample, iter, purpose 1)
            x = oracle.philox_blocks(seed, sample0, alpha_iter, 1, B)
            alpha = ((x[:, 0] >> 8).astype(np.float64) / 16777216.0).astype(np.float32)
        else:
            alpha = alpha.numpy()
        grads, stats = oracle.critic_step(clean.numpy(), noisy.numpy(), fake.numpy(), alpha, dparams.numpy(), gp_weight, slope)
        w = B / float(b_global or B)
        out.zero_()
        out[:521] = torch.from_numpy(grads) * w
        out[521:526] = torch.from_numpy(stats) * w
        return out

    @staticmethod
    def gen_step(clean, noisy, dparams, gparams, adv_weight=1.0, rec_weight=100.0, slope=0.2, b_global=None, out=None, fake_out=None,
                 fake=None):
        B = clean.shape[0]
        grads, stats, _ = oracle.gen_step(clean.numpy(), noisy.numpy(), dparams.numpy(), gparams.numpy(), adv_weight, rec_weight, slope)
        w = B / float(b_global or B)
        out.zero_()
        out[:258] = torch.from_numpy(grads) * w
        out[258:261] = torch.from_numpy(stats) * w
        return out

    @staticmethod
    def adam(p, m, v, g, lr, beta1, beta2, eps, step, grad_scale=1.0):
        pn, mn, vn = oracle.adam(p.numpy(), m.numpy(), v.numpy(), g.numpy()[:p.numel()], lr, beta1, beta2, eps, step, grad_scale)
        p.copy_(torch.from_numpy(pn)); m.copy_(torch.from_numpy(mn)); v.copy_(torch.from_numpy(vn))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import ofdm_gan_sr_b200 as pkg
        from ofdm_gan_sr_b200.sweep import run_benchmark, run_sweep, shard_range
        from ofdm_gan_sr_b200.train_step import CWGANGPStep
        r = dict(np.load(os.path.join(GOLDEN, "ref_fp32.npz")))
        # ---- sweep: frame-sharded, one final all-reduce
        kw = dict(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=50)
        cfg = pkg.ops.make_cfg(**kw)
        total = 7 * 50 * 3 + 17
        t = run_sweep(cfg, total, 0, gparams=r["gparams"], seed=4, backend=OracleBackend, chunk=300)
        res = run_benchmark(r["gparams"], n_trials=40, nonlinear=True, pa_saturation=0.8, seed=2, backend=OracleBackend)
        # ---- trainer: batch-sharded, one all-reduce per optimizer step
        B = 96
        lo, hi = shard_range(B, rank, world)
        assert hi - lo == B // world
        clean, noisy = torch.from_numpy(r["tr_clean"].reshape(-1, 2, 16)[:B]), torch.from_numpy(r["tr_noisy"].reshape(-1, 2, 16)[:B])
        tr = CWGANGPStep(r["tr_g0"], r["tr_d0"], seed=11, backend=OracleBackend)
        assert tr.world == world and tr.rank == rank
        for _ in range(2):
            tr.step(clean[lo:hi].contiguous(), noisy[lo:hi].contiguous())
        q.put((rank, t, res, tr.g.numpy().copy(), tr.d.numpy().copy(), tr.stats()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:      # surface the failure in the parent
        import traceback
        q.put((rank, "error", traceback.format_exc(), None, None, None))
        raise e


@pytest.mark.timeout(600)
def test_world_size_2_gloo_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, q)) for rk in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for g in got:
        assert g[1] is not None and not isinstance(g[1], str), g[2]
    got.sort(key=lambda g: g[0])

    import ofdm_gan_sr_b200 as pkg
    from ofdm_gan_sr_b200.sweep import run_benchmark, run_sweep
    from ofdm_gan_sr_b200.train_step import CWGANGPStep
    r = dict(np.load(os.path.join(GOLDEN, "ref_fp32.npz")))
    kw = dict(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=50)
    total = 7 * 50 * 3 + 17
    single = run_sweep(pkg.ops.make_cfg(**kw), total, 0, gparams=r["gparams"], seed=4, backend=OracleBackend)
    assert single[:, :2, 0].sum() == 2 * total
    res1 = run_benchmark(r["gparams"], n_trials=40, nonlinear=True, pa_saturation=0.8, seed=2, backend=OracleBackend)
    B = 96
    clean, noisy = torch.from_numpy(r["tr_clean"].reshape(-1, 2, 16)[:B]), torch.from_numpy(r["tr_noisy"].reshape(-1, 2, 16)[:B])
    tr = CWGANGPStep(r["tr_g0"], r["tr_d0"], seed=11, backend=OracleBackend)
    for _ in range(2):
        tr.step(clean, noisy)
    for rank, t, res, g, d, st in got:
        assert np.array_equal(t[:, :, 0], single[:, :, 0])                      # counts are exact
        np.testing.assert_allclose(t, single, rtol=1e-12, atol=0)               # sums equal up to addition order
        for m in ("GAN", "NoEQ"):
            for snr in res1[m]:
                for k in ("mse", "evm", "mse_std", "evm_std"):
                    assert abs(res[m][snr][k] - res1[m][snr][k]) <= 1e-9 * max(1.0, abs(res1[m][snr][k]))
        np.testing.assert_allclose(g, tr.g.numpy(), rtol=0, atol=2e-6 * np.abs(tr.g.numpy()).max())
        np.testing.assert_allclose(d, tr.d.numpy(), rtol=0, atol=2e-6 * np.abs(tr.d.numpy()).max())
        assert abs(st["d_loss"] - tr.stats()["d_loss"]) <= 1e-5 * max(1.0, abs(tr.stats()["d_loss"]))
    # both ranks hold identical replicas after the all-reduced updates
    assert np.array_equal(got[0][3], got[1][3]) and np.array_equal(got[0][4], got[1][4])


def test_shard_ranges_cover_everything():
    from ofdm_gan_sr_b200.sweep import shard_range
    for total in (0, 1, 7, 1 << 26, (1 << 26) + 5):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


class _FakeDataset:
    """stands in for SyntheticOFDMDataset (its batches are made on the GPU): records which frame ranges a loader asks for"""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def set_epoch(self, e):
        self.epoch = e

    def batch(self, lo, B):
        return (lo, B)


@pytest.mark.parametrize("n,bs,world", [(1000, 64, 1), (1000, 64, 4), (64 * 8 + 3, 64, 4), (64 * 8 + 37, 64, 8), (10, 64, 4)])
def test_loader_shards_are_equal_and_disjoint(n, bs, world):
    """every rank steps on the same local batch size (CWGANGPStep scales by 1 / (B_local x world)), also on a ragged last batch"""
    from ofdm_gan_sr_b200.utils.dataset import GPUBatchLoader
    for drop_last in (True, False):
        per_rank = [list(GPUBatchLoader(_FakeDataset(n), bs, drop_last=drop_last, rank=r, world_size=world)) for r in range(world)]
        assert len({len(p) for p in per_rank}) == 1                          # same number of steps everywhere
        assert all(len(p) == len(GPUBatchLoader(_FakeDataset(n), bs, drop_last=drop_last, rank=0, world_size=world)) for p in per_rank)
        covered = []
        for step in zip(*per_rank):
            assert len({B for _, B in step}) == 1 and step[0][1] > 0         # equal, non-empty shards
            covered += [(lo, lo + B) for lo, B in step]
        covered.sort()
        assert all(a[1] <= b[0] for a, b in zip(covered[:-1], covered[1:]))  # disjoint
        used = sum(hi - lo for lo, hi in covered)
        assert used <= n and (drop_last or n - used < world)                 # drop_last=False loses at most world-1 frames


def test_trainer_state_round_trips_and_maps_to_torch_adam():
    """state_dict / adam_state_dict: resume continues bit for bit, and the moments land in torch.optim.Adam's own layout
    (train.py:411-445 stores optimizer_G/D_state_dict)."""
    from ofdm_gan_sr_b200.train_step import CWGANGPStep
    r = dict(np.load(os.path.join(GOLDEN, "ref_fp32.npz")))
    B = 32
    clean, noisy = torch.from_numpy(r["tr_clean"].reshape(-1, 2, 16)[:B]), torch.from_numpy(r["tr_noisy"].reshape(-1, 2, 16)[:B])
    a = CWGANGPStep(r["tr_g0"], r["tr_d0"], seed=3, backend=OracleBackend)
    a.step(clean, noisy)
    sd = a.state_dict()
    b = CWGANGPStep(np.zeros(258, np.float32), np.zeros(521, np.float32), seed=3, backend=OracleBackend)
    b.load_state_dict(sd)
    a.step(clean, noisy)
    b.step(clean, noisy)
    assert torch.equal(a.g, b.g) and torch.equal(a.d, b.d) and torch.equal(a.d_v, b.d_v) and (a.d_steps, a.g_steps) == (b.d_steps, b.g_steps)
    # torch.optim.Adam accepts the exported state and a step from it equals the flat kernel's step
    G = torch.nn.ModuleList([torch.nn.Conv1d(2, 4, 3), torch.nn.Conv1d(4, 8, 3), torch.nn.Conv1d(8, 4, 3), torch.nn.Conv1d(4, 2, 3)])
    assert sum(p.numel() for p in G.parameters()) == 258
    a.store_to(generator=G)
    opt = torch.optim.Adam(G.parameters(), lr=a.lr_g, betas=a.betas, eps=a.eps)
    opt.load_state_dict(a.adam_state_dict("g", G))
    st = opt.state_dict()["state"]
    assert int(st[0]["step"]) == a.g_steps == 2
    assert torch.equal(torch.cat([st[i]["exp_avg"].reshape(-1) for i in sorted(st)]), a.g_m)
    c = CWGANGPStep(np.zeros(258, np.float32), np.zeros(521, np.float32), backend=OracleBackend)
    c.load_adam_state_dict("g", opt.state_dict())
    assert torch.equal(c.g_m, a.g_m) and torch.equal(c.g_v, a.g_v) and c.g_steps == 2 and c.lr_g == a.lr_g
    with pytest.raises(Exception):
        a.step(clean[:0], noisy[:0])                                        # an empty local batch is an error, not a silent no-op

"""Multi-GPU parity checks of the data-parallel paths (one process per GPU over torch.distributed / NCCL), shared by
tests/multi_gpu_worker.py (pytest, needs >= 2 visible GPUs) and bench.py (which runs them OUTSIDE its timed regions whenever it is
launched on more than one GPU, so that the driver's scaling run carries the parity verdict):
  1. the fused peer-memory exchange alone: the sum is the rank-ordered sum bit for bit, every rank ends with the same bits, the
     fused Adam equals the stand-alone kernel (replaces all_reduce + optimizer.step() of train.py:253,297 under data parallelism);
  2. the training step: peer exchange == NCCL exchange == ONE GPU on the whole batch to 2e-5, replayed CUDA graph == eager;
  3. the frame-sharded sweep == the whole frame range on one GPU (counts exact), every rank holds the same table.
Raises AssertionError on the first mismatch; returns a dict of what was checked."""
import numpy as np
import torch
import torch.distributed as dist


def run_all(dev, rank, world):
    import ofdm_gan_sr_b200 as pkg
    from ofdm_gan_sr_b200.train_step import CWGANGPStep
    ops = pkg.ops
    worst = {}

    # ---- 1. the exchange alone: rank-ordered sum, identical bits on every rank, Adam applied to the prefix
    comm = ops.PeerComm(None, dev)
    rng = np.random.default_rng(1234)
    msgs = rng.standard_normal((world, 7, 528)).astype(np.float32)           # every rank knows every rank's messages
    p0 = rng.standard_normal(521).astype(np.float32)
    p, m, v = torch.as_tensor(p0).to(dev), torch.zeros(521, device=dev), torch.zeros(521, device=dev)
    pr, mr, vr = p.clone(), m.clone(), v.clone()
    for it in range(7):
        g = torch.as_tensor(msgs[rank, it]).to(dev)
        comm.allreduce_adam(g, p, m, v, 2e-4, 0.0, 0.9, 1e-8, it + 1)
        want = torch.zeros(528, device=dev)
        for r in range(world):                                               # the kernel's order: rank 0, 1, ...
            want = want + torch.as_tensor(msgs[r, it]).to(dev)
        assert torch.equal(g, want), (rank, it, float((g - want).abs().max()))
        ops.adam(pr, mr, vr, want, 2e-4, 0.0, 0.9, 1e-8, it + 1)
        assert torch.equal(p, pr) and torch.equal(m, mr) and torch.equal(v, vr)
    plain = torch.full((264,), float(rank + 1), device=dev)
    comm.allreduce_adam(plain)                                               # n_params = 0: plain all-reduce
    assert torch.equal(plain, torch.full((264,), world * (world + 1) / 2.0, device=dev))
    comm.check()
    everyone = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(everyone, p)
    assert all(torch.equal(e, everyone[0]) for e in everyone)                # replicas bit-identical
    comm.close()

    # ---- 2. the training step: peer exchange == NCCL exchange == one GPU on the whole batch (to rounding)
    B = 2048
    cfg = ops.make_cfg(normalize=1, snr_lo=0.0, snr_hi=30.0)
    clean_all, noisy_all, _ = ops.chan_sim(cfg, B * world, seed=5, frame0=0, device=dev)
    clean, noisy = clean_all[rank * B:(rank + 1) * B].contiguous(), noisy_all[rank * B:(rank + 1) * B].contiguous()
    gp = (np.random.default_rng(7).standard_normal(258) * 0.3).astype(np.float32)
    dp = (np.random.default_rng(8).standard_normal(521) * 0.2).astype(np.float32)
    runs = {}
    for mode in ("peer", "nccl", "peer+graph"):
        t = CWGANGPStep(gp, dp, device=dev, exchange=mode.split("+")[0], graph=mode.endswith("graph"))
        assert (t.comm is not None) == mode.startswith("peer") and t.use_graph == mode.endswith("graph")
        for _ in range(4):
            t.step(clean, noisy)
        runs[mode] = (t.g.clone(), t.d.clone(), t.stats())
        t.close()
    solo = CWGANGPStep(gp, dp, device=dev, process_group=None, exchange="nccl")
    solo.distributed, solo.world, solo.rank = False, 1, 0                    # the whole batch on this GPU, no exchange
    for _ in range(4):
        solo.step(clean_all, noisy_all)
    for name, (g, d, st) in runs.items():
        for a, b, what in ((g, solo.g, "G"), (d, solo.d, "D")):
            err = float((a - b).abs().max()) / float(b.abs().max())
            worst[name + ":" + what] = err
            assert err < 2e-5, (name, what, err)
        assert abs(st["d_loss"] - solo.stats()["d_loss"]) < 1e-4 * max(1.0, abs(solo.stats()["d_loss"]))
    err = float((runs["peer"][0] - runs["nccl"][0]).abs().max())
    assert err < 1e-6, err                                                   # same sums up to the order NCCL happens to use
    err = float((runs["peer+graph"][1] - runs["peer"][1]).abs().max()) / float(runs["peer"][1].abs().max())
    assert err < 1e-6, err                                                   # replayed graph == eager launches
    # ---- 3. the sweep: frame-sharded over the ranks == the whole range on one GPU (counts exact, sums to rounding)
    from ofdm_gan_sr_b200.sweep import run_benchmark, run_sweep
    cfg = ops.make_cfg(nonlinear=True, pa_saturation=0.8, normalize=1, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=1000)
    total = 7 * 1000 * 37 + 13                                               # ragged on purpose
    sharded = run_sweep(cfg, total, gparams=gp, seed=9, device=dev)          # all ranks, all-reduced
    whole = ops.sim_gen_metrics(cfg, total, gparams=gp, seed=9, device=dev).cpu().numpy()
    assert np.array_equal(sharded[:, :, 0], whole[:, :, 0]) and sharded[:, :2, 0].sum() == 2 * total
    assert np.allclose(sharded[:, :2, 1:5], whole[:, :2, 1:5], rtol=5e-6, atol=0)      # fp32 per-thread partial sums regroup
    res = run_benchmark(gp, n_trials=5000, nonlinear=True, pa_saturation=0.8, device=dev, seed=4)
    every = [None] * world
    dist.all_gather_object(every, {m: {s: v["evm"] for s, v in res[m].items()} for m in res})
    assert all(e == every[0] for e in every)                                 # every rank holds the same table

    dist.barrier()
    return {"exchange_bit_exact": True, "replicas_identical": True, "train_vs_single_gpu_rel_err": max(worst.values()),
            "train_vs_single_gpu_tol": 2e-5, "sweep_counts_exact": True, "world": world}

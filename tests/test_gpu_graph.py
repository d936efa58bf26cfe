"""GPU: the trainer iteration replayed as one CUDA graph (device-resident step counters) equals the eager iteration."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import ofdm_gan_sr_b200 as pkg
    from ofdm_gan_sr_b200.train_step import CWGANGPStep
    assert torch.cuda.is_available()
    gp = (np.random.default_rng(7).standard_normal(258) * 0.3).astype(np.float32)
    dp = (np.random.default_rng(8).standard_normal(521) * 0.2).astype(np.float32)
    return pkg.ops, CWGANGPStep, gp, dp


def _batches(ops, B, n, seed=3):
    cfg = ops.make_cfg(normalize=1, snr_lo=0.0, snr_hi=30.0)
    return [ops.chan_sim(cfg, B, seed=seed, frame0=i * B)[:2] for i in range(n)]


def test_graph_replay_equals_eager(setup):
    ops, Step, gp, dp = setup
    data = _batches(ops, 4096, 6)
    eager, graph = Step(gp, dp), Step(gp, dp, graph=True)
    for clean, noisy in data:
        eager.step(clean, noisy)
        graph.step(clean, noisy)
    assert 1 <= len(graph._graphs) <= 5 and graph.d_steps == eager.d_steps == 30 and graph.g_steps == 6
    assert graph._ctr.tolist() == [30, 6]                      # the device counters advanced with every replay
    for a, b in ((graph.g, eager.g), (graph.d, eager.d), (graph.g_m, eager.g_m), (graph.d_v, eager.d_v)):
        # same alphas, same kernels; Adam's bias correction comes from beta^t by squaring on the device vs pow() on the host
        assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())
    se, sg = eager.stats(), graph.stats()
    for k in ("d_loss", "gradient_penalty", "g_loss", "rec_loss"):
        assert abs(se[k] - sg[k]) <= 1e-5 * max(1.0, abs(se[k]))


def test_graph_recaptures_on_shape_change_and_rejects_injected_alpha(setup):
    ops, Step, gp, dp = setup
    t = Step(gp, dp, graph=True)
    ref = Step(gp, dp)
    for B in (1024, 1024, 1024, 2048, 2048, 2048, 1024, 1024):
        clean, noisy = _batches(ops, B, 1, seed=B)[0]
        t.step(clean, noisy)
        ref.step(clean, noisy)
    assert float((t.d - ref.d).abs().max()) <= 1e-6 * float(ref.d.abs().max())
    with pytest.raises(Exception):
        t.step(clean, noisy, alphas=torch.rand(5, 1024, device="cuda"))


def test_adam_ctr_matches_host_stepped_adam(setup):
    ops = setup[0]
    rng = np.random.default_rng(0)
    p0 = torch.as_tensor(rng.standard_normal(521).astype(np.float32)).cuda()
    pa, ma, va = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    pb, mb, vb = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t in range(1, 41):
        g = torch.as_tensor(rng.standard_normal(521).astype(np.float32)).cuda()
        ops.adam(pa, ma, va, g, 2e-4, 0.5, 0.9, 1e-8, t)
        ops.adam(pb, mb, vb, g, 2e-4, 0.5, 0.9, 1e-8, 0, step_dev=ctr)
    assert int(ctr) == 40
    assert float((pa - pb).abs().max()) <= 1e-6 * float(pa.abs().max())
    assert torch.equal(ma, mb) and torch.equal(va, vb)         # the moments do not depend on the bias corrections


def test_asynchronous_stats_equal_synchronous(setup):
    ops, Step, gp, dp = setup
    t = Step(gp, dp)
    tickets = []
    want = []
    for clean, noisy in _batches(ops, 2048, 3):
        t.step(clean, noisy)
        tickets.append(t.request_stats())
        want.append(t.stats())
        got = t.collect_stats(tickets[-1])
        assert got == want[-1]


def test_fused_critic_iteration_is_bitwise_the_unfused_one(setup):
    """ofdmgan_critic_train_ctr (kernel + one tail launch) == ofdmgan_critic_step_ctr + ofdmgan_adam_ctr, bit for bit, over several
    iterations, with and without the image_is_current shortcut."""
    ops, _, gp, dp = setup
    clean, noisy = _batches(ops, 3000, 1)[0]
    fake = ops.gen_fwd_f32(noisy, gp)
    dev = clean.device
    d0 = torch.as_tensor(dp).to(dev)
    a = dict(d=d0.clone(), m=torch.zeros_like(d0), v=torch.zeros_like(d0), ctr=torch.zeros(1, dtype=torch.int32, device=dev))
    b = dict(d=d0.clone(), m=torch.zeros_like(d0), v=torch.zeros_like(d0), ctr=torch.zeros(1, dtype=torch.int32, device=dev))
    for it in range(6):
        oa = ops.critic_step(clean, noisy, fake, a["d"], seed=11, gp_weight=10.0, alpha_iter_dev=a["ctr"])
        ops.adam(a["d"], a["m"], a["v"], oa, 2e-4, 0.0, 0.9, 1e-8, 0, step_dev=a["ctr"])
        ob = ops.critic_train(clean, noisy, fake, b["d"], b["m"], b["v"], b["ctr"], 2e-4, 0.0, 0.9, 1e-8, seed=11, gp_weight=10.0,
                              image_is_current=it % 3 != 0)
        assert torch.equal(oa, ob), it
        for k in ("d", "m", "v", "ctr"):
            assert torch.equal(a[k], b[k]), (it, k)
    assert int(b["ctr"]) == 6


def test_graph_is_recaptured_when_a_hyper_parameter_changes():
    """the learning rate is a host scalar baked into the captured graph: assigning lr_d / lr_g (StepLR, train.py:497-514) must take
    effect on the next step - graph replay == eager launches with the same schedule"""
    import ofdm_gan_sr_b200 as pkg
    from ofdm_gan_sr_b200.train_step import CWGANGPStep
    ops = pkg.ops
    rng = np.random.default_rng(3)
    gp, dp = (rng.standard_normal(258) * 0.3).astype(np.float32), (rng.standard_normal(521) * 0.2).astype(np.float32)
    clean, noisy, _ = ops.chan_sim(ops.make_cfg(), 512, seed=2)
    runs = []
    for graph in (True, False):
        t = CWGANGPStep(gp, dp, graph=graph, seed=5)
        for i in range(6):
            if i == 3:
                t.lr_d, t.lr_g = t.lr_d * 0.5, t.lr_g * 0.5
            t.step(clean, noisy)
        runs.append((t.g.clone(), t.d.clone()))
    for a, b in zip(*runs):
        assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())
    # and the halved rate really was used: a run that never halves ends elsewhere
    t = CWGANGPStep(gp, dp, graph=True, seed=5)
    for i in range(6):
        t.step(clean, noisy)
    assert float((t.d - runs[0][1]).abs().max()) > 1e-5 * float(t.d.abs().max())


def test_graphs_are_keyed_by_the_callers_buffers():
    """a double-buffered loader alternates between two buffer pairs: two graphs, no copies, same result as eager launches"""
    import ofdm_gan_sr_b200 as pkg
    from ofdm_gan_sr_b200.train_step import CWGANGPStep
    ops = pkg.ops
    rng = np.random.default_rng(4)
    gp, dp = (rng.standard_normal(258) * 0.3).astype(np.float32), (rng.standard_normal(521) * 0.2).astype(np.float32)
    bufs = [ops.chan_sim(ops.make_cfg(), 256, seed=s)[:2] for s in (1, 2, 3)]
    g, e = CWGANGPStep(gp, dp, graph=True, seed=2), CWGANGPStep(gp, dp, graph=False, seed=2)
    order = [0, 1, 0, 1, 2, 0, 1, 2, 2, 0]
    for i in order:
        g.step(*bufs[i])
        e.step(*bufs[i])
    assert len(g._graphs) == 4                                  # the static pair (first sightings) + one per recurring buffer pair
    for a, b in ((g.g, e.g), (g.d, e.d)):
        assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())
    # new contents in an old buffer are what the replay reads
    bufs[0][0].copy_(bufs[2][0]); bufs[0][1].copy_(bufs[2][1])
    g.step(*bufs[0]); e.step(*bufs[2])
    assert float((g.d - e.d).abs().max()) <= 1e-6 * float(e.d.abs().max())

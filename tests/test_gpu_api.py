"""GPU parity at the level of the reference's Python call surface: the drop-in nn.Modules,
compute_gradient_penalty, SyntheticOFDMDataset, run_benchmark and the fused trainer step, against a plain PyTorch fp32
restatement of the same architecture running stock ATen / autograd on the same device (<= 1e-5 relative)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import assert_close

pytestmark = pytest.mark.gpu
TOL = 1e-5


# ---- plain PyTorch restatement (test-side reference; same parameter names as models/generator.py / discriminator.py)
class TorchG(nn.Module):
    def __init__(self):
        super().__init__()
        blk = lambda i, o, s: nn.ModuleDict({"conv": nn.Conv1d(i, o, 3, stride=s, padding=1)})
        self.enc1, self.bottleneck, self.dec1 = blk(2, 4, 2), blk(4, 8, 2), blk(8, 4, 1)
        self.out_conv = nn.Conv1d(4, 2, 3, padding=1)

    def forward(self, x):
        e = F.leaky_relu(self.enc1["conv"](x), 0.2)
        b = F.leaky_relu(self.bottleneck["conv"](e), 0.2)
        d = F.leaky_relu(self.dec1["conv"](F.interpolate(b, scale_factor=2, mode="nearest")), 0.2)
        return torch.tanh(self.out_conv(F.interpolate(d + e, scale_factor=2, mode="nearest")))


class TorchD(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1, self.conv2, self.dense = nn.Conv1d(4, 8, 3, stride=2, padding=1), nn.Conv1d(8, 16, 3, stride=2, padding=1), nn.Linear(16, 1)

    def forward(self, cand, cond):
        h = F.leaky_relu(self.conv1(torch.cat([cand, cond], 1)), 0.2)
        return self.dense(F.leaky_relu(self.conv2(h), 0.2).sum(2))


def torch_gp(D, real, fake, cond):
    alpha = torch.rand(real.size(0), 1, 1, device=real.device)
    xh = (alpha * real + (1 - alpha) * fake).requires_grad_(True)
    g = torch.autograd.grad(D(xh, cond), xh, torch.ones(real.size(0), 1, device=real.device), create_graph=True)[0]
    return ((g.view(real.size(0), -1).norm(2, dim=1) - 1) ** 2).mean()


@pytest.fixture(scope="module")
def pkg():
    import ofdm_gan_sr_b200 as p
    import ofdm_gan_sr_b200.models  # noqa: F401
    import ofdm_gan_sr_b200.utils  # noqa: F401
    assert torch.cuda.is_available()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return p


def _pair(pkg, seed=0):
    torch.manual_seed(seed)
    G, D = pkg.models.MiniGenerator().cuda(), pkg.models.MiniDiscriminator().cuda()
    with torch.no_grad():                                   # non-zero biases make the check stronger
        for p in list(G.parameters()) + list(D.parameters()):
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
    TG, TD = TorchG().cuda(), TorchD().cuda()
    TG.load_state_dict(G.state_dict())                      # same names and shapes: checkpoints interchange
    TD.load_state_dict(D.state_dict())
    return G, D, TG, TD


def _grads(mod):
    # a parameter autograd never reached (the penalty has no bias gradient) counts as zero
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in mod.parameters()]).cpu().numpy()


def test_module_surface(pkg):
    G, D = pkg.models.MiniGenerator(), pkg.models.MiniDiscriminator()
    assert G.count_parameters() == 258 and D.count_parameters() == 521
    assert G.estimate_macs() == 1728 and D.estimate_macs() == 2384
    assert list(G.state_dict()) == ["enc1.conv.weight", "enc1.conv.bias", "bottleneck.conv.weight", "bottleneck.conv.bias",
                                    "dec1.conv.weight", "dec1.conv.bias", "out_conv.weight", "out_conv.bias"]
    assert list(D.state_dict()) == ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "dense.weight", "dense.bias"]
    assert pkg.models.UNetGenerator is pkg.models.MiniGenerator and pkg.models.Discriminator is pkg.models.MiniDiscriminator
    assert [l["name"] for l in G.get_layer_info()][:2] == ["enc1", "bottleneck"]
    assert all(float(p.detach().abs().sum()) == 0 for n, p in G.named_parameters() if n.endswith("bias"))     # zero-bias init
    with pytest.raises(pkg.OfdmGanError):
        G(torch.zeros(2, 2, 16))                             # CPU tensors: no fallback
    with pytest.raises(pkg.OfdmGanError):
        pkg.models.MiniGenerator(frame_length=32).cuda()(torch.zeros(2, 2, 32).cuda())


def test_generator_forward_backward_vs_autograd(pkg):
    G, D, TG, TD = _pair(pkg, 1)
    x = torch.randn(777, 2, 16, device="cuda", requires_grad=True)
    xt = x.detach().clone().requires_grad_(True)
    w = torch.randn(777, 2, 16, device="cuda")
    y, yt = G(x), TG(xt)
    assert y.shape == (777, 2, 16)
    assert_close(y.detach().cpu().numpy(), yt.detach().cpu().numpy(), TOL, "G forward")
    (y * w).sum().backward()
    (yt * w).sum().backward()
    assert_close(x.grad.cpu().numpy(), xt.grad.cpu().numpy(), TOL, "G dx")
    assert_close(_grads(G), _grads(TG), TOL, "G dparams")


def test_critic_forward_backward_vs_autograd(pkg):
    G, D, TG, TD = _pair(pkg, 2)
    a = torch.randn(500, 2, 16, device="cuda", requires_grad=True)
    c = torch.randn(500, 2, 16, device="cuda", requires_grad=True)
    at, ct = a.detach().clone().requires_grad_(True), c.detach().clone().requires_grad_(True)
    w = torch.randn(500, 1, device="cuda")
    s, st = D(a, c), TD(at, ct)
    assert s.shape == (500, 1)
    assert_close(s.detach().cpu().numpy(), st.detach().cpu().numpy(), TOL, "D forward")
    (s * w).sum().backward()
    (st * w).sum().backward()
    assert_close(a.grad.cpu().numpy(), at.grad.cpu().numpy(), TOL, "D dcand")
    assert_close(c.grad.cpu().numpy(), ct.grad.cpu().numpy(), TOL, "D dcond")
    assert_close(_grads(D), _grads(TD), TOL, "D dparams")


def test_gradient_penalty_vs_autograd_double_backward(pkg):
    G, D, TG, TD = _pair(pkg, 3)
    real, fake, cond = (torch.randn(640, 2, 16, device="cuda") for _ in range(3))
    torch.manual_seed(123)
    gp = pkg.models.compute_gradient_penalty(D, real, fake, cond, device=real.device)
    torch.manual_seed(123)                                   # same torch.rand draw
    gpt = torch_gp(TD, real, fake, cond)
    assert gp.dim() == 0
    assert abs(float(gp.detach()) - float(gpt.detach())) <= TOL * abs(float(gpt.detach()))
    (10.0 * gp).backward()
    (10.0 * gpt).backward()
    assert_close(_grads(D), _grads(TD), TOL, "GP dparams")


def test_reference_training_iteration_through_the_modules(pkg):
    """train.py:201-305 written against the drop-in modules (autograd.Function path) == the same code on stock torch."""
    G, D, TG, TD = _pair(pkg, 4)
    B = 256
    clean = torch.rand(B, 2, 16, device="cuda") * 2 - 1
    noisy = (clean + 0.2 * torch.randn_like(clean)).clamp(-1, 1)

    def iteration(Gm, Dm, gp_fn):
        oD = torch.optim.Adam(Dm.parameters(), lr=2e-4, betas=(0.0, 0.9))
        oG = torch.optim.Adam(Gm.parameters(), lr=2e-4, betas=(0.0, 0.9))
        torch.manual_seed(7)
        for _ in range(5):
            oD.zero_grad()
            with torch.no_grad():
                fake = Gm(noisy)
            d_loss = Dm(fake, noisy).mean() - Dm(clean, noisy).mean() + 10.0 * gp_fn(Dm, clean, fake, noisy)
            d_loss.backward()
            oD.step()
        oG.zero_grad()
        fake = Gm(noisy)
        g_loss = -Dm(fake, noisy).mean() + 100.0 * F.l1_loss(fake, clean)
        g_loss.backward()
        oG.step()
        return float(d_loss.detach()), float(g_loss.detach())

    l1 = iteration(G, D, lambda Dm, r, f, c: pkg.models.compute_gradient_penalty(Dm, r, f, c))
    l2 = iteration(TG, TD, torch_gp)
    assert abs(l1[0] - l2[0]) <= 2e-5 * max(1, abs(l2[0])) and abs(l1[1] - l2[1]) <= 2e-5 * max(1, abs(l2[1]))
    for (n, p), (_, q) in zip(list(G.named_parameters()) + list(D.named_parameters()),
                              list(TG.named_parameters()) + list(TD.named_parameters())):
        assert_close(p.detach().cpu().numpy(), q.detach().cpu().numpy(), TOL, n)


def test_fused_trainer_step_vs_torch_eager(pkg):
    from ofdm_gan_sr_b200.train_step import CWGANGPStep
    G, D, TG, TD = _pair(pkg, 5)
    B = 512
    clean = torch.rand(B, 2, 16, device="cuda") * 2 - 1
    noisy = (clean + 0.2 * torch.randn_like(clean)).clamp(-1, 1)
    tr = CWGANGPStep(G, D)
    oD = torch.optim.Adam(TD.parameters(), lr=2e-4, betas=(0.0, 0.9))
    oG = torch.optim.Adam(TG.parameters(), lr=2e-4, betas=(0.0, 0.9))
    for it in range(2):
        alphas = torch.rand(5, B, device="cuda")
        tr.step(clean, noisy, alphas=alphas)
        for c in range(5):
            oD.zero_grad()
            with torch.no_grad():
                fake = TG(noisy)
            a = alphas[c].view(B, 1, 1)
            xh = (a * clean + (1 - a) * fake).requires_grad_(True)
            g = torch.autograd.grad(TD(xh, noisy), xh, torch.ones(B, 1, device="cuda"), create_graph=True)[0]
            gp = ((g.view(B, -1).norm(2, dim=1) - 1) ** 2).mean()
            d_loss = TD(fake, noisy).mean() - TD(clean, noisy).mean() + 10.0 * gp
            d_loss.backward()
            oD.step()
        oG.zero_grad()
        fake = TG(noisy)
        g_loss = -TD(fake, noisy).mean() + 100.0 * F.l1_loss(fake, clean)
        g_loss.backward()
        oG.step()
        st = tr.stats()
        assert abs(st["d_loss"] - float(d_loss.detach())) <= 2e-5 * max(1, abs(float(d_loss.detach())))
        assert abs(st["g_loss"] - float(g_loss.detach())) <= 2e-5 * max(1, abs(float(g_loss.detach())))
    tr.store_to(G, D)
    for (n, p), (_, q) in zip(list(G.named_parameters()) + list(D.named_parameters()),
                              list(TG.named_parameters()) + list(TD.named_parameters())):
        assert_close(p.detach().cpu().numpy(), q.detach().cpu().numpy(), TOL, n)


def test_synthetic_dataset_surface(pkg):
    DS = pkg.utils.SyntheticOFDMDataset
    ds = DS(n_samples=1000, snr_range=(5, 20), nonlinear=True, pa_saturation=0.8, seed=3)
    assert len(ds) == 1000
    b = ds.batch(0, 1000)
    assert b["noisy"].shape == (1000, 2, 16) and b["clean"].shape == (1000, 2, 16) and b["snr"].shape == (1000,)
    assert b["noisy"].is_cuda and b["noisy"].dtype == torch.float32
    assert float(b["snr"].min()) >= 5 and float(b["snr"].max()) < 20
    peak = torch.maximum(b["noisy"].abs().amax(dim=(1, 2)), b["clean"].abs().amax(dim=(1, 2)))
    assert torch.all((peak - 1).abs() < 1e-6)                 # joint max-abs normalisation (utils/dataset.py:284-287)
    item = ds[17]
    assert torch.equal(item["noisy"], b["noisy"][17]) and torch.equal(item["clean"], b["clean"][17]) and item["snr"].dim() == 0
    with pytest.raises(IndexError):
        ds[1000]
    loader = pkg.utils.create_dataloader(ds, batch_size=64, shuffle=True, num_workers=0, drop_last=True)
    batches = list(loader)
    assert len(batches) == len(loader) == 15 and all(x["noisy"].shape == (64, 2, 16) for x in batches)
    again = list(loader)                                     # new epoch -> new frames, like the reference's fresh draws
    assert not torch.equal(again[0]["noisy"], batches[0]["noisy"])
    # two ranks see disjoint halves of the same global batch
    r0 = next(iter(pkg.utils.create_dataloader(DS(n_samples=256, seed=9), batch_size=64, rank=0, world_size=2)))
    r1 = next(iter(pkg.utils.create_dataloader(DS(n_samples=256, seed=9), batch_size=64, rank=1, world_size=2)))
    whole = DS(n_samples=256, seed=9).batch(0, 128)
    assert torch.equal(torch.cat([r0["noisy"], r1["noisy"]]), whole["noisy"])
    with pytest.raises(ValueError):
        DS(channel_type="no-such-channel")
    ts = pkg.utils.generate_test_samples(n_samples=50, snr_values=[5, 20])
    assert list(ts) == [5, 20] and len(ts[5]) == 50 and ts[20][3]["snr"] == 20 and ts[5][0]["noisy"].shape == (2, 16)
    err = lambda k: float(torch.stack([(d["noisy"] - d["clean"]).pow(2).mean() for d in ts[k]]).mean())
    assert err(5) > 5 * err(20)                                # less noise at the higher SNR


def test_run_benchmark_surface_and_consistency(pkg):
    from ofdm_gan_sr_b200.sweep import run_benchmark
    G, D, TG, TD = _pair(pkg, 6)
    res = run_benchmark(G, n_trials=2000, nonlinear=True, pa_saturation=0.8, seed=1)
    assert set(res) == {"GAN", "ZF", "MMSE", "NoEQ"} and list(res["GAN"]) == [0.0, 5.0, 10.0, 15.0, 20.0, 25.0, 30.0]
    assert set(res["GAN"][0.0]) == {"mse", "mse_std", "evm", "evm_std"}
    evm = [res["NoEQ"][s]["evm"] for s in res["NoEQ"]]
    assert all(a > b for a, b in zip(evm[:3], evm[1:4]))     # NoEQ EVM falls with SNR while noise dominates (the PA floor is near -5 dB)
    # the same frames through the unfused public pieces and the stock-torch generator
    ops = pkg.ops
    cfg = ops.make_cfg(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=2000)
    clean, noisy, _ = ops.chan_sim(cfg, 14000, seed=1)
    with torch.no_grad():
        est = TG(noisy)
    e = ((est - clean) ** 2).sum(dim=(1, 2)).double()
    r = (clean ** 2).sum(dim=(1, 2)).double()
    evm_db = (20 * torch.log10(torch.sqrt(e / r) + 1e-10)).view(7, 2000)
    mse = (e / 32).view(7, 2000)
    for i, s in enumerate(res["GAN"]):
        assert abs(res["GAN"][s]["evm"] - float(evm_db[i].mean())) <= 2e-5 * abs(float(evm_db[i].mean()))
        assert abs(res["GAN"][s]["mse"] - float(mse[i].mean())) <= 2e-5 * float(mse[i].mean())
        assert abs(res["GAN"][s]["evm_std"] - float(evm_db[i].std(unbiased=False))) <= 1e-3 * float(evm_db[i].std(unbiased=False))
    irregular = run_benchmark(G, n_trials=500, snr_values=[3, 4, 12], seed=1)
    assert list(irregular["GAN"]) == [3.0, 4.0, 12.0]


def test_q_rom_export_and_integer_inference(pkg):
    import oracle
    G, D, TG, TD = _pair(pkg, 7)
    wrom, brom = pkg.utils.export_q_roms(G)
    assert wrom.dtype == np.int8 and wrom.shape == (2048,) and brom.dtype == np.int16 and brom.shape == (64,)
    sd = G.state_dict()
    w = sd["bottleneck.conv.weight"].cpu()
    ref = pkg.utils.quantize_tensor(w, torch.tensor(1 / 128.0), 8).numpy().astype(np.int8).reshape(-1)
    assert np.array_equal(wrom[24:120], ref)
    assert np.array_equal(wrom[216:224], np.clip(np.rint(sd["out_conv.weight"].cpu().numpy()[:, :, 1] * 128), -128, 127).astype(np.int8).reshape(-1))
    x = torch.randn(4096, 2, 16, device="cuda").clamp(-1, 1)
    xq = pkg.utils.float_to_q88(x)
    for mode, omode in (("spec", 0), ("rtl_literal", 1)):
        yq = G.forward_q88(xq, wrom, brom, mode=mode)
        assert np.array_equal(yq.cpu().numpy(), oracle.gen_fwd_q(xq.cpu().numpy(), wrom, brom, omode))
    # the clean-dataflow integer model tracks the float model it was exported from (1x1 output conv, clip instead of tanh
    # and a 0.3125 LeakyReLU make it an approximation, not a bit-level twin)
    assert pkg.utils.q88_to_float(xq).sub(x).abs().max() <= 1 / 256


def test_quantize_helpers_on_the_device_match_reference(ref_channel):
    """compute_scale / quantize_tensor / dequantize_tensor (utils/quantization.py:73-161) as libofdmgan kernels, bit-exact against the
    values recorded from the reference, for per-tensor and per-channel scales and for a channel dimension that is not the first."""
    import ofdm_gan_sr_b200.utils as utils
    r = ref_channel
    t = torch.as_tensor(r["qt_in"]).cuda()
    assert np.array_equal(utils.quantize_tensor(t, torch.tensor(1.0 / 128), 8).cpu().numpy(), r["qt_q17"])
    scale = utils.compute_scale(t, 8)
    assert scale.is_cuda and scale.dim() == 0 and float(scale) == float(r["qt_scale"])
    q = utils.quantize_tensor(t, scale, 8)
    assert np.array_equal(q.cpu().numpy(), r["qt_q8"])
    assert np.array_equal(utils.dequantize_tensor(q, scale).cpu().numpy(), r["qt_deq"])
    w = torch.as_tensor(r["qt_w"]).cuda()
    ws = utils.compute_scale(w, 8, per_channel=True, channel_dim=0)
    assert ws.shape == r["qt_w_scale"].shape and np.array_equal(ws.cpu().numpy(), r["qt_w_scale"])
    assert np.array_equal(utils.quantize_tensor(w, ws, 8).cpu().numpy(), r["qt_w_q"])
    # another channel dimension, large tensor: the same numbers as the host expression
    x = torch.randn(6, 40, 33, device="cuda") * 3
    for dim in (0, 1, 2):
        s_dev = utils.compute_scale(x, 6, per_channel=True, channel_dim=dim)
        s_host = utils.compute_scale(x.cpu(), 6, per_channel=True, channel_dim=dim)
        assert torch.equal(s_dev.cpu(), s_host)
        q_dev, q_host = utils.quantize_tensor(x, s_dev, 6), utils.quantize_tensor(x.cpu(), s_host, 6)
        assert torch.equal(q_dev.cpu(), q_host)
        assert torch.equal(utils.dequantize_tensor(q_dev, s_dev).cpu(), utils.dequantize_tensor(q_host, s_host))
    fq = utils.FakeQuantize(8, per_channel=False).cuda().train()
    y = fq(t.clone().requires_grad_(True))
    assert torch.allclose(y.detach().cpu(), torch.as_tensor(r["qt_deq"]), atol=1e-7)


def test_inference_calls_on_two_streams_are_independent(pkg):
    """Host-resident weights travel by value: no constant-image refresh, no library lock, per-stream scratch.  A call on stream B
    completes while stream A is still busy - with the weights in device memory (the training path) B would be ordered after A."""
    ops = pkg.ops
    rng = np.random.default_rng(0)
    gp = (rng.standard_normal(258) * 0.3).astype(np.float32)
    x = torch.randn(8192, 2, 16, device="cuda")
    cfg = ops.make_cfg(nonlinear=True, pa_saturation=0.8, normalize=1, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=64)
    ref_y = ops.gen_fwd_f32(x, gp)
    ref_m = ops.sim_gen_metrics(cfg, 7 * 64 * 9, gparams=gp, seed=3)
    W = rng.integers(-128, 128, 2048).astype(np.int8)
    Bq = np.zeros(64, np.int16)
    xq = ops.quantize_q88(x)
    ref_q = ops.gen_fwd_q(xq, W, Bq, mode=ops.GEN_Q_SPEC)
    torch.cuda.synchronize()
    sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(sA):
        torch.cuda._sleep(int(3e9))                               # ~1.5 s of spinning in front of stream A's calls
        yA = ops.gen_fwd_f32(x, gp)
        mA = ops.sim_gen_metrics(cfg, 7 * 64 * 9, gparams=gp, seed=3)
    with torch.cuda.stream(sB):
        yB = ops.gen_fwd_f32(x, gp)
        mB = ops.sim_gen_metrics(cfg, 7 * 64 * 9, gparams=gp, seed=3)
        qB = ops.gen_fwd_q(xq, W, Bq, mode=ops.GEN_Q_SPEC)
    sB.synchronize()
    assert not sA.query(), "stream B's calls waited for stream A"
    assert torch.equal(yB, ref_y) and torch.equal(mB, ref_m) and torch.equal(qB, ref_q)
    sA.synchronize()
    assert torch.equal(yA, ref_y) and torch.equal(mA, ref_m)
    # two host threads, each on its own stream, at the same time
    import threading
    out = {}

    def work(i):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(20):
                y = ops.gen_fwd_f32(x, gp)
                m = ops.sim_gen_metrics(cfg, 7 * 64 * 9, gparams=gp, seed=3)
        st.synchronize()
        out[i] = (y, m)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert all(torch.equal(out[i][0], ref_y) and torch.equal(out[i][1], ref_m) for i in range(4))


@pytest.mark.parametrize("channel_type", ["rayleigh", "rician", "multipath"])
def test_run_benchmark_fading_channels_match_the_unfused_pieces(pkg, channel_type):
    """run_benchmark(channel_type=...) (benchmark_comparison.py:154): the fused sweep with equaliser rows AND a fading channel ==
    the same frames through ofdmgan_chan_sim -> generator / ofdmgan_equalize -> metrics"""
    from ofdm_gan_sr_b200.sweep import run_benchmark
    ops = pkg.ops
    G, D, TG, TD = _pair(pkg, 8)
    n = 1500
    res = run_benchmark(G, n_trials=n, nonlinear=True, pa_saturation=0.8, channel_type=channel_type, seed=6)
    assert set(res) == {"GAN", "ZF", "MMSE", "NoEQ"}
    cfg = ops.make_cfg(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=n,
                       channel_type=channel_type)
    clean, noisy, snr = ops.chan_sim(cfg, 7 * n, seed=6)

    def evm_rows(est):
        e = ((est - clean) ** 2).sum(dim=(1, 2)).double()
        r = (clean ** 2).sum(dim=(1, 2)).double()
        return (20 * torch.log10(torch.sqrt(e / r) + 1e-10)).view(7, n).mean(dim=1)

    with torch.no_grad():
        rows = {"GAN": evm_rows(TG(noisy)), "NoEQ": evm_rows(noisy), "MMSE": evm_rows(ops.equalize(noisy, clean, pkg._lib.METHOD_MMSE, snr_db=snr))}
    for m, want in rows.items():
        for i, s in enumerate(res[m]):
            assert abs(res[m][s]["evm"] - float(want[i])) <= 5e-5 * max(1.0, abs(float(want[i]))), (m, s)
    with pytest.raises(ValueError):
        run_benchmark(G, n_trials=10, channel_type="no-such-channel")


@pytest.mark.parametrize("gen_kind", [1, 2])
def test_integer_generators_inside_the_benchmark_sweep(pkg, gen_kind, monkeypatch):
    """integer generator x equaliser rows x fading channel in ONE fused launch: the generator row equals the separate integer
    entry point on the same Q8.8 frames, the equaliser rows equal the fp32 generator's sweep (they do not depend on the generator)"""
    import oracle
    ops = pkg.ops
    monkeypatch.setenv("OFDMGAN_SIM_IMPL", "general")              # the stand-alone simulator call on the same kernel as the fused one
    rng = np.random.default_rng(gen_kind)
    W = rng.integers(-128, 128, 2048).astype(np.int8)
    Bq = np.zeros(64, np.int16)
    Bq[:18] = rng.integers(-300, 300, 18)
    gp = (rng.standard_normal(258) * 0.3).astype(np.float32)
    for kw in (dict(equalizers=True), dict(channel_type="rayleigh"), dict(equalizers=True, channel_type="multipath")):
        cfg = ops.make_cfg(nonlinear=True, pa_saturation=0.8, normalize=2, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7, frames_per_snr=128, **kw)
        B = 7 * 128 * 5
        fused = ops.sim_gen_metrics(cfg, B, gen_kind=gen_kind, wrom=W, brom=Bq, seed=3).cpu().numpy()
        f32 = ops.sim_gen_metrics(cfg, B, gen_kind=ops.GEN_F32, gparams=gp, seed=3).cpu().numpy()
        assert np.array_equal(fused[:, :, 0], f32[:, :, 0])
        np.testing.assert_allclose(fused[:, 1:, 1:], f32[:, 1:, 1:], rtol=1e-12, atol=0)     # NoEQ / ZF / MMSE rows: same frames
        clean, noisy, _ = ops.chan_sim(cfg, B, seed=3)
        yq = ops.gen_fwd_q(ops.quantize_q88(noisy), W, Bq, mode=gen_kind)
        bins = torch.as_tensor((np.arange(B) // 128 % 7).astype(np.int32)).cuda()
        m = ops.frame_metrics(ops.dequantize_q88(yq), clean, bins, method=0, n_snr=7).cpu().numpy()
        assert np.array_equal(m[:, 0, 0], fused[:, 0, 0])
        for c in (1, 3):
            np.testing.assert_allclose(fused[:, 0, c], m[:, 0, c], rtol=1e-6)


def test_module_path_reflattens_only_when_parameters_change(pkg):
    """the flat parameter vector of the fused modules is cached on the parameters' version counters: an optimiser step (in-place
    write) invalidates it, repeated forwards reuse it"""
    G, D, TG, TD = _pair(pkg, 11)
    x = torch.randn(64, 2, 16, device="cuda")
    y0 = G(x)
    f0 = pkg.ops.flat_cached(list(G.parameters()))
    assert pkg.ops.flat_cached(list(G.parameters())) is f0                   # unchanged parameters: the same tensor, no new cat
    opt = torch.optim.SGD(G.parameters(), lr=0.1)
    G(x).sum().backward()
    opt.step()
    f1 = pkg.ops.flat_cached(list(G.parameters()))
    assert f1 is not f0 and not torch.equal(f1, f0)
    with torch.no_grad():
        TG.load_state_dict(G.state_dict())
        assert float((G(x) - TG(x)).abs().max()) < 1e-5 and float((G(x) - y0).abs().max()) > 1e-6

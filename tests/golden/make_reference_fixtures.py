#!/usr/bin/env python3
"""Record golden input/output fixtures from the UNMODIFIED reference, imported from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_reference_fixtures.py

Writes (all small, committed):
  tests/golden/ref_fp32.npz      MiniGenerator / MiniDiscriminator forward, autograd gradients, gradient penalty,
                                 train_discriminator / train_generator loss+grad, torch.optim.Adam trajectories
  tests/golden/ref_channel.npz   SyntheticOFDMDataset samples (AWGN, non-linear) with the np.random draws that
                                 produced them, benchmark_comparison frames + per-trial MSE/EVM, QAMModulator /
                                 OFDMModulator frames and hard decisions, quantize_tensor results
  tests/golden/verification_golden.npz  copy of verification_output/golden_vectors/*.npy (+ parsed .hex)

Everything that uses randomness records the draws so the CUDA path and the oracle can be fed identical inputs
(draw order randn Re, randn Im, [randn phase], uniform snr, randn noise Re, Im).
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
# benchmark_comparison imports matplotlib, which this image does not have
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))

import torch  # noqa: E402

torch.set_num_threads(1)

from models import MiniGenerator, MiniDiscriminator  # noqa: E402
from models.discriminator import compute_gradient_penalty  # noqa: E402
from utils.dataset import SyntheticOFDMDataset  # noqa: E402
from utils.ofdm_utils import QAMModulator, OFDMModulator  # noqa: E402
from utils import quantization as refq  # noqa: E402
import benchmark_comparison as bc  # noqa: E402


def flat(params):
    return torch.cat([p.detach().reshape(-1) for p in params]).numpy().copy()


def flat_grads(params):
    # a parameter autograd never reached (e.g. biases under the gradient penalty) has grad None == zero
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().reshape(-1)
                      for p in params]).numpy().copy()


def randomise(model, gen, scale_b=0.1, scale_w=1.0):
    """Xavier weights keep their init; biases (zero at init) get small random values so bias paths are tested."""
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias"):
                p.copy_(scale_b * torch.randn(p.shape, generator=gen))
            else:
                p.mul_(scale_w)


class DrawRecorder:
    """Wraps np.random.randn / uniform, recording every draw in call order."""

    def __init__(self):
        self.log = []
        self._randn, self._uniform = np.random.randn, np.random.uniform

    def __enter__(self):
        def randn(*shape):
            v = self._randn(*shape)
            self.log.append(("randn", np.array(v, dtype=np.float64).reshape(-1)))
            return v

        def uniform(lo=0.0, hi=1.0, size=None):
            v = self._uniform(lo, hi, size)
            self.log.append(("uniform", np.array(v, dtype=np.float64).reshape(-1)))
            return v

        np.random.randn, np.random.uniform = randn, uniform
        return self

    def __exit__(self, *a):
        np.random.randn, np.random.uniform = self._randn, self._uniform


def fp32_fixtures():
    out = {}
    gen = torch.Generator().manual_seed(1234)
    torch.manual_seed(0)
    G, D = MiniGenerator(), MiniDiscriminator()
    randomise(G, gen)
    randomise(D, gen)
    out["gparams"], out["dparams"] = flat(G.parameters()), flat(D.parameters())
    assert out["gparams"].size == 258 and out["dparams"].size == 521

    B = 64
    x = (torch.rand(B, 2, 16, generator=gen) * 2 - 1)
    cond = (torch.rand(B, 2, 16, generator=gen) * 2 - 1)
    out["x"], out["cond"] = x.numpy(), cond.numpy()
    with torch.no_grad():
        out["g_y"] = G(x).numpy()
        out["d_score"] = D(x, cond).reshape(-1).numpy()

    # generator backward: dy -> dx, dparams
    xg = x.clone().requires_grad_(True)
    dy = torch.randn(B, 2, 16, generator=gen)
    G.zero_grad()
    G(xg).backward(dy)
    out["g_dy"], out["g_dx"], out["g_dparams"] = dy.numpy(), xg.grad.numpy().copy(), flat_grads(G.parameters())

    # critic backward: upstream g -> dcand, dcond, dparams
    ca, co = x.clone().requires_grad_(True), cond.clone().requires_grad_(True)
    gup = torch.randn(B, generator=gen)
    D.zero_grad()
    (D(ca, co).reshape(-1) * gup).sum().backward()
    out["d_gup"], out["d_dcand"], out["d_dcond"] = gup.numpy(), ca.grad.numpy().copy(), co.grad.numpy().copy()
    out["d_dparams"] = flat_grads(D.parameters())

    # gradient penalty alone (models/discriminator.py:172-236): alpha = the torch.rand(B,1,1) draw
    real = (torch.rand(B, 2, 16, generator=gen) * 2 - 1)
    fake = G(cond).detach()
    torch.manual_seed(77)
    alpha = torch.rand(B, 1, 1)
    torch.manual_seed(77)
    D.zero_grad()
    gp = compute_gradient_penalty(D, real, fake, cond)
    gp.backward()
    out["gp_real"], out["gp_fake"], out["gp_alpha"] = real.numpy(), fake.numpy(), alpha.reshape(-1).numpy()
    out["gp_value"], out["gp_dparams"] = np.float32(gp.item()), flat_grads(D.parameters())

    # train_discriminator loss + backward (train.py:220-250), gp_weight 10; no optimizer here
    torch.manual_seed(78)
    alpha2 = torch.rand(B, 1, 1)
    torch.manual_seed(78)
    D.zero_grad()
    with torch.no_grad():
        fake_signal = G(cond)
    d_real, d_fake = D(real, cond), D(fake_signal, cond)
    w_loss = d_fake.mean() - d_real.mean()
    gp2 = compute_gradient_penalty(D, real, fake_signal, cond, None)
    d_loss = w_loss + 10.0 * gp2
    d_loss.backward()
    out["cs_alpha"], out["cs_fake"] = alpha2.reshape(-1).numpy(), fake_signal.numpy()
    out["cs_grads"] = flat_grads(D.parameters())
    out["cs_stats"] = np.array([d_loss.item(), -w_loss.item(), gp2.item(), d_real.mean().item(), d_fake.mean().item()],
                               np.float32)

    # train_generator loss + backward (train.py:285-296), adv 1, rec 100
    G.zero_grad()
    D.zero_grad()
    fs = G(cond)
    adv = -D(fs, cond).mean()
    rec = torch.nn.functional.l1_loss(fs, real)
    g_loss = 1.0 * adv + 100.0 * rec
    g_loss.backward()
    out["gs_grads"] = flat_grads(G.parameters())
    out["gs_stats"] = np.array([g_loss.item(), adv.item(), rec.item()], np.float32)

    # Adam trajectories (train.py:114-127: lr 2e-4 betas (0,0.9); plus a beta1=0.5 case)
    for tag, betas in (("adam0", (0.0, 0.9)), ("adam5", (0.5, 0.999))):
        p = torch.nn.Parameter(torch.from_numpy(out["dparams"]).clone())
        opt = torch.optim.Adam([p], lr=2e-4, betas=betas)
        gs = []
        for step in range(4):
            g = torch.randn(521, generator=gen) * (10.0 ** (step - 2))
            gs.append(g.numpy())
            p.grad = g.clone()
            opt.step()
        st = opt.state[p]
        out[tag + "_g"] = np.stack(gs)
        out[tag + "_p"], out[tag + "_m"], out[tag + "_v"] = p.detach().numpy().copy(), st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()

    # a 3-iteration mini training run through the reference trainer maths (5 critic + 1 G per iteration, B=32)
    torch.manual_seed(5)
    G2, D2 = MiniGenerator(), MiniDiscriminator()
    out["tr_g0"], out["tr_d0"] = flat(G2.parameters()), flat(D2.parameters())
    oG = torch.optim.Adam(G2.parameters(), lr=2e-4, betas=(0.0, 0.9))
    oD = torch.optim.Adam(D2.parameters(), lr=2e-4, betas=(0.0, 0.9))
    Bt = 32
    clean_t = (torch.rand(3, Bt, 2, 16, generator=gen) * 2 - 1)
    noisy_t = clean_t + 0.1 * torch.randn(3, Bt, 2, 16, generator=gen)
    alphas, dstats, gstats = [], [], []
    for it in range(3):
        real_b, noisy_b = clean_t[it], noisy_t[it]
        for c in range(5):
            oD.zero_grad()
            with torch.no_grad():
                fk = G2(noisy_b)
            dr, df = D2(real_b, noisy_b), D2(fk, noisy_b)
            wl = df.mean() - dr.mean()
            st = torch.get_rng_state()
            a = torch.rand(Bt, 1, 1)
            torch.set_rng_state(st)
            gpv = compute_gradient_penalty(D2, real_b, fk, noisy_b, None)
            dl = wl + 10.0 * gpv
            dl.backward()
            oD.step()
            alphas.append(a.reshape(-1).numpy())
            dstats.append([dl.item(), -wl.item(), gpv.item(), dr.mean().item(), df.mean().item()])
        oG.zero_grad()
        fk = G2(noisy_b)
        advl = -D2(fk, noisy_b).mean()
        recl = torch.nn.functional.l1_loss(fk, real_b)
        gl = advl + 100.0 * recl
        gl.backward()
        oG.step()
        gstats.append([gl.item(), advl.item(), recl.item()])
    out["tr_clean"], out["tr_noisy"] = clean_t.numpy(), noisy_t.numpy()
    out["tr_alpha"] = np.stack(alphas).reshape(3, 5, Bt)
    out["tr_dstats"], out["tr_gstats"] = np.array(dstats, np.float32).reshape(3, 5, 5), np.array(gstats, np.float32)
    out["tr_g3"], out["tr_d3"] = flat(G2.parameters()), flat(D2.parameters())
    np.savez_compressed(os.path.join(HERE, "ref_fp32.npz"), **out)
    print("ref_fp32.npz:", {k: v.shape for k, v in out.items()})
    return G, out


def dataset_samples(n, **kw):
    ds = SyntheticOFDMDataset(n_samples=n, **kw)
    nonlinear = kw.get("nonlinear", False)
    sym, pn, snr, noise, clean, noisy = [], [], [], [], [], []
    for i in range(n):
        with DrawRecorder() as rec:
            s = ds[i]
        kinds = [k for k, _ in rec.log]
        assert kinds == (["randn"] * (3 if nonlinear else 2) + ["uniform"] + ["randn"] * 2), kinds
        vals = [v for _, v in rec.log]
        sym.append(np.concatenate(vals[0:2]))
        j = 2
        if nonlinear:
            pn.append(vals[2]); j = 3
        else:
            pn.append(np.zeros(16))
        snr.append(vals[j][0])
        noise.append(np.concatenate(vals[j + 1:j + 3]))
        clean.append(s["clean"].numpy()); noisy.append(s["noisy"].numpy())
        assert abs(float(s["snr"]) - np.float32(vals[j][0])) < 1e-6
    return dict(sym=np.array(sym), pn=np.array(pn), snr=np.array(snr), noise=np.array(noise),
                clean=np.array(clean, np.float32), noisy=np.array(noisy, np.float32))


def channel_fixtures(G):
    out = {}
    np.random.seed(0)
    for tag, kw in (("awgn10", dict(snr_range=(10, 10))),
                    ("awgn", dict(snr_range=(0, 30))),
                    ("nl08", dict(snr_range=(0, 30), nonlinear=True, pa_saturation=0.8)),
                    ("nl10", dict(snr_range=(5, 20), nonlinear=True, pa_saturation=1.0, iq_imbalance_db=0.5,
                                  iq_phase_deg=-3.0, phase_noise_dbchz=-85))):
        d = dataset_samples(64, **kw)
        for k, v in d.items():
            out[f"{tag}_{k}"] = v

    # benchmark_comparison inner loop (benchmark_comparison.py:184-214), separate normalisation, fixed SNR grid
    np.random.seed(1)
    G.eval()
    for tag, nonlinear in (("bm_lin", False), ("bm_nl", True)):
        rows = []
        sym, pn, noise, snrs, clean, noisy, gan, mets = [], [], [], [], [], [], [], []
        for snr in (0, 5, 10, 15, 20, 25, 30):
            for trial in range(6):
                with DrawRecorder() as rec:
                    cc = bc.generate_test_signal(16, "ofdm")
                    nc, _ = bc.apply_channel_and_impairments(cc, snr, "awgn", nonlinear, 0.8 if nonlinear else 1.0)
                vals = [v for _, v in rec.log]
                sym.append(np.concatenate(vals[0:2]))
                j = 2
                if nonlinear:
                    pn.append(vals[2]); j = 3
                else:
                    pn.append(np.zeros(16))
                noise.append(np.concatenate(vals[j:j + 2]))
                snrs.append(float(snr))
                ci, ni = bc.complex_to_iq(cc), bc.complex_to_iq(nc)
                nn_, _ = bc.normalize_iq(ni)
                cn_, _ = bc.normalize_iq(ci)
                with torch.no_grad():
                    go = G(torch.from_numpy(nn_).unsqueeze(0).float()).squeeze(0).numpy()
                clean.append(cn_); noisy.append(nn_); gan.append(go)
                mets.append([bc.compute_mse(go, cn_), bc.compute_evm(go, cn_), bc.compute_mse(nn_, cn_), bc.compute_evm(nn_, cn_)])
        out[tag + "_sym"], out[tag + "_pn"], out[tag + "_noise"] = np.array(sym), np.array(pn), np.array(noise)
        out[tag + "_snr"] = np.array(snrs)
        out[tag + "_clean"], out[tag + "_noisy"] = np.array(clean, np.float32), np.array(noisy, np.float32)
        out[tag + "_gan"], out[tag + "_metrics"] = np.array(gan, np.float32), np.array(mets, np.float64)

    # QAMModulator('QPSK') + OFDMModulator: stream truncated to 16 samples (ImageOFDMConverter, ofdm_utils.py:923-929)
    rng = np.random.RandomState(3)
    qam = QAMModulator("QPSK")
    for tag, (N, cp, sp) in (("q16", (16, 0, 8)), ("q8", (8, 2, 4)), ("q16cp", (16, 2, 16))):
        ofdm = OFDMModulator(n_subcarriers=N, cp_length=cp, pilot_spacing=sp, pilot_value=1 + 0j)
        words, frames, syms_dec = [], [], []
        for i in range(32):
            w = int(rng.randint(0, 2 ** 32, dtype=np.uint64))
            bits = np.array([(w >> (31 - b)) & 1 for b in range(32)])
            nsym_stream = int(np.ceil(16 / (N + cp)))
            n_data = ofdm.n_data_subcarriers * nsym_stream
            data_bits = np.concatenate([bits, np.zeros(max(0, 2 * n_data - 32), int)])[:2 * n_data]
            sig = ofdm.modulate(qam.modulate(data_bits))
            fr = np.zeros(16, complex)
            fr[:min(16, len(sig))] = sig[:16]
            frames.append(np.stack([fr.real, fr.imag]))
            # hard decisions on the complete symbols contained in the frame
            ncomplete = 16 // (N + cp)
            d, _ = ofdm.demodulate(fr[:ncomplete * (N + cp)])
            dec = qam.demodulate(d)
            ref_bits = data_bits[:len(dec)]
            assert np.array_equal(dec, ref_bits), (tag, i)
            words.append(w)
        out[tag + "_words"] = np.array(words, np.uint32)
        out[tag + "_frames"] = np.array(frames, np.float64)
    # QPSK demodulate decisions on noisy symbols, including exact ties (Re==0 / Im==0)
    s = (rng.randn(256) + 1j * rng.randn(256)) * 0.7
    s[:8] = [0, 1j, -1j, 1, -1, 0.5, -0.5j, 0 + 0j]
    out["qpsk_syms"] = np.stack([s.real, s.imag])
    out["qpsk_bits"] = qam.demodulate(s).astype(np.int8)
    out["qpsk_table"] = np.stack([qam.constellation.real, qam.constellation.imag])

    # utils/quantization.py: compute_scale / quantize_tensor / dequantize_tensor
    t = torch.from_numpy(np.concatenate([np.linspace(-1.2, 1.2, 97), [0.5 / 128, 1.5 / 128, 2.5 / 128, -0.5 / 128, -1.5 / 128]]).astype(np.float32))
    out["qt_in"] = t.numpy()
    out["qt_q17"] = refq.quantize_tensor(t, torch.tensor(1.0 / 128), 8).numpy()
    sc = refq.compute_scale(t, 8)
    out["qt_scale"] = np.float32(sc.item())
    out["qt_q8"] = refq.quantize_tensor(t, sc, 8).numpy()
    out["qt_deq"] = refq.dequantize_tensor(refq.quantize_tensor(t, sc, 8), sc).numpy()
    w = torch.from_numpy(np.random.RandomState(4).randn(4, 2, 3).astype(np.float32))
    scc = refq.compute_scale(w, 8, per_channel=True, channel_dim=0)
    out["qt_w"], out["qt_w_scale"], out["qt_w_q"] = w.numpy(), scc.numpy(), refq.quantize_tensor(w, scc, 8).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_channel.npz"), **out)
    print("ref_channel.npz:", {k: v.shape for k, v in out.items()})


def verification_golden():
    d = os.path.join(REF, "verification_output", "golden_vectors")
    out = {n: np.load(os.path.join(d, n + ".npy")) for n in ("input_float", "input_q88", "output_float", "output_q88")}
    for n in ("input", "output"):
        words = [int(l.strip(), 16) for l in open(os.path.join(d, n + ".hex")) if l.strip()]
        out[n + "_hex"] = np.array(words, np.uint16)
    np.savez_compressed(os.path.join(HERE, "verification_golden.npz"), **out)
    print("verification_golden.npz:", {k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    G, _ = fp32_fixtures()
    channel_fixtures(G)
    verification_golden()

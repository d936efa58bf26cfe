"""GPU parity, training side: critic forward/backward, gradient penalty + closed-form double backward, the fused
critic / generator steps and Adam - libofdmgan through the C ABI vs fixtures recorded from the reference's autograd
path and vs the CPU oracle.  <= 1e-5 relative for fp32 outputs and gradients (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

import oracle
from conftest import assert_close

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    import ofdm_gan_sr_b200 as pkg
    assert torch.cuda.is_available()
    return pkg.ops


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def _batch(B, seed):
    rng = np.random.default_rng(seed)
    clean = rng.uniform(-1, 1, (B, 2, 16)).astype(np.float32)
    noisy = (clean + 0.3 * rng.standard_normal((B, 2, 16))).astype(np.float32)
    alpha = rng.uniform(0, 1, B).astype(np.float32)
    return clean, noisy, alpha


# ------------------------------------------------------------------------------------------------ reference fixtures
def test_generator_backward_matches_reference(ops, ref_fp32):
    r = ref_fp32
    dx, dparams = ops.gen_bwd_f32(cu(r["x"]), r["gparams"], cu(r["g_dy"]))
    assert_close(host(dx), r["g_dx"], TOL, "G dx")
    assert_close(host(dparams), r["g_dparams"], TOL, "G dparams")
    _, dparams2 = ops.gen_bwd_f32(cu(r["x"]), r["gparams"], cu(r["g_dy"]), need_dx=False)
    assert torch.equal(dparams, dparams2)


def test_critic_forward_backward_matches_reference(ops, ref_fp32):
    r = ref_fp32
    assert_close(host(ops.disc_fwd_f32(cu(r["x"]), cu(r["cond"]), r["dparams"])), r["d_score"], TOL, "D forward")
    dcand, dcond, grads = ops.disc_bwd_f32(cu(r["x"]), cu(r["cond"]), cu(r["dparams"]), cu(r["d_gup"]))
    assert_close(host(dcand), r["d_dcand"], TOL, "D dcand")
    assert_close(host(dcond), r["d_dcond"], TOL, "D dcond")
    assert_close(host(grads), r["d_dparams"], TOL, "D dparams")
    _, _, grads2 = ops.disc_bwd_f32(cu(r["x"]), cu(r["cond"]), cu(r["dparams"]), cu(r["d_gup"]), need_dcand=False,
                                    need_dcond=False)
    assert_close(host(grads2), r["d_dparams"], TOL, "D dparams (no input grads)")


def test_gradient_penalty_matches_reference(ops, ref_fp32):
    r = ref_fp32
    gp, grads = ops.gradient_penalty(cu(r["gp_real"]), cu(r["gp_fake"]), cu(r["cond"]), r["dparams"], alpha=cu(r["gp_alpha"]))
    assert abs(float(gp) - float(r["gp_value"])) <= TOL * abs(float(r["gp_value"]))
    g = host(grads)
    assert_close(g, r["gp_dparams"], TOL, "GP dparams")
    for lo, hi in ((96, 104), (488, 504), (520, 521)):          # zero gradient for every bias
        assert np.all(g[lo:hi] == 0)


def test_critic_and_generator_step_match_reference(ops, ref_fp32):
    r = ref_fp32
    out = host(ops.critic_step(cu(r["gp_real"]), cu(r["cond"]), cu(r["cs_fake"]), r["dparams"], alpha=cu(r["cs_alpha"]),
                               gp_weight=10.0))
    assert_close(out[:521], r["cs_grads"], TOL, "critic grads")
    assert_close(out[521:526], r["cs_stats"], TOL, "critic stats")
    fake = torch.empty(64, 2, 16, device="cuda")
    out = host(ops.gen_step(cu(r["gp_real"]), cu(r["cond"]), r["dparams"], r["gparams"], 1.0, 100.0, fake_out=fake))
    assert_close(out[:258], r["gs_grads"], TOL, "generator grads")
    assert_close(out[258:261], r["gs_stats"], TOL, "generator stats")
    assert_close(host(fake), r["cs_fake"], TOL, "fake")


def test_adam_matches_reference_and_oracle_bitwise(ops, ref_fp32):
    r = ref_fp32
    for tag, (b1, b2) in (("adam0", (0.0, 0.9)), ("adam5", (0.5, 0.999))):
        p, m, v = cu(r["dparams"].copy()), torch.zeros(521, device="cuda"), torch.zeros(521, device="cuda")
        po, mo, vo = r["dparams"].copy(), np.zeros(521, np.float32), np.zeros(521, np.float32)
        for step in range(4):
            ops.adam(p, m, v, cu(r[tag + "_g"][step]), 2e-4, b1, b2, 1e-8, step + 1)
            po, mo, vo = oracle.adam(po, mo, vo, r[tag + "_g"][step], 2e-4, b1, b2, 1e-8, step + 1)
        assert_close(host(p), r[tag + "_p"], 1e-6, tag + " p")
        assert_close(host(m), r[tag + "_m"], 1e-6, tag + " m")
        assert_close(host(v), r[tag + "_v"], 1e-6, tag + " v")
        assert np.array_equal(host(p), po) and np.array_equal(host(m), mo) and np.array_equal(host(v), vo)


def test_three_training_iterations_match_reference(ops, ref_fp32):
    """train.py:327-344 replayed with the reference's torch.rand alphas: 15 critic + 3 generator optimizer steps."""
    r = ref_fp32
    g, d = cu(r["tr_g0"].copy()), cu(r["tr_d0"].copy())
    gm, gv, dm, dv = (torch.zeros(n, device="cuda") for n in (258, 258, 521, 521))
    dstep = gstep = 0
    for it in range(3):
        clean, noisy = cu(r["tr_clean"][it]), cu(r["tr_noisy"][it])
        fake = ops.gen_fwd_f32(noisy, g)                       # once per batch: G does not change inside the critic loop
        for c in range(5):
            out = ops.critic_step(clean, noisy, fake, d, alpha=cu(r["tr_alpha"][it, c]), gp_weight=10.0)
            assert_close(host(out[521:526]), r["tr_dstats"][it, c], 2e-5, f"dstats {it},{c}")
            dstep += 1
            ops.adam(d, dm, dv, out, 2e-4, 0.0, 0.9, 1e-8, dstep)
        out = ops.gen_step(clean, noisy, d, g, 1.0, 100.0)
        assert_close(host(out[258:261]), r["tr_gstats"][it], 2e-5, f"gstats {it}")
        gstep += 1
        ops.adam(g, gm, gv, out, 2e-4, 0.0, 0.9, 1e-8, gstep)
    assert_close(host(d), r["tr_d3"], TOL, "D params after 3 iterations")
    assert_close(host(g), r["tr_g3"], TOL, "G params after 3 iterations")


# ------------------------------------------------------------------------------------------------ oracle, larger / ragged
@pytest.mark.parametrize("B", [1, 33, 127, 129, 4096, 10007])
def test_critic_step_vs_oracle(ops, ref_fp32, B):
    clean, noisy, alpha = _batch(B, B)
    dp = (ref_fp32["dparams"] * 1.3).astype(np.float32)
    fake = oracle.gen_fwd_f32(noisy, ref_fp32["gparams"])
    out = host(ops.critic_step(cu(clean), cu(noisy), cu(fake), dp, alpha=cu(alpha), gp_weight=10.0))
    grads, stats = oracle.critic_step(clean, noisy, fake, alpha, dp, 10.0)
    assert_close(out[:521], grads, TOL, f"critic grads B={B}")
    assert_close(out[521:526], stats, TOL, f"critic stats B={B}")
    assert np.all(out[526:] == 0)


@pytest.mark.parametrize("B", [1, 33, 129, 4096, 10007])
def test_generator_step_vs_oracle(ops, ref_fp32, B):
    clean, noisy, _ = _batch(B, 100 + B)
    dp, gp = ref_fp32["dparams"], ref_fp32["gparams"]
    out = host(ops.gen_step(cu(clean), cu(noisy), dp, gp, 1.0, 100.0))
    grads, stats, _ = oracle.gen_step(clean, noisy, dp, gp, 1.0, 100.0)
    assert_close(out[:258], grads, TOL, f"generator grads B={B}")
    assert_close(out[258:261], stats, TOL, f"generator stats B={B}")


@pytest.mark.parametrize("B", [1, 33, 129, 4096, 10007])
def test_generator_step_from_a_precomputed_fake_vs_oracle(ops, ref_fp32, B):
    """ofdmgan_gen_step_fake: the step fed with G(noisy) from ofdmgan_gen_fwd_f32 (what a training iteration has at hand) is the
    same step - against the oracle, and against the entry point that computes the forward itself."""
    clean, noisy, _ = _batch(B, 300 + B)
    dp, gp = ref_fp32["dparams"], ref_fp32["gparams"]
    fake = ops.gen_fwd_f32(cu(noisy), gp)
    out = host(ops.gen_step(cu(clean), cu(noisy), dp, gp, 1.0, 100.0, fake=fake))
    grads, stats, _ = oracle.gen_step(clean, noisy, dp, gp, 1.0, 100.0)
    assert_close(out[:258], grads, TOL, f"generator grads from fake B={B}")
    assert_close(out[258:261], stats, TOL, f"generator stats from fake B={B}")
    assert np.all(out[261:] == 0)
    own = host(ops.gen_step(cu(clean), cu(noisy), dp, gp, 1.0, 100.0))
    assert_close(out[:261], own[:261], 2e-6, f"from fake vs own forward B={B}")


def test_fused_generator_update_equals_step_plus_adam(ops, ref_fp32):
    """ofdmgan_gen_train_ctr (step + one tail launch: reduction, Adam) == ofdmgan_gen_step_fake followed by ofdmgan_adam_ctr, bit for
    bit; and with the weight images taken from the staging buffers (as inside a training iteration) instead of rebuilt."""
    B = 4097
    clean, noisy, _ = _batch(B, 77)
    c, n = cu(clean), cu(noisy)
    d = cu(ref_fp32["dparams"]).clone()
    g0 = cu(ref_fp32["gparams"]).clone()
    hp = dict(lr=2e-4, beta1=0.0, beta2=0.9, eps=1e-8)
    z = lambda k: torch.zeros(k, dtype=torch.float32, device="cuda")

    def separate():
        p, m, v, ctr = g0.clone(), z(258), z(258), torch.zeros(1, dtype=torch.int32, device="cuda")
        fake = ops.gen_fwd_f32(n, p)
        out = ops.gen_step(c, n, d, p, 1.0, 100.0, fake=fake)
        ops.adam(p, m, v, out, hp["lr"], hp["beta1"], hp["beta2"], hp["eps"], 0, step_dev=ctr)
        return out.clone(), p, m, v, ctr

    def fused(staged):
        p, m, v, ctr = g0.clone(), z(258), z(258), torch.zeros(1, dtype=torch.int32, device="cuda")
        if staged:                                           # a critic update first: its tail leaves the image of `dd` staged
            dd, dm, dv, dctr = d.clone(), z(521), z(521), torch.zeros(1, dtype=torch.int32, device="cuda")
            fk = ops.gen_fwd_f32(n, p)
            ops.critic_train(c, n, fk, dd, dm, dv, dctr, 2e-4, 0.0, 0.9, 1e-8)
        else:
            dd = d
        fake = ops.gen_fwd_f32(n, p)                          # (stages the generator image of p)
        out = ops.gen_train(c, n, fake, p, m, v, ctr, hp["lr"], hp["beta1"], hp["beta2"], hp["eps"], dd, 1.0, 100.0,
                            d_image_staged=staged, g_image_staged=staged)
        return out.clone(), p, m, v, ctr, dd

    o0, p0, m0, v0, c0 = separate()
    o1, p1, m1, v1, c1, _ = fused(False)
    assert torch.equal(o0, o1) and torch.equal(p0, p1) and torch.equal(m0, m1) and torch.equal(v0, v1)
    assert int(c0.item()) == 1 and int(c1.item()) == 1
    o2, p2, m2, v2, c2, dd = fused(True)                      # different critic (one update applied): compare with rebuilt images
    p, m, v, ctr = g0.clone(), z(258), z(258), torch.zeros(1, dtype=torch.int32, device="cuda")
    fake = ops.gen_fwd_f32(n, p)
    o3 = ops.gen_train(c, n, fake, p, m, v, ctr, hp["lr"], hp["beta1"], hp["beta2"], hp["eps"], dd, 1.0, 100.0)
    assert torch.equal(o2, o3) and torch.equal(p2, p) and torch.equal(m2, m) and torch.equal(v2, v)
    assert not torch.equal(o2, o1)                            # (the critic update did change the step)


def test_generator_step_from_fake_argument_errors(ops, ref_fp32):
    clean, noisy, _ = _batch(64, 1)
    dp, gp = ref_fp32["dparams"], ref_fp32["gparams"]
    fake = ops.gen_fwd_f32(cu(noisy), gp)
    with pytest.raises(Exception):
        ops.gen_step(cu(clean), cu(noisy), dp, gp, fake=fake[:32])
    with pytest.raises(Exception):
        ops.gen_step(cu(clean), cu(noisy), dp, gp, fake=fake, fake_out=torch.empty_like(fake))


@pytest.mark.parametrize("B", [1, 95, 4097])
def test_backward_entry_points_vs_oracle(ops, ref_fp32, B):
    clean, noisy, alpha = _batch(B, 200 + B)
    dp, gp = ref_fp32["dparams"], ref_fp32["gparams"]
    g = (alpha - 0.5).astype(np.float32)
    dcand, dcond, grads = ops.disc_bwd_f32(cu(clean), cu(noisy), dp, cu(g))
    oc, on, og = oracle.disc_bwd_f32(clean, noisy, dp, g)
    assert_close(host(dcand), oc, TOL, "dcand")
    assert_close(host(dcond), on, TOL, "dcond")
    assert_close(host(grads), og, TOL, "D grads")
    assert_close(host(ops.disc_fwd_f32(cu(clean), cu(noisy), dp)), oracle.disc_fwd_f32(clean, noisy, dp), TOL, "score")
    dy = (noisy - clean).astype(np.float32)
    dx, dparams = ops.gen_bwd_f32(cu(noisy), gp, cu(dy))
    ox, op = oracle.gen_bwd_f32(noisy, gp, dy)
    assert_close(host(dx), ox, TOL, "G dx")
    assert_close(host(dparams), op, TOL, "G dparams")
    fake = oracle.gen_fwd_f32(noisy, gp)
    gpv, gpg = ops.gradient_penalty(cu(clean), cu(fake), cu(noisy), dp, alpha=cu(alpha))
    ov, ogr, _ = oracle.gradient_penalty(clean, fake, noisy, alpha, dp)
    assert abs(float(gpv) - ov) <= TOL * abs(ov)
    assert_close(host(gpg), ogr, TOL, "GP grads")


def test_data_parallel_shards_add_up(ops, ref_fp32):
    """What the NCCL allreduce does: per-rank outputs scaled by 1/B_global sum to the single-GPU result."""
    B = 8192
    clean, noisy, alpha = _batch(B, 7)
    dp, gp = ref_fp32["dparams"], ref_fp32["gparams"]
    fake = ops.gen_fwd_f32(cu(noisy), gp)
    whole = host(ops.critic_step(cu(clean), cu(noisy), fake, dp, alpha=cu(alpha)))
    parts = np.zeros_like(whole)
    gwhole = host(ops.gen_step(cu(clean), cu(noisy), dp, gp))
    gparts = np.zeros_like(gwhole)
    for lo, hi in ((0, 3000), (3000, 3001), (3001, B)):
        parts += host(ops.critic_step(cu(clean[lo:hi]), cu(noisy[lo:hi]), fake[lo:hi].contiguous(), dp, alpha=cu(alpha[lo:hi]),
                                      b_global=B))
        gparts += host(ops.gen_step(cu(clean[lo:hi]), cu(noisy[lo:hi]), dp, gp, b_global=B))
    # stats[0] (d_loss) is linear in the partial sums too
    assert_close(parts[:526], whole[:526], 2e-6, "critic shards")
    assert_close(gparts[:261], gwhole[:261], 2e-6, "generator shards")


def test_philox_alpha_is_rank_independent(ops, ref_fp32):
    """alpha = Philox(seed, global sample index, iteration): the penalty of a shard equals the same rows of the whole."""
    B = 4096
    clean, noisy, _ = _batch(B, 9)
    dp, gp = ref_fp32["dparams"], ref_fp32["gparams"]
    fake = ops.gen_fwd_f32(cu(noisy), gp)
    whole = host(ops.critic_step(cu(clean), cu(noisy), fake, dp, seed=5, sample0=1000, alpha_iter=3))
    parts = np.zeros_like(whole)
    for lo, hi in ((0, 1024), (1024, B)):
        parts += host(ops.critic_step(cu(clean[lo:hi]), cu(noisy[lo:hi]), fake[lo:hi].contiguous(), dp, seed=5, sample0=1000 + lo,
                                      alpha_iter=3, b_global=B))
    assert_close(parts[:526], whole[:526], 2e-6, "philox alpha shards")
    # and it is the documented stream: u = (x0 >> 8) * 2^-24 of block (sample, iter, purpose 1)
    x = host(ops.philox_blocks(5, 1000, 3, 1, B)).view(np.uint32)
    alpha = ((x[:, 0] >> 8).astype(np.float64) / 16777216.0).astype(np.float32)
    ref = host(ops.critic_step(cu(clean), cu(noisy), fake, dp, alpha=cu(alpha)))
    assert np.array_equal(ref, whole)
    other = host(ops.critic_step(cu(clean), cu(noisy), fake, dp, seed=5, sample0=1000, alpha_iter=4))
    assert not np.array_equal(other[:521], whole[:521])


def test_steps_are_deterministic(ops, ref_fp32):
    B = 65536
    clean, noisy, alpha = _batch(B, 3)
    dp, gp = ref_fp32["dparams"], ref_fp32["gparams"]
    c, n, a = cu(clean), cu(noisy), cu(alpha)
    fake = ops.gen_fwd_f32(n, gp)
    o1 = ops.critic_step(c, n, fake, dp, alpha=a).clone()
    o2 = ops.critic_step(c, n, fake, dp, alpha=a)
    assert torch.equal(o1, o2)
    g1 = ops.gen_step(c, n, dp, gp).clone()
    g2 = ops.gen_step(c, n, dp, gp)
    assert torch.equal(g1, g2)
    assert torch.isfinite(o1).all() and torch.isfinite(g1).all()

"""Oracle pinning, fp32 models and training maths: the C restatement vs fixtures recorded from the imported
reference (tests/golden/make_reference_fixtures.py).  CPU only."""
import numpy as np

import oracle
from conftest import assert_close

TOL = 1e-5   # BASELINE.json north_star: <= 1e-5 relative for fp32 outputs and gradients


def test_generator_forward(ref_fp32):
    r = ref_fp32
    assert_close(oracle.gen_fwd_f32(r["x"], r["gparams"]), r["g_y"], TOL, "G forward")


def test_generator_backward(ref_fp32):
    r = ref_fp32
    dx, dparams = oracle.gen_bwd_f32(r["x"], r["gparams"], r["g_dy"])
    assert_close(dx, r["g_dx"], TOL, "G dx")
    assert_close(dparams, r["g_dparams"], TOL, "G dparams")


def test_critic_forward_backward(ref_fp32):
    r = ref_fp32
    assert_close(oracle.disc_fwd_f32(r["x"], r["cond"], r["dparams"]), r["d_score"], TOL, "D forward")
    dcand, dcond, grads = oracle.disc_bwd_f32(r["x"], r["cond"], r["dparams"], r["d_gup"])
    assert_close(dcand, r["d_dcand"], TOL, "D dcand")
    assert_close(dcond, r["d_dcond"], TOL, "D dcond")
    assert_close(grads, r["d_dparams"], TOL, "D dparams")


def test_gradient_penalty_and_double_backward(ref_fp32):
    r = ref_fp32
    gp, grads, norms = oracle.gradient_penalty(r["gp_real"], r["gp_fake"], r["cond"], r["gp_alpha"], r["dparams"])
    assert abs(gp - float(r["gp_value"])) <= TOL * abs(float(r["gp_value"]))
    assert_close(grads, r["gp_dparams"], TOL, "GP dparams")
    # the penalty has zero gradient w.r.t. every bias (D is piecewise linear): SURVEY 3.4
    for lo, hi in ((96, 104), (488, 504), (520, 521)):
        assert np.all(grads[lo:hi] == 0) and np.all(r["gp_dparams"][lo:hi] == 0)


def test_critic_step(ref_fp32):
    r = ref_fp32
    grads, stats = oracle.critic_step(r["gp_real"], r["cond"], r["cs_fake"], r["cs_alpha"], r["dparams"], 10.0)
    assert_close(grads, r["cs_grads"], TOL, "critic grads")
    assert_close(stats, r["cs_stats"], TOL, "critic stats")


def test_generator_step(ref_fp32):
    r = ref_fp32
    grads, stats, fake = oracle.gen_step(r["gp_real"], r["cond"], r["dparams"], r["gparams"], 1.0, 100.0)
    assert_close(grads, r["gs_grads"], TOL, "generator grads")
    assert_close(stats, r["gs_stats"], TOL, "generator stats")
    assert_close(fake, r["cs_fake"], TOL, "fake")


def test_adam(ref_fp32):
    r = ref_fp32
    for tag, (b1, b2) in (("adam0", (0.0, 0.9)), ("adam5", (0.5, 0.999))):
        p, m, v = r["dparams"].copy(), np.zeros(521, np.float32), np.zeros(521, np.float32)
        for step in range(4):
            p, m, v = oracle.adam(p, m, v, r[tag + "_g"][step], 2e-4, b1, b2, 1e-8, step + 1)
        assert_close(p, r[tag + "_p"], 1e-6, tag + " p")
        assert_close(m, r[tag + "_m"], 1e-6, tag + " m")
        assert_close(v, r[tag + "_v"], 1e-6, tag + " v")


def test_three_training_iterations(ref_fp32):
    """5 critic + 1 generator updates x 3 batches, replaying the reference's torch.rand alphas: parameters after
    18 optimizer steps must match the reference trainer maths (train.py:327-344)."""
    r = ref_fp32
    g, d = r["tr_g0"].copy(), r["tr_d0"].copy()
    gm, gv, dm, dv = (np.zeros(n, np.float32) for n in (258, 258, 521, 521))
    dstep = gstep = 0
    for it in range(3):
        clean, noisy = r["tr_clean"][it], r["tr_noisy"][it]
        for c in range(5):
            fake = oracle.gen_fwd_f32(noisy, g)
            grads, stats = oracle.critic_step(clean, noisy, fake, r["tr_alpha"][it, c], d, 10.0)
            assert_close(stats, r["tr_dstats"][it, c], 2e-5, f"dstats {it},{c}")
            dstep += 1
            d, dm, dv = oracle.adam(d, dm, dv, grads, 2e-4, 0.0, 0.9, 1e-8, dstep)
        grads, stats, _ = oracle.gen_step(clean, noisy, d, g, 1.0, 100.0)
        assert_close(stats, r["tr_gstats"][it], 2e-5, f"gstats {it}")
        gstep += 1
        g, gm, gv = oracle.adam(g, gm, gv, grads, 2e-4, 0.0, 0.9, 1e-8, gstep)
    assert_close(d, r["tr_d3"], 1e-5, "D params after 3 iterations")
    assert_close(g, r["tr_g3"], 1e-5, "G params after 3 iterations")

#!/usr/bin/env python
"""bench.py - the headline measurement of the ofdm-gan-sr hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Primary line (BASELINE.json metric "OFDM frames/s: fused channel sim + G inference"): one step = one pass of the fused
kernel (on-chip Philox channel simulation with the non-linear chain, SNR grid 0..30 dB, -> fp32 MiniGenerator -> EVM/MSE
accumulation) over FRAMES_PER_GPU frames on every rank, sharded by global frame index, plus the final metric allreduce.
`also` carries the other BASELINE configs measured in the same run: the Q1.7/Q8.8 integer generator over 2^24 HBM-resident
frames (config 2) and the CWGAN-GP training step at 65,536 frames per GPU (config 3).

`--impl reference` times the CPU restatement of the reference path (oracle/, OpenMP over all host cores) on a bounded
sample of the same workload: the reference is Python/NumPy/Verilog, nothing of it compiles into oracle/_ref.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_GPU = 1 << 24           # frames per rank per step of the primary workload (weak scaling)
Q_FRAMES = 1 << 24                 # BASELINE config 2
TRAIN_FRAMES_PER_GPU = 65536       # BASELINE config 3
FLOP_SIM = 900                     # nominal channel-sim FLOP per frame (Philox integer work not credited)
FLOP_GEN = 3456                    # 2 x 1728 MAC, models/generator.py:227-233
FLOP_TRAIN = 311424                # 5 x 57,504 + 23,904 per sample, SURVEY.md 8(d)
FLOP_TRAIN_EXECUTED = FLOP_TRAIN - 5 * FLOP_GEN   # the step runs ONE G(noisy) forward for its five critic updates and its generator update
Q_BYTES = 128                      # int16 frame in + out
WORKLOAD = dict(nonlinear=True, pa_saturation=0.8, normalize=1, snr_mode=1, snr_lo=0.0, snr_step=5.0, n_snr=7,
                frames_per_snr=1 << 10)
WORKLOAD_NAME = ("C4: Gaussian-symbol OFDM 2x16 frames, ifft*sqrt(N), Rapp PA (A=0.8,p=3) + IQ imbalance (1 dB, 5 deg) + "
                 "phase noise (-80 dBc/Hz) + AWGN on an SNR grid 0..30 dB step 5, joint normalisation, fp32 MiniGenerator "
                 "inference, per-SNR MSE/EVM accumulation; 2^24 frames per GPU per step, generated on-chip (Philox4x32-10)")


def seed_params():
    """Deterministic Xavier-like random-init weights of the two architectures (no checkpoints offline)."""
    rng = np.random.default_rng(0)
    gp = np.zeros(258, np.float32)
    for off, (oc, ic) in ((0, (4, 2)), (28, (8, 4)), (132, (4, 8)), (232, (2, 4))):
        a = np.sqrt(6.0 / (ic * 3 + oc * 3))
        gp[off:off + oc * ic * 3] = rng.uniform(-a, a, oc * ic * 3)
    dp = np.zeros(521, np.float32)
    for off, (oc, ic, k) in ((0, (8, 4, 3)), (104, (16, 8, 3)), (504, (1, 16, 1))):
        a = np.sqrt(6.0 / (ic * k + oc * k))
        dp[off:off + oc * ic * k] = rng.uniform(-a, a, oc * ic * k)
    return gp, dp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_uuid):
        self.rows, self.proc, self.uuid = [], None, gpu_uuid

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm),
                "reasons": sorted(reasons)}


def torch_eager_bar(gp, dp, clean, noisy, K):
    """The same box without custom kernels: MiniGenerator / MiniDiscriminator (models/generator.py:180-208,
    models/discriminator.py:112-152) and the 5+1 CWGAN-GP step of train.py:201-305 written with stock PyTorch CUDA ops
    (F.conv1d, autograd incl. the double backward of the penalty, torch.optim.Adam), fp32, TF32 off.  Returns timings and
    the largest difference between its generator output and libofdmgan's on the same frames."""
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True                        # let cuDNN pick its fastest algorithm for these shapes
    torch.backends.cuda.matmul.allow_tf32 = False

    def split(v, shapes):
        out, o = [], 0
        for sh in shapes:
            n = int(np.prod(sh))
            out.append(v[o:o + n].view(*sh).clone().requires_grad_(True))
            o += n
        return out

    G = split(gp, [(4, 2, 3), (4,), (8, 4, 3), (8,), (4, 8, 3), (4,), (2, 4, 3), (2,)])
    D = split(dp, [(8, 4, 3), (8,), (16, 8, 3), (16,), (1, 16), (1,)])

    up = [lambda t: F.interpolate(t, scale_factor=2, mode="nearest")]

    def gen(x):
        e1 = F.leaky_relu(F.conv1d(x, G[0], G[1], stride=2, padding=1), 0.2)
        b = F.leaky_relu(F.conv1d(e1, G[2], G[3], stride=2, padding=1), 0.2)
        d1 = F.leaky_relu(F.conv1d(up[0](b), G[4], G[5], padding=1), 0.2) + e1
        return torch.tanh(F.conv1d(up[0](d1), G[6], G[7], padding=1))

    def disc(c, cond):
        h = F.leaky_relu(F.conv1d(torch.cat([c, cond], dim=1), D[0], D[1], stride=2, padding=1), 0.2)
        h = F.leaky_relu(F.conv1d(h, D[2], D[3], stride=2, padding=1), 0.2)
        return F.linear(h.sum(dim=2), D[4], D[5])

    opt_g = torch.optim.Adam(G, lr=2e-4, betas=(0.0, 0.9))
    opt_d = torch.optim.Adam(D, lr=2e-4, betas=(0.0, 0.9))

    def step():
        for _ in range(5):
            with torch.no_grad():
                fake = gen(noisy)
            alpha = torch.rand(clean.shape[0], 1, 1, device=clean.device)
            inter = (alpha * clean + (1 - alpha) * fake).requires_grad_(True)
            grad = torch.autograd.grad(disc(inter, noisy).sum(), inter, create_graph=True)[0]
            gp_term = ((grad.reshape(grad.shape[0], -1).norm(2, dim=1) - 1) ** 2).mean()
            d_loss = disc(fake, noisy).mean() - disc(clean, noisy).mean() + 10.0 * gp_term
            opt_d.zero_grad(set_to_none=True)
            d_loss.backward()
            opt_d.step()
        fake = gen(noisy)
        g_loss = -disc(fake, noisy).mean() + 100.0 * F.l1_loss(fake, clean)
        opt_g.zero_grad(set_to_none=True)
        g_loss.backward()
        opt_g.step()

    with torch.no_grad():
        y = gen(noisy)
    import ofdm_gan_sr_b200 as pkg
    diff = float((y - pkg.ops.gen_fwd_f32(noisy, gp)).abs().max())

    def timed(fn, n):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def fwd():
        with torch.no_grad():
            gen(noisy)

    B = clean.shape[0]
    out = {"what": "stock PyTorch CUDA eager (F.conv1d + autograd + torch.optim.Adam), fp32, same B200, same batch; 'as_written' "
                   "upsamples with nn.Upsample(nearest) like models/generator.py:141,154 (9.8 ms per call at this batch, "
                   "tools/eager_probe.py), 'tuned' replaces it by repeat_interleave",
           "frames": B, "gen_fwd_max_abs_diff_vs_libofdmgan": diff}
    for tag, fn in (("as_written", up[0]), ("tuned", lambda t: t.repeat_interleave(2, dim=2))):
        up[0] = fn
        ms_fwd, ms_step = timed(fwd, K), timed(step, max(2, K // 2))
        out[tag] = {"gen_fwd_frames_per_s": B / (ms_fwd * 1e-3), "gen_fwd_ms": ms_fwd,
                    "train_samples_per_s": B / (ms_step * 1e-3), "train_ms_per_step": ms_step}
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)", d.get("sm_max_mhz", 1965.0)
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, saved_stdout):
    """CPU restatement of the reference path on the host cores, bounded sample per step."""
    if rank != 0:
        return
    import oracle
    gp, _ = seed_params()
    cfg = oracle.make_cfg(**WORKLOAD)
    cores = oracle.set_threads()                               # torchrun exports OMP_NUM_THREADS=1: take the whole host back
    t0 = time.perf_counter()
    oracle.sim_gen_metrics(cfg, 0, 1 << 17, gparams=gp, seed=1)
    rate = (1 << 17) / (time.perf_counter() - t0)
    budget = 90.0 / (args.steps + args.warmup)                 # whole run ~1.5 min
    sample = int(min(FRAMES_PER_GPU, max(1 << 16, 1 << int(np.log2(max(rate * budget, 1.0))))))
    for _ in range(args.warmup):
        oracle.sim_gen_metrics(cfg, 0, sample, gparams=gp, seed=1)
    t0 = time.perf_counter()
    for s in range(args.steps):
        oracle.sim_gen_metrics(cfg, 0, sample, gparams=gp, seed=1, frame0=s * sample)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": "OFDM frames/s: fused channel sim + G inference", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 channel / f32 generator",
            "data": "synthetic", "config": {"workload": WORKLOAD_NAME, "frames_per_step": sample},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} frames per step of the same workload (oracle/channel.c + fp32_models.c, OpenMP)"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    _emit(saved_stdout, line)


# ------------------------------------------------------------------------------------------------ B200 arm
def _quiet_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner to stdout) are sent to stderr."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(saved_fd, line):
    sys.stdout.flush()
    os.write(saved_fd, (json.dumps(line) + "\n").encode())


def main():
    saved_stdout = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=FRAMES_PER_GPU)
    ap.add_argument("--skip-also", action="store_true", help="primary workload only (used under ncu)")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, saved_stdout)
        return

    import torch
    import torch.distributed as dist
    import ofdm_gan_sr_b200 as pkg
    from ofdm_gan_sr_b200.train_step import CWGANGPStep
    ops = pkg.ops
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: libofdmgan has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    gp_h, dp_h = seed_params()
    gp_d = torch.as_tensor(gp_h).to(dev)
    cfg = ops.make_cfg(**WORKLOAD)
    F = args.frames_per_gpu
    K, W = args.steps, args.warmup
    hbm_peak, peak_src, sm_max = measured_peaks()

    # ---- primary: fused sim + G + metrics, device-resident weights and accumulator
    table = torch.zeros(7, pkg._lib.N_METHODS, pkg._lib.METRIC_COLS, dtype=torch.float64, device=dev)

    # A sweep is frame-sharded with ONE exchange at its end (BASELINE.json north_star: "SNR sweeps shard frames with only a final
    # BER/EVM reduce"): every step adds its 2^24 frames per rank into the rank's device-resident table, and the < 2 KB table is
    # all-reduced once after the last step - inside the timed region.
    def fused_step(s):
        ops.sim_gen_metrics(cfg, F, gparams=gp_d, seed=1, frame0=(s * world + rank) * F, out=table)

    def final_reduce():
        if world > 1:
            dist.all_reduce(table)                               # the sweep's only exchange

    for s in range(W):
        fused_step(s)
    final_reduce()
    barrier()
    uuid = str(getattr(torch.cuda.get_device_properties(dev), "uuid", "")) or str(local_rank)
    clocks = ClockSampler(uuid if uuid.startswith("GPU-") or uuid.isdigit() else "GPU-" + uuid)
    if rank == 0:
        clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 2)]
    table.zero_()
    barrier()
    ev[0].record()
    for s in range(K):
        fused_step(W + s)
        ev[s + 1].record()
    final_reduce()
    ev[K + 1].record()
    barrier()
    ms_total = max_over_ranks(ev[0].elapsed_time(ev[K + 1]))
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / K
    value = F * world / (ms_step * 1e-3)
    per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    n_frames_seen = float(table[:, :2, 0].sum().item())
    assert n_frames_seen == 2.0 * F * world * K, "metric table does not account for every frame"

    # ---- sustained: the same step back to back for >= 2 s, clocks and power sampled over the whole stretch (the K-step headline above
    # lasts tens of milliseconds; this shows what the rate does once the part has warmed up under load)
    sustained = None
    if not args.skip_also:
        n_sus = max(K, int(2.2e3 / ms_step) + 1)
        sclk = ClockSampler(clocks.uuid)
        barrier()
        if rank == 0:
            sclk.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for s_ in range(n_sus):
            fused_step(W + K + s_)
        final_reduce()
        s1.record()
        barrier()
        sus_ms = max_over_ranks(s0.elapsed_time(s1))
        sc = sclk.stop() if rank == 0 else None
        sustained = {"seconds": sus_ms * 1e-3, "steps": n_sus, "ms_per_step": sus_ms / n_sus, "value": F * world * n_sus / (sus_ms * 1e-3),
                     "unit": "frames/s", "vs_headline": (sus_ms / n_sus) / ms_step, "clocks": sc}

    # ---- multi-GPU parity, outside every timed region: the checks of tests/multi_gpu_checks.py (peer exchange bit-equal to the rank-
    # ordered sum, replicas identical, peer == NCCL == one GPU on the whole batch to 2e-5, sharded sweep == whole sweep).  A failure
    # ends the run with a non-zero exit code.
    parity = None
    if world > 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import multi_gpu_checks
        try:
            parity = multi_gpu_checks.run_all(dev, rank, world)
            parity["status"] = "ok"
        except BaseException as e:                               # noqa: BLE001 - any failure on any rank fails the whole launch
            sys.stderr.write(f"[rank {rank}] multi-GPU parity FAILED: {type(e).__name__}: {e}\n")
            sys.stderr.flush()
            os._exit(1)

    # ---- e2e: the host-buffer C-ABI call (weights from host memory in, metric table to host memory out), every step
    def e2e_step(s):
        out = ops.sim_gen_metrics_host(cfg, F, gparams=gp_h, seed=1, frame0=(s * world + rank) * F)
        if world > 1:
            t = torch.as_tensor(out).to(dev)
            dist.all_reduce(t)
            out = t.cpu().numpy()
        return out

    for s in range(2):
        e2e_step(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        e2e_step(2 + s)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = F * world * K / e2e_s
    h2d = 258 * 4
    d2h = 7 * pkg._lib.N_METHODS * pkg._lib.METRIC_COLS * 8

    # ---- FP32 issue-rate peak measured in the same run (denominator of the fp32 roofline)
    ffma = ops.ffma_peak(8192, device=dev)
    flop_frame = FLOP_SIM + FLOP_GEN
    achieved = F * flop_frame / (float(np.mean(per_step_ms)) * 1e-3) / 1e12
    # dram bytes per launch and issue-slot utilisation cannot be measured outside a profiler: they are PROFILE CONSTANTS, read from
    # the committed ncu capture of this kernel at this size and labelled as such (profiles/r2_kernel_counters.json)
    traffic = issue_frac = prof_src = None
    tpath = os.path.join(ROOT, "profiles", "r2_kernel_counters.json")
    if os.path.exists(tpath) and F == FRAMES_PER_GPU:
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        issue_frac = tj.get("issue_slots_busy_pct", 0.0) / 100.0 or None
        prof_src = tj.get("source")
    roofline = {"bound": "fp32", "achieved": achieved, "peak": ffma, "unit": "TFLOP/s", "frac": achieved / ffma, "traffic": traffic,
                "issue_frac": issue_frac, "traffic_and_issue_frac_source": prof_src,
                "kernel": "og::k_sim_lean<GEN_F32, CHAIN_NONLINEAR>", "algorithmic_flop_per_frame": flop_frame, "frames_per_launch": F,
                "peak_source": "FFMA issue-rate microbenchmark ofdmgan_ffma_peak, same run (148 SMs x 128 lanes x 2 flop x clock)",
                "note": "inputs are generated on-chip and outputs reduced on-chip: HBM traffic per launch is the per-CTA metric "
                        "partials only, so the binding roofline is the FP32/INT issue rate, not HBM (BASELINE.json north_star)",
                "hbm_peak_gbs": hbm_peak, "hbm_peak_source": peak_src, "sustained": sustained}

    also = {}
    launches = K * 3                                             # prep_g_image + k_sim_lean + k_reduce_partials per step
    if not args.skip_also:
        # ---- config 2: integer generator over 2^24 HBM-resident frames (1 GPU worth per rank)
        g = torch.Generator(device=dev).manual_seed(1 + rank)
        xq = ops.quantize_q88(torch.randn(Q_FRAMES, 2, 16, generator=g, device=dev).clamp_(-4, 4))
        rng = np.random.default_rng(5)
        Wrom = np.zeros(2048, np.int8)
        Wrom[:224] = np.clip(np.rint(rng.standard_normal(224) * 40), -128, 127)
        Brom = np.zeros(64, np.int16)
        Brom[:18] = rng.integers(-64, 64, 18)
        for mode, tag in ((ops.GEN_Q_SPEC, "spec"), (ops.GEN_Q_RTL, "rtl_literal")):
            for _ in range(3):
                yq = ops.gen_fwd_q(xq, Wrom, Brom, mode=mode)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                yq = ops.gen_fwd_q(xq, Wrom, Brom, mode=mode)      # 2 GiB of traffic per launch >> 126 MB L2
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1)) / K
            gbs = Q_FRAMES * Q_BYTES / (ms * 1e-3) / 1e9
            also["fixed_point_" + tag] = {"frames_per_s": Q_FRAMES * world / (ms * 1e-3), "ms_per_launch": ms, "frames": Q_FRAMES,
                                          "hbm_gbs": gbs, "hbm_frac_of_measured_peak": gbs / hbm_peak}
        # ---- integer critic (discriminator_mini.v) over the same frames as candidate, the generator's output as condition
        Wrom[256:752] = np.clip(np.rint(rng.standard_normal(496) * 30), -128, 127)
        Brom[32:57] = rng.integers(-64, 64, 25)
        for mode, tag in ((ops.GEN_Q_SPEC, "spec"), (ops.GEN_Q_RTL, "rtl_literal")):
            for _ in range(3):
                sq = ops.disc_fwd_q(yq, xq, Wrom, Brom, mode=mode)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                sq = ops.disc_fwd_q(yq, xq, Wrom, Brom, mode=mode)   # 2 GiB read per launch >> 126 MB L2
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1)) / K
            gbs = Q_FRAMES * 130 / (ms * 1e-3) / 1e9                  # 64 + 64 bytes in, 2 bytes out per frame
            also["fixed_point_critic_" + tag] = {"frames_per_s": Q_FRAMES * world / (ms * 1e-3), "ms_per_launch": ms,
                                                 "frames": Q_FRAMES, "hbm_gbs": gbs, "hbm_frac_of_measured_peak": gbs / hbm_peak}
        del xq, yq, sq
        # ---- config 1 at scale: fp32 MiniGenerator forward over HBM-resident frames (256 B of traffic per frame)
        xf = torch.randn(Q_FRAMES, 2, 16, generator=g, device=dev)
        for _ in range(3):
            yf = ops.gen_fwd_f32(xf, gp_d)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            yf = ops.gen_fwd_f32(xf, gp_d)                          # 4 GiB of traffic per launch >> 126 MB L2
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / K
        gbs = Q_FRAMES * 256 / (ms * 1e-3) / 1e9
        also["generator_fp32_hbm"] = {"frames_per_s": Q_FRAMES * world / (ms * 1e-3), "ms_per_launch": ms, "frames": Q_FRAMES,
                                      "hbm_gbs": gbs, "hbm_frac_of_measured_peak": gbs / hbm_peak,
                                      "fp32_tflops": Q_FRAMES * FLOP_GEN / (ms * 1e-3) / 1e12}
        del xf, yf
        # ---- the dataset path: SyntheticOFDMDataset batches written to HBM (clean + noisy + snr = 260 B per frame), AWGN with SNR ~ U(0, 30)
        # (config.yaml:20) and the --nonlinear chain; 2^22 frames per launch (1.1 GB of writes >> 126 MB L2)
        DS_FRAMES = 1 << 22
        for tag, dkw in (("awgn", dict()), ("nonlinear", dict(nonlinear=True, pa_saturation=0.8))):
            dcfg = ops.make_cfg(normalize=1, snr_lo=0.0, snr_hi=30.0, **dkw)
            for _ in range(3):
                dc, dn, dsn = ops.chan_sim(dcfg, DS_FRAMES, seed=3, frame0=rank * DS_FRAMES, device=dev)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s_ in range(K):
                dc, dn, dsn = ops.chan_sim(dcfg, DS_FRAMES, seed=3, frame0=(s_ * world + rank) * DS_FRAMES, device=dev)
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1)) / K
            also["dataset_" + tag] = {"frames_per_s": DS_FRAMES * world / (ms * 1e-3), "ms_per_launch": ms, "frames": DS_FRAMES,
                                      "hbm_write_gbs": DS_FRAMES * 260 / (ms * 1e-3) / 1e9,
                                      "what": "ofdmgan_chan_sim = SyntheticOFDMDataset.__getitem__ x frames (utils/dataset.py:236-293), k_sim_lean<-1>"}
            del dc, dn, dsn
        # ---- the separately named fast-RNG workload: the primary workload with Philox4x32-7 (the smallest round count Random123 documents
        # as passing BigCrush) instead of Philox4x32-10.  NOT the headline and NOT the parity workload: a different random stream.
        fcfg = ops.make_cfg(rng_rounds=7, **WORKLOAD)
        ftab = torch.zeros_like(table)
        for s_ in range(3):
            ops.sim_gen_metrics(fcfg, F, gparams=gp_d, seed=1, frame0=(s_ * world + rank) * F, out=ftab)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_ in range(K):
            ops.sim_gen_metrics(fcfg, F, gparams=gp_d, seed=1, frame0=(s_ * world + rank) * F, out=ftab)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / K
        also["fused_philox7"] = {"workload": "C4-fastrng: the primary workload with Philox4x32-7 draws (ofdmgan_chan_cfg.rng_rounds = 7)",
                                 "frames_per_s": F * world / (ms * 1e-3), "ms_per_step": ms, "frames_per_gpu": F,
                                 "fp32_frac_of_ffma_peak": F * flop_frame / (ms * 1e-3) / 1e12 / ffma}
        # ---- QPSK variant of the primary workload (QAMModulator + OFDMModulator source, N = 16, no pilots / CP): adds hard-decision BER
        qcfg = ops.make_cfg(symbol_source=ops.SYM_QPSK, n_fft=16, cp_len=0, pilot_spacing=0, **WORKLOAD)
        qtab = None
        for _ in range(3):
            qtab = ops.sim_gen_metrics(qcfg, F, gparams=gp_d, seed=2, frame0=rank * F)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(K):
            qtab = ops.sim_gen_metrics(qcfg, F, gparams=gp_d, seed=2, frame0=(s * world + rank) * F)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / K
        qs = ops.metrics_summary(qtab)
        also["fused_qpsk"] = {"frames_per_s": F * world / (ms * 1e-3), "ms_per_step": ms, "frames_per_gpu": F,
                              "noeq_ber_per_snr": [float(v) for v in qs["ber"][:, 1]], "gan_ber_per_snr": [float(v) for v in qs["ber"][:, 0]]}
        # ---- config 5: benchmark_comparison.run_benchmark with n_trials x 10^4 (10^6 trials x 7 SNRs x 2 scenarios), GAN / ZF / MMSE /
        # NoEQ rows, sharded by frame index over the ranks, one final all-reduce
        from ofdm_gan_sr_b200.sweep import run_benchmark
        run_benchmark(gp_d, n_trials=1000, nonlinear=True, pa_saturation=0.8, device=dev)
        barrier()
        t0 = time.perf_counter()
        res_lin = run_benchmark(gp_d, n_trials=1_000_000, nonlinear=False, device=dev, seed=3)
        res_nl = run_benchmark(gp_d, n_trials=1_000_000, nonlinear=True, pa_saturation=0.8, device=dev, seed=4)
        barrier()
        c5_s = max_over_ranks(time.perf_counter() - t0)
        also["benchmark_c5"] = {"trials": 14_000_000, "seconds": c5_s, "trials_per_s": 14_000_000 / c5_s, "methods": list(res_nl),
                                "evm_db_at_10dB_nonlinear": {m: res_nl[m][10.0]["evm"] for m in res_nl},
                                "evm_db_at_10dB_linear": {m: res_lin[m][10.0]["evm"] for m in res_lin},
                                "note": "untrained random-init generator: the GAN row is not a quality claim"}
        # ---- config 3: CWGAN-GP step, 65,536 frames per GPU, data-parallel
        Bt = TRAIN_FRAMES_PER_GPU
        tcfg = ops.make_cfg(normalize=1, snr_lo=0.0, snr_hi=30.0)
        clean, noisy, _ = ops.chan_sim(tcfg, Bt, seed=0, frame0=rank * Bt, device=dev)
        # the iteration is replayed as one CUDA graph (device-resident step counters; with several GPUs the peer-memory exchange
        # kernel is a node of that graph)
        trainer = CWGANGPStep(gp_h, dp_h, device=dev, graph=True)
        for _ in range(3):
            trainer.step(clean, noisy)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            trainer.step(clean, noisy)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / K
        st = trainer.stats()
        assert np.isfinite(st["d_loss"]) and np.isfinite(st["g_loss"])
        ms_eager = small = None
        if world == 1:                                           # the same iteration launched eagerly, and the reference's default batch of 64
            def timed_steps(tr, c, n_, reps):
                for _ in range(3):
                    tr.step(c, n_)
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(reps):
                    tr.step(c, n_)
                g1.record()
                torch.cuda.synchronize()
                return g0.elapsed_time(g1) / reps
            ms_eager = timed_steps(CWGANGPStep(gp_h, dp_h, device=dev), clean, noisy, K)
            c64, n64 = clean[:64].contiguous(), noisy[:64].contiguous()
            small = {"batch": 64, "ms_per_step_graph": timed_steps(CWGANGPStep(gp_h, dp_h, device=dev, graph=True), c64, n64, 50),
                     "ms_per_step_eager": timed_steps(CWGANGPStep(gp_h, dp_h, device=dev), c64, n64, 50)}
        ms_nccl = None
        if world > 1 and trainer.comm is not None:               # the same step with dist.all_reduce + the Adam kernel, for comparison
            t2 = CWGANGPStep(gp_h, dp_h, device=dev, exchange="nccl")
            for _ in range(3):
                t2.step(clean, noisy)
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(K):
                t2.step(clean, noisy)
            f1.record()
            barrier()
            ms_nccl = max_over_ranks(f0.elapsed_time(f1)) / K
        # Where the data-parallel step's extra time goes: (a) the same iteration on every rank WITHOUT any exchange (each rank alone on
        # its shard: the compute), (b) the exchange alone in a tight loop (ranks in lockstep: pure store + poll latency over NVLink).
        # (step - compute) / 6 exchanges - latency = what is left for waiting on the slowest rank (skew between the ranks' kernels).
        exch = None
        if world > 1 and trainer.comm is not None:
            t3 = CWGANGPStep(gp_h, dp_h, device=dev, graph=True, data_parallel=False)   # no exchange: every rank alone on its shard
            for _ in range(3):
                t3.step(clean, noisy)
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(K):
                t3.step(clean, noisy)
            f1.record()
            barrier()
            ms_compute = max_over_ranks(f0.elapsed_time(f1)) / K
            buf = torch.zeros(528, device=dev)
            for _ in range(20):
                trainer.comm.allreduce_adam(buf)
            barrier()
            n_ex = 400
            f0.record()
            for _ in range(n_ex):
                trainer.comm.allreduce_adam(buf)
            f1.record()
            barrier()
            us_exchange = max_over_ranks(f0.elapsed_time(f1)) / n_ex * 1e3
            per_ex = (ms - ms_compute) * 1e3 / 6.0
            exch = {"ms_per_step_compute_only": ms_compute, "exchanges_per_step": 6, "us_per_exchange_in_step": per_ex,
                    "us_per_exchange_alone_incl_launch": us_exchange,
                    "note": "in step = (data-parallel step - the same iteration on every rank without any exchange) / 6: store to every "
                            "peer, poll, rank-ordered sum, plus waiting for the slowest rank; alone = one stand-alone launch of the "
                            "exchange kernel per call in a tight loop (inside the step five of the six exchanges are the tail of the "
                            "critic kernel and cost no launch)"}
        tflops = Bt * FLOP_TRAIN / (ms * 1e-3) / 1e12
        # end to end: every step's batch comes from pinned host memory (H2D inside the timed region, double-buffered on a copy
        # stream so the transfer of batch i+1 overlaps the compute of batch i - what a prefetching loader does), and the step's
        # statistics go back to the host
        hc, hn = clean.cpu().pin_memory(), noisy.cpu().pin_memory()
        bufs = [(torch.empty_like(clean), torch.empty_like(noisy)) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[i % 2])
                bufs[i % 2][0].copy_(hc, non_blocking=True)
                bufs[i % 2][1].copy_(hn, non_blocking=True)
                ready[i % 2].record(copy_stream)

        for e in freed:
            e.record()
        trainer.collect_stats(trainer.request_stats())           # allocate the pinned read-back buffers outside the timed region
        barrier()
        Ke = max(K, 40)                                          # long enough that the un-overlappable first copy (0.7 ms) is amortised
        t0 = time.perf_counter()
        prefetch(0)
        ticket = None
        for i in range(Ke):
            if i + 1 < Ke:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            trainer.step(*bufs[i % 2])
            freed[i % 2].record()
            nxt = trainer.request_stats()                        # D2H of this step's losses into pinned memory, asynchronous
            if ticket is not None:
                trainer.collect_stats(ticket)                    # the previous step's losses are read while this step runs
            ticket = nxt
        last_stats = trainer.collect_stats(ticket)
        assert np.isfinite(last_stats["d_loss"])
        barrier()
        e2e_train = Bt * world * Ke / max_over_ranks(time.perf_counter() - t0)
        also["train"] = {"samples_per_s": Bt * world / (ms * 1e-3), "ms_per_step": ms, "frames_per_gpu": Bt, "n_critic": 5,
                         "fp32_tflops_per_gpu": tflops, "fp32_frac_of_ffma_peak": tflops / ffma,
                         "fp32_frac_of_ffma_peak_executed_work": tflops / ffma * FLOP_TRAIN_EXECUTED / FLOP_TRAIN,
                         "flop_per_sample": {"credited": FLOP_TRAIN, "executed": FLOP_TRAIN_EXECUTED,
                                             "note": "credited = SURVEY.md 8(d), the reference's own iteration (a generator forward per critic update); executed = "
                                                     "what this step computes: the generator is unchanged during the critic updates, so its forward runs once"},
                         "e2e_samples_per_s": e2e_train, "e2e_steps": Ke, "e2e_h2d_bytes_per_step": 2 * Bt * 128, "e2e_d2h_bytes_per_step": 28 * 4,
                         "launches_per_step": trainer.launches_per_step(), "d_loss": st["d_loss"], "g_loss": st["g_loss"],
                         "exchange": "none (1 GPU)" if world == 1 else ("peer-memory all-reduce fused with Adam" if trainer.comm is not None else "nccl"),
                         "ms_per_step_with_nccl_exchange": ms_nccl, "exchange_breakdown": exch,
                         "launch": "one CUDA graph per iteration" if trainer.use_graph else "eager, 27 launches per iteration",
                         "ms_per_step_eager": ms_eager, "reference_default_batch": small}
        launches += 4 * K * 1 + K * 2 + K * 3 + K * trainer.launches_per_step() + 2 * K
        if rank == 0:
            also["torch_eager"] = torch_eager_bar(gp_d, torch.as_tensor(dp_h, device=dev), clean, noisy, K)
        barrier()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        import oracle                                            # checker / CPU baseline leg only
        ocfg = oracle.make_cfg(**WORKLOAD)
        cpu_threads = oracle.set_threads()
        t0 = time.perf_counter()
        oracle.sim_gen_metrics(ocfg, 0, 1 << 17, gparams=gp_h, seed=1)
        rate = (1 << 17) / (time.perf_counter() - t0)
        sample = int(min(1 << 25, max(1 << 17, 1 << int(np.log2(max(rate * 12.0, 1.0))))))
        t0 = time.perf_counter()
        om = oracle.sim_gen_metrics(ocfg, 0, sample, gparams=gp_h, seed=1)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": sample / dt, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                        "sample": f"{sample} frames of the same workload, oracle/channel.c + fp32_models.c with OpenMP on all host cores"}
        # the same sample through the GPU path must tell the same story (parity spot check inside the bench)
        gm = ops.sim_gen_metrics(cfg, sample, gparams=gp_d, seed=1).cpu().numpy()
        assert np.array_equal(gm[:, :2, 0], om[:, :2, 0])
        assert np.allclose(gm[:, :2, 3], om[:, :2, 3], rtol=1e-4), "GPU and CPU sweeps disagree"

    # the UNMODIFIED reference (Python / NumPy / PyTorch CPU) cannot travel to the GPU box (/root/reference does not exist there): its
    # numbers were taken in the build container by tools/time_reference_here.py and are quoted with their host
    cpu_unmodified = None
    upath = os.path.join(ROOT, "profiles", "r1_reference_cpu_container.json")
    if os.path.exists(upath):
        cpu_unmodified = {"host": "build container (8 cores), not this box", "source": "profiles/r1_reference_cpu_container.json",
                          "numbers": json.load(open(upath))}
    if rank == 0:
        tr = also.get("train") or {}
        summary = {"frames_per_s": value, "ms_per_step": ms_step, "roofline_frac": achieved / ffma, "e2e_frames_per_s": e2e_value,
                   "sustained_frames_per_s": sustained and sustained["value"],
                   "fast_rng_workload_frames_per_s": (also.get("fused_philox7") or {}).get("frames_per_s"),
                   "fast_rng_workload_frac": (also.get("fused_philox7") or {}).get("fp32_frac_of_ffma_peak"),
                   "train_ms_per_step": tr.get("ms_per_step"),
                   "train_samples_per_s": tr.get("samples_per_s"), "train_frac_of_ffma_peak": tr.get("fp32_frac_of_ffma_peak"), "train_frac_of_ffma_peak_executed_work": tr.get("fp32_frac_of_ffma_peak_executed_work"),
                   "train_ms_per_step_nccl": tr.get("ms_per_step_with_nccl_exchange"), "train_exchange_breakdown": tr.get("exchange_breakdown"),
                   "multi_gpu_parity": parity["status"] if parity else ("n/a (1 GPU)" if world == 1 else None), "n_gpus": world}
        line = {"metric": "OFDM frames/s: fused channel sim + G inference", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_NAME, "frames_per_gpu_per_step": F, "parallelism": f"frame-sharded x{world}",
                           "l2": "no HBM inputs: frames are generated on-chip from Philox counters and reduced on-chip",
                           "weights": "random-init (Xavier-uniform, seed 0)",
                           "train": {"workload": "C3: CWGAN-GP step (n_critic 5, GP weight 10) at 65,536 frames per GPU, data-parallel",
                                     "ms_per_step": tr.get("ms_per_step"), "samples_per_s": tr.get("samples_per_s"),
                                     "frac_of_ffma_peak": tr.get("fp32_frac_of_ffma_peak"), "exchange": tr.get("exchange"),
                                     "ms_per_step_with_nccl_exchange": tr.get("ms_per_step_with_nccl_exchange")},
                           "multi_gpu_parity": parity},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "ofdmgan_sim_gen_metrics_host (host weights in, host metric table out, synchronous)"},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "cpu_baseline_unmodified": cpu_unmodified, "per_step_ms": per_step_ms, "also": also, "summary": summary}
        _emit(saved_stdout, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""The loop of the reference's `python train.py --synthetic [--nonlinear]` (train.py:447-536) on the B200 path:
GPU-generated batches -> fused 5+1 CWGAN-GP step -> periodic validation with the fused sweep.

    python examples/train_synthetic.py [--steps 2000] [--batch 4096] [--nonlinear]
    torchrun --nproc-per-node N examples/train_synthetic.py ...      (data-parallel)
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ofdm_gan_sr_b200 as pkg  # noqa: E402
from ofdm_gan_sr_b200.models import MiniDiscriminator, MiniGenerator  # noqa: E402
from ofdm_gan_sr_b200.sweep import run_benchmark  # noqa: E402
from ofdm_gan_sr_b200.train_step import CWGANGPStep  # noqa: E402
from ofdm_gan_sr_b200.utils import SyntheticOFDMDataset, create_dataloader  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=4096, help="per-GPU batch")
    ap.add_argument("--nonlinear", action="store_true")
    ap.add_argument("--pa_saturation", type=float, default=0.8)
    ap.add_argument("--lr", type=float, default=2e-4)
    ap.add_argument("--log_every", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--graph", action="store_true", help="replay each iteration as one CUDA graph")
    args = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    torch.manual_seed(args.seed)
    G, D = MiniGenerator().cuda(), MiniDiscriminator().cuda()
    ds = SyntheticOFDMDataset(n_samples=args.batch * world * args.steps, snr_range=(0, 30), nonlinear=args.nonlinear,
                              pa_saturation=args.pa_saturation, seed=args.seed)
    loader = create_dataloader(ds, batch_size=args.batch, rank=rank, world_size=world)
    step = CWGANGPStep(G, D, lr_g=args.lr, lr_d=args.lr, seed=args.seed, graph=args.graph)
    history, t0 = [], time.time()
    for i, batch in enumerate(loader):
        step.step(batch["clean"], batch["noisy"])
        if (i + 1) % args.log_every == 0 or i == 0:
            st = step.stats()
            history.append((i + 1, st["rec_loss"], st["d_loss"], st["gradient_penalty"]))
            if rank == 0:
                print(f"step {i + 1:6d}  rec {st['rec_loss']:.4f}  d_loss {st['d_loss']:+.4f}  gp {st['gradient_penalty']:.4f}  "
                      f"W {st['wasserstein_distance']:+.4f}  {(i + 1) * args.batch * world / (time.time() - t0):.3g} samples/s", flush=True)
    step.store_to(G, D)
    step.close()                                                 # unmaps the peer exchange blocks (a barrier: every rank calls it)
    res = run_benchmark(G, n_trials=20000, nonlinear=args.nonlinear, pa_saturation=args.pa_saturation, seed=123)
    if rank == 0:
        for snr in res["GAN"]:
            print(f"SNR {snr:4.0f} dB   EVM  GAN {res['GAN'][snr]['evm']:7.2f}   MMSE {res['MMSE'][snr]['evm']:7.2f}   NoEQ {res['NoEQ'][snr]['evm']:7.2f} dB")
    if world > 1:
        dist.destroy_process_group()
    return history, res


if __name__ == "__main__":
    main()

"""The CWGAN-GP optimisation step of the reference trainer (train.py:201-344: n_critic x train_discriminator, then
train_generator), as fused libofdmgan launches on flat parameter vectors, data-parallel over torch.distributed.

Per step and per rank:  1 generator forward (the reference recomputes the same G(noisy) in each of the 5 critic
iterations, train.py:225-226,333-334 - it is computed once here), n_critic x [critic_step -> allreduce(528 floats)
-> Adam(521)], then gen_step -> allreduce(264 floats) -> Adam(258).  Loss statistics ride in the same buffers as the
gradients, so a step needs no device->host synchronisation; `stats()` fetches them with one copy when asked.
"""
import torch
import torch.distributed as dist

from . import ops
from ._lib import CRITIC_OUT, D_NPARAMS, G_NPARAMS, GEN_OUT, OfdmGanError


class CWGANGPStep:
    """State: flat generator / critic parameters + Adam moments on one CUDA device.

    Hyper-parameters and their defaults follow CWGANGPTrainer._setup_config (train.py:146-185) and config/config.yaml:
    Adam(lr 2e-4, betas (0.0, 0.9), eps 1e-8), n_critic 5, gp_weight 10, rec_weight 100, adv_weight 1.
    """

    def __init__(self, gparams, dparams, lr_g=2e-4, lr_d=2e-4, betas=(0.0, 0.9), eps=1e-8, n_critic=5, gp_weight=10.0,
                 rec_weight=100.0, adv_weight=1.0, leaky_slope=0.2, seed=0, process_group=None, device=None, backend=ops, exchange="auto", graph=False,
                 data_parallel=True, reuse_fake=True):
        """`backend` is the kernel namespace (default: libofdmgan through `ops`).  It exists so the host-side logic
        of this class (sharding, all-reduce, optimiser bookkeeping) can be exercised by the CPU test-suite with a
        stand-in; the product never passes anything but `ops`."""
        self.k = backend
        # reuse_fake: the generator update is fed the G(noisy) of this iteration's critic updates (the generator has not changed in
        # between, so train.py:285's forward IS that one) instead of recomputing it; False = the step computes its own forward
        self.reuse_fake = bool(reuse_fake)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if backend is ops else torch.device("cpu")
        self.device = torch.device(device)
        if backend is ops and self.device.type != "cuda":
            raise OfdmGanError("CWGANGPStep needs a CUDA device (libofdmgan has no CPU path)")
        f = lambda t, n: self._flat(t, n)
        self.g, self.d = f(gparams, G_NPARAMS), f(dparams, D_NPARAMS)
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)
        self.g_m, self.g_v, self.d_m, self.d_v = z(G_NPARAMS), z(G_NPARAMS), z(D_NPARAMS), z(D_NPARAMS)
        self._graph = None
        self.lr_g, self.lr_d, self.betas, self.eps = lr_g, lr_d, tuple(betas), eps
        self.n_critic, self.gp_weight, self.rec_weight, self.adv_weight = n_critic, gp_weight, rec_weight, adv_weight
        self.slope, self.seed = leaky_slope, seed
        self.group = process_group
        # data_parallel=False: a single replica even inside an initialised process group (no exchange; bench.py's compute-only leg)
        self.distributed = bool(data_parallel) and dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        self.world = dist.get_world_size(process_group) if self.distributed else 1
        self.d_steps = self.g_steps = 0
        self._dout = torch.zeros(max(n_critic, 1), CRITIC_OUT, dtype=torch.float32, device=self.device)
        self._gout = torch.zeros(GEN_OUT, dtype=torch.float32, device=self.device)
        self._fake = None
        # graph=True: the 27 launches of an iteration are captured once per batch shape and replayed as one CUDA graph; the step
        # counters the kernels need (Philox alpha counter, Adam bias-correction step, the peer exchange's sequence number) then live
        # on the device.  With several ranks it needs the peer-memory exchange (a captured graph cannot hold the NCCL fallback here).
        self.use_graph = bool(graph) and backend is ops
        self._ctr = torch.zeros(2, dtype=torch.int32, device=self.device) if self.use_graph else None   # [critic steps, generator steps]
        self._graphs, self._seen, self._static, self._warm_shape = {}, [], None, None
        # gradient exchange: "peer" = all-reduce fused with Adam over NVLink peer memory (one launch, ops.PeerComm),
        # "nccl" = dist.all_reduce then the Adam kernel, "auto" = peer when the ranks can map each other's memory
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError(f"exchange must be 'auto', 'peer' or 'nccl', got {exchange!r}")
        self.comm = None
        if self.distributed and backend is ops and exchange != "nccl":
            try:
                self.comm = ops.PeerComm(process_group, self.device)
            except OfdmGanError:
                if exchange == "peer":
                    raise
            # all ranks must agree: one rank without peer access sends everybody to NCCL
            ok = torch.tensor([1 if self.comm is not None else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
            if int(ok.item()) == 0 and self.comm is not None:
                self.comm.close()
                self.comm = None
        if self.distributed and self.comm is None:
            self.use_graph = False                               # NCCL exchange: eager launches

    # Host scalars are baked into a captured CUDA graph: assigning any of them (e.g. the StepLR halving of train.py:497-514 writing
    # lr_g / lr_d) drops the captured graph, and the next step() captures a new one with the current values.
    _MAX_GRAPHS = 5                                              # the static pair's graph + four caller-buffer graphs
    _GRAPH_SCALARS = ("lr_g", "lr_d", "betas", "eps", "n_critic", "gp_weight", "rec_weight", "adv_weight", "slope", "seed")

    def __setattr__(self, name, value):
        if name in CWGANGPStep._GRAPH_SCALARS and getattr(self, "_graph", None) is not None and getattr(self, name, None) != value:
            object.__setattr__(self, "_graph", None)             # every captured graph is stale; scratch stays sized
        object.__setattr__(self, name, value)

    # ---- checkpointing: everything train.py:411-445 saves for the two optimisers, in torch.optim.Adam's own layout
    def state_dict(self):
        """Parameters, Adam moments and step counters (flat fp32 vectors in state_dict order)."""
        ctr = self._ctr.cpu().tolist() if self._ctr is not None and self.use_graph else None
        return {"g": self.g.clone(), "d": self.d.clone(), "g_m": self.g_m.clone(), "g_v": self.g_v.clone(), "d_m": self.d_m.clone(),
                "d_v": self.d_v.clone(), "d_steps": self.d_steps, "g_steps": self.g_steps, "device_counters": ctr,
                "hyper": {k: getattr(self, k) for k in self._GRAPH_SCALARS}}

    def load_state_dict(self, sd):
        for k in ("g", "d", "g_m", "g_v", "d_m", "d_v"):
            getattr(self, k).copy_(torch.as_tensor(sd[k]).to(self.device))
        self.d_steps, self.g_steps = int(sd["d_steps"]), int(sd["g_steps"])
        if self._ctr is not None:
            self._ctr.copy_(torch.tensor([self.d_steps, self.g_steps], dtype=torch.int32))
        for k, v in sd.get("hyper", {}).items():
            setattr(self, k, tuple(v) if k == "betas" else v)

    def adam_state_dict(self, which, module):
        """The state_dict a torch.optim.Adam over `module.parameters()` would hold after the same steps (`which` = "g" | "d"): what
        train.py:417-418 stores under 'optimizer_G_state_dict' / 'optimizer_D_state_dict'."""
        m, v, t, lr = (self.g_m, self.g_v, self.g_steps, self.lr_g) if which == "g" else (self.d_m, self.d_v, self.d_steps, self.lr_d)
        state, off = {}, 0
        for i, p in enumerate(module.parameters()):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(t)), "exp_avg": m[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": v[off:off + n].view_as(p).clone()}
            off += n
        group = {"lr": lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": list(range(len(state)))}
        return {"state": state, "param_groups": [group]}

    def load_adam_state_dict(self, which, sd):
        """Inverse of adam_state_dict: resume from a checkpoint written by the reference trainer (train.py:434-445)."""
        m, v = (self.g_m, self.g_v) if which == "g" else (self.d_m, self.d_v)
        off, t = 0, 0
        for i in sorted(sd["state"]):
            st = sd["state"][i]
            n = st["exp_avg"].numel()
            m[off:off + n].copy_(st["exp_avg"].reshape(-1).to(self.device))
            v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1).to(self.device))
            t = int(st["step"])
            off += n
        if off != m.numel():
            raise OfdmGanError(f"optimizer state holds {off} values, expected {m.numel()}")
        grp = sd["param_groups"][0]
        if which == "g":
            self.g_steps, self.lr_g = t, grp["lr"]
        else:
            self.d_steps, self.lr_d = t, grp["lr"]
        self.betas, self.eps = tuple(grp["betas"]), grp["eps"]
        if self._ctr is not None:
            self._ctr.copy_(torch.tensor([self.d_steps, self.g_steps], dtype=torch.int32))

    def _flat(self, t, n):
        if isinstance(t, torch.nn.Module):
            t = ops.flatten_params(t)
        t = torch.as_tensor(t, dtype=torch.float32).detach().reshape(-1).to(self.device).clone()
        if t.numel() != n:
            raise OfdmGanError(f"expected {n} parameters, got {t.numel()}")
        return t

    # ---- parameters <-> nn.Module (state_dict order, train.py:413-424 checkpoints stay interchangeable)
    def store_to(self, generator=None, discriminator=None):
        for vec, mod in ((self.g, generator), (self.d, discriminator)):
            if mod is None:
                continue
            off = 0
            with torch.no_grad():
                for p in mod.parameters():
                    p.copy_(vec[off:off + p.numel()].view_as(p))
                    off += p.numel()

    def load_from(self, generator=None, discriminator=None):
        if generator is not None:
            self.g.copy_(ops.flatten_params(generator).to(self.device))
        if discriminator is not None:
            self.d.copy_(ops.flatten_params(discriminator).to(self.device))

    def _allreduce(self, buf):
        if self.distributed:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)

    def _reduce_and_update(self, buf, p, m, v, lr, t):
        """sum `buf` (gradients + loss statistics) over the ranks, then Adam on p with its first p.numel() entries"""
        if self.comm is not None:
            self.comm.allreduce_adam(buf, p, m, v, lr, self.betas[0], self.betas[1], self.eps, t)
        else:
            self._allreduce(buf)
            self.k.adam(p, m, v, buf, lr, self.betas[0], self.betas[1], self.eps, t)

    def close(self):
        if self.comm is not None:
            self.comm.close()
            self.comm = None

    def _iteration_ctr(self, clean, noisy):
        """The same iteration as step() with every per-step scalar read from device memory: no argument changes between calls."""
        B = clean.shape[0]
        Bg = B * self.world

        def update(buf, p, m, v, lr, ctr):
            if self.comm is not None:
                self.comm.allreduce_adam(buf, p, m, v, lr, self.betas[0], self.betas[1], self.eps, step_dev=ctr)
            else:
                self.k.adam(p, m, v, buf, lr, self.betas[0], self.betas[1], self.eps, 0, step_dev=ctr)

        fused = hasattr(self.k, "critic_train") and hasattr(self.k, "gen_train")
        self._fake = self.k.gen_fwd_f32(noisy, self.g, self.slope)
        for c in range(self.n_critic):
            out = self._dout[c]
            if fused:
                # loss + backward, then ONE tail launch (gradient reduction, the sum over the ranks through peer memory, Adam,
                # weight-image refresh); from the second iteration on the image installed by the previous tail is still current
                self.k.critic_train(clean, noisy, self._fake, self.d, self.d_m, self.d_v, self._ctr[0:1], self.lr_d, self.betas[0],
                                    self.betas[1], self.eps, seed=self.seed, sample0=self.rank * B, gp_weight=self.gp_weight,
                                    slope=self.slope, out=out, image_is_current=c > 0, comm=self.comm, b_global=Bg)
                continue
            self.k.critic_step(clean, noisy, self._fake, self.d, seed=self.seed, sample0=self.rank * B, gp_weight=self.gp_weight,
                               slope=self.slope, b_global=Bg, out=out, alpha_iter_dev=self._ctr[0:1])
            update(out, self.d, self.d_m, self.d_v, self.lr_d, self._ctr[0:1])
        if self.reuse_fake and fused:
            # loss + backward, then ONE tail launch (reduction, sum over the ranks, Adam).  Inside this iteration the staging buffers
            # still hold the critic image the last critic tail wrote and the generator image gen_fwd_f32 built: copied, not rebuilt
            self.k.gen_train(clean, noisy, self._fake, self.g, self.g_m, self.g_v, self._ctr[1:2], self.lr_g, self.betas[0], self.betas[1],
                             self.eps, self.d, adv_weight=self.adv_weight, rec_weight=self.rec_weight, slope=self.slope, out=self._gout,
                             d_image_staged=self.n_critic > 0, g_image_staged=True, comm=self.comm, b_global=Bg)
            return
        self.k.gen_step(clean, noisy, self.d, self.g, self.adv_weight, self.rec_weight, self.slope, b_global=Bg, out=self._gout,
                        fake=self._fake if self.reuse_fake else None)
        update(self._gout, self.g, self.g_m, self.g_v, self.lr_g, self._ctr[1:2])

    def _step_graph(self, clean, noisy):
        """Replay the iteration as a CUDA graph.  Buffers that come back (a loop stepping on the same tensors, a double-buffered
        loader alternating between a few) get a graph captured ON THEM, keyed by (address of clean, address of noisy, shape): no
        copy at all.  An address seen for the first time goes through one pair of static buffers (two device-to-device copies,
        then the static graph), so a loader that hands out fresh memory every step never pays for a capture per step."""
        if not (clean.is_contiguous() and noisy.is_contiguous()):
            clean, noisy = clean.contiguous(), noisy.contiguous()
        shape = tuple(clean.shape)
        key = (clean.data_ptr(), noisy.data_ptr(), shape)
        if self._graph is None:                                   # (also after a hyper-parameter change: every captured graph is stale)
            self._graphs, self._seen = {}, []
            self._graph = self._graphs
        if self._warm_shape != shape:
            self._iteration_ctr(clean, noisy)                    # first call with this shape: eager (also sizes the library's scratch)
            self._warm_shape = shape
            self._static = (torch.empty_like(clean), torch.empty_like(noisy))
            self._graphs.clear()
        else:
            g = self._graphs.get(key)
            if g is None and key in self._seen and len(self._graphs) < self._MAX_GRAPHS:
                g = self._capture(clean, noisy)
                self._graphs[key] = g
            if g is None:                                         # first sighting of these addresses: the static pair
                self._seen = (self._seen + [key])[-8:]
                skey = ("static", shape)
                self._static[0].copy_(clean)
                self._static[1].copy_(noisy)
                g = self._graphs.get(skey)
                if g is None:
                    g = self._graphs[skey] = self._capture(*self._static)
            g.replay()
        self.d_steps += self.n_critic
        self.g_steps += 1

    def _capture(self, clean, noisy):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):                                 # capturing does not execute
            self._iteration_ctr(clean, noisy)
        return g

    def step(self, clean, noisy, alphas=None):
        """One trainer iteration on this rank's shard of the batch (train.py:327-344).

        clean, noisy: [B_local,2,16] CUDA tensors; B_local must be the same on every rank (the global batch is B_local x world).  alphas (optional, [n_critic,B_local]): the torch.rand draws of
        compute_gradient_penalty for parity runs; by default alpha comes from Philox(seed, global sample index,
        critic-step counter), so the global batch does not depend on the number of ranks.
        """
        if clean.shape[0] < 1:
            raise OfdmGanError("CWGANGPStep.step: empty local batch - every rank must hold the same, non-empty number of frames "
                               "(gradients are scaled by 1 / (B_local x world); GPUBatchLoader shards that way)")
        if self.use_graph:
            if alphas is not None:
                raise OfdmGanError("graph=True draws alpha from Philox; injected alphas need graph=False")
            return self._step_graph(clean, noisy)
        B = clean.shape[0]
        Bg = B * self.world
        self._fake = self.k.gen_fwd_f32(noisy, self.g, self.slope)
        for c in range(self.n_critic):
            out = self._dout[c]
            self.k.critic_step(clean, noisy, self._fake, self.d, alpha=None if alphas is None else alphas[c], seed=self.seed,
                            sample0=self.rank * B, alpha_iter=self.d_steps, gp_weight=self.gp_weight, slope=self.slope,
                            b_global=Bg, out=out)
            self.d_steps += 1
            self._reduce_and_update(out, self.d, self.d_m, self.d_v, self.lr_d, self.d_steps)
        self.k.gen_step(clean, noisy, self.d, self.g, self.adv_weight, self.rec_weight, self.slope, b_global=Bg, out=self._gout,
                        fake=self._fake if self.reuse_fake else None)
        self.g_steps += 1
        self._reduce_and_update(self._gout, self.g, self.g_m, self.g_v, self.lr_g, self.g_steps)

    def _packed_stats(self):
        d = self._dout[:, D_NPARAMS:D_NPARAMS + 5]
        g = self._gout[G_NPARAMS:G_NPARAMS + 3]
        return torch.cat([d.reshape(-1), g])

    def request_stats(self):
        """Start the device->host copy of this step's statistics into pinned memory without waiting for it: a training loop logs
        the previous step's numbers (`collect_stats`) while the GPU runs the next one, instead of draining the stream every step."""
        if getattr(self, "_pin", None) is None:
            self._pin = [torch.empty(self.n_critic * 5 + 3, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._pin_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._pin_i = 0
        i = self._pin_i = 1 - self._pin_i
        self._pin[i].copy_(self._packed_stats(), non_blocking=True)
        self._pin_ev[i].record()
        return i

    def collect_stats(self, ticket):
        """The statistics requested under `ticket` (waits only for that copy)."""
        self._pin_ev[ticket].synchronize()
        return self._unpack(self._pin[ticket].clone())

    def stats(self):
        """The scalars train.py:255-261,301-305 log, from the last step: one device->host copy."""
        if self.comm is not None:
            self.comm.check()
        return self._unpack(self._packed_stats().cpu())

    def _unpack(self, packed):
        last = packed[(self.n_critic - 1) * 5:self.n_critic * 5] if self.n_critic else torch.zeros(5)
        gs = packed[self.n_critic * 5:]
        return {"d_loss": float(last[0]), "wasserstein_distance": float(last[1]), "gradient_penalty": float(last[2]),
                "d_real_mean": float(last[3]), "d_fake_mean": float(last[4]), "g_loss": float(gs[0]), "adv_loss": float(gs[1]),
                "rec_loss": float(gs[2]), "critic_iterations": packed[:self.n_critic * 5].view(-1, 5).tolist()}

    # kernels launched by one step() (for bench.py's gpu_launches): G fwd 2 (weight image + kernel), per critic iteration
    # 4 (image, k_critic, finalize, adam), generator step 6 (2 images, k_gen_step, finalize, adam)... see DESIGN.md
    def launches_per_step(self):
        if self.use_graph:                                       # fused critic iterations: kernel + tail, one image refresh per step
            return 2 + (2 * self.n_critic + 1) + 2
        return 2 + self.n_critic * 4 + 5                         # (the fused exchange kernel takes the Adam kernel's place)

"""CUDA-graph execution of the adaptation iteration.

At batch 2 the iteration is launch-bound: about two thousand small kernels per step, each costing more host time than
device time.  ``GraphedRickAdapter`` records the sub-steps of ``RickAdapter.step`` -- D step (one joint pass over fake +
real), R1, G step, path-length, EMA -- once (forward, backward, the fused masks + Adam launch, and for the D step the
weight re-pack + tcgen05 generator forward) and replays them.  Everything data dependent is a device buffer written
before the replay:

    real images        static (B, 3, S, S) buffer, filled by ``copy_`` from the caller's tensor
    style mixing       z1, z2 are always drawn (device RNG inside the graph); the crossover index is a device scalar,
                       ``n_latent`` meaning "no mixing" -- latent = where(layer < index, w1, w2), which equals the
                       reference's concat of repeats (model_probe_tune.py:544-560)
    masks              the byte masks are device resident and updated in place by the Fisher round
    Adam step counts   per-parameter device floats (rick_b200.optim.FusedMaskedAdam)

The Fisher round (once per ``fisher_freq`` iterations) replays a sixth graph -- its per-image body: G forward, joint D
pass, two backward passes, grad**2 accumulation -- and keeps the exchange step, the percentile and the mask update eager
(164 -> 50 ms per round at 256 px).  Capturing a graph runs its body twice as warm-up; everything those executions
touch (the four networks, both optimisers' moments and step counts, the path-length mean, the Fisher accumulators, the
CUDA RNG state) is snapshotted before and restored after, so capture is invisible to the training trajectory:
iteration 0 of a graphed run starts from exactly the state an eager run starts from (``prepare()`` captures every
graph up front).  Under torchrun (world_size > 1) the DDP gradient all-reduce is captured inside the graphs (NCCL
supports stream capture), so every rank replays the same sequence of collectives.

Randomness.  Throughput mode (default): latents and per-layer noise are drawn by the device RNG inside the graphs; the
style-mixing decision and crossover index come from the ``draws`` stream handed to ``step`` (or, without one, from a
generator owned by the adapter and seeded at construction) -- never from the global ``random`` module.
``explicit_inputs=True`` (parity runs): latents, per-layer noise and the path-length image noise are static buffers
filled from ``draws`` in the reference's order before each replay, so a CPU oracle consuming the same ``DrawStream``
sees identical inputs.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Tuple

import torch
from torch import autograd, optim

from . import dist as rdist
from .adapt import (AdaptConfig, RickAdapter, d_logistic_loss, d_pair, d_r1_loss, g_nonsaturating_loss,
                    g_path_regularize)
from .fused import FusedGenerator


class GraphedRickAdapter(RickAdapter):
    def __init__(self, cfg: AdaptConfig, generator, discriminator, g_ema, d_ema, fused_generator: bool = True,
                 fused_optim: bool = True, explicit_inputs: bool = False, seed: int = 0):
        super().__init__(cfg, generator, discriminator, g_ema, d_ema, fused_adam=True, fused_generator=False,
                         fused_optim=fused_optim)
        # world_size > 1: the gradient all-reduce (NCCL) is recorded inside the graphs, between backward and Adam
        if cfg.warmup_iter != 0:
            raise RuntimeError("GraphedRickAdapter captures the post-warm-up iteration (warmup_iter must be 0)")
        g_ratio = cfg.g_reg_every / (cfg.g_reg_every + 1)
        d_ratio = cfg.d_reg_every / (cfg.d_reg_every + 1)
        if not self.fused_optim:                        # (the fused optimiser keeps its step count on the device already)
            self.g_optim = optim.Adam(self.g_train, lr=cfg.lr * g_ratio, betas=(0 ** g_ratio, 0.99 ** g_ratio),
                                      fused=True, capturable=True)
            self.d_optim = optim.Adam(self.d_train, lr=cfg.lr * d_ratio, betas=(0 ** d_ratio, 0.99 ** d_ratio),
                                      fused=True, capturable=True)
        dev = self.device
        self.fg = FusedGenerator(generator) if (fused_generator and FusedGenerator.supports(generator)) else None
        self._real = torch.zeros(cfg.batch, 3, cfg.size, cfg.size, device=dev)
        self._f_lat = torch.zeros(1, cfg.latent, device=dev)
        self._f_real = torch.zeros(1, 3, cfg.size, cfg.size, device=dev)
        self._inject = {k: torch.full((), generator.n_latent, dtype=torch.long, device=dev) for k in ("d", "g", "path")}
        self._layer = torch.arange(generator.n_latent, device=dev).view(1, -1, 1)
        self.explicit_inputs = bool(explicit_inputs)
        self._rng = random.Random(seed)                 # mixing decisions when step() is given no draw stream
        pb = max(1, cfg.batch // cfg.path_batch_shrink)
        self._batch_of = {"d": cfg.batch, "g": cfg.batch, "path": pb}
        if self.explicit_inputs:
            log_size = int(math.log2(cfg.size))
            shapes = [4] + [2 ** i for i in range(3, log_size + 1) for _ in range(2)]
            self._z = {k: torch.zeros(2, b, cfg.latent, device=dev) for k, b in self._batch_of.items()}
            self._noise = {k: [torch.zeros(b, 1, r, r, device=dev) for r in shapes] for k, b in self._batch_of.items()}
            self._path_noise = torch.zeros(pb, 3, cfg.size, cfg.size, device=dev)
        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self._outs: Dict[str, Dict[str, torch.Tensor]] = {}
        self._graph_launches: Dict[str, int] = {}      # rick_b200 kernel nodes recorded in each graph
        self.replayed_launches = 0                      # rick_b200 kernels executed through graph replays so far
        self._warm = False

    # ---- graph bodies -----------------------------------------------------------------------------------
    def _latent(self, batch: int, key: str) -> torch.Tensor:
        cfg = self.cfg
        z = self._z[key] if self.explicit_inputs else torch.randn(2, batch, cfg.latent, device=self.device)
        # the mapping network is not among the trained parameters (train:908-917): no autograd through it, which is
        # result-preserving and lets both latents go through the fused forward in one pass
        with torch.no_grad():
            w = self.g.map_latent(z.reshape(2 * batch, cfg.latent)).reshape(2, batch, cfg.latent)
            return torch.where(self._layer < self._inject[key], w[0].unsqueeze(1), w[1].unsqueeze(1))

    def _layer_noise(self, key: str):
        return self._noise[key] if self.explicit_inputs else None

    def _body_d(self):
        cfg = self.cfg
        with torch.no_grad():
            latent = self._latent(cfg.batch, "d")
            if self.fg is not None:
                fake_img, _ = self.fg([latent], input_is_latent=True, noise=self._layer_noise("d"))
            else:
                fake_img, _ = self.g([latent], input_is_latent=True, noise=self._layer_noise("d"))
        self._grad_mode("d")
        fake_pred, real_pred = d_pair(self.d, fake_img, self._real)
        d_loss = d_logistic_loss(real_pred, fake_pred)
        self.d.zero_grad(set_to_none=True)
        self._backward("d", d_loss)
        self._optim_step("d", True, force_masks=True)
        return {"d": d_loss.detach(), "real_score": real_pred.mean().detach(), "fake_score": fake_pred.mean().detach()}

    def _body_r1(self):
        cfg = self.cfg
        self._grad_mode("d")
        real_r = self._real.detach().clone().requires_grad_(True)
        real_pred, _ = self.d(real_r)
        real_pred = real_pred.view(real_r.size(0), -1).mean(dim=1).unsqueeze(1)
        r1_loss = d_r1_loss(real_pred, real_r)
        self.d.zero_grad(set_to_none=True)
        self._backward("d", cfg.r1 / 2 * r1_loss * cfg.d_reg_every + 0 * real_pred[0])
        self._optim_step("d", True, force_masks=True)
        return {"r1": r1_loss.detach()}

    def _body_g(self):
        cfg = self.cfg
        self._grad_mode("g")
        latent = self._latent(cfg.batch, "g")
        fake_img, _ = self.g([latent], input_is_latent=True, noise=self._layer_noise("g"))
        fake_pred, _ = self.d(fake_img)
        g_loss = g_nonsaturating_loss(fake_pred)
        self.g.zero_grad(set_to_none=True)
        self._backward("g", g_loss)
        self._optim_step("g", True, force_masks=True)
        return {"g": g_loss.detach()}

    def _body_path(self):
        cfg = self.cfg
        pb = max(1, cfg.batch // cfg.path_batch_shrink)
        self._grad_mode("g")
        latent = self._latent(pb, "path").requires_grad_(True)
        fake_img, latents = self.g([latent], input_is_latent=True, return_latents=True, noise=self._layer_noise("path"))
        path_noise = self._path_noise if self.explicit_inputs else torch.randn_like(fake_img)
        path_loss, path_mean, path_lengths = g_path_regularize(fake_img, latents, self.mean_path_length, path_noise)
        self.g.zero_grad(set_to_none=True)
        weighted = cfg.path_regularize * cfg.g_reg_every * path_loss
        if cfg.path_batch_shrink:
            weighted = weighted + 0 * fake_img[0, 0, 0, 0]
        self._backward("g", weighted)
        self._optim_step("g", True, force_masks=True)
        self.mean_path_length.copy_(path_mean)
        return {"path": path_loss.detach(), "path_length": path_lengths.mean().detach()}

    def _body_ema(self):
        if self.fused_optim:                            # one EMA-only launch per network
            self.g_optim.ema_only()
            self.d_optim.ema_only()
        else:
            self._ema()
        return {}

    # ---- capture / replay -------------------------------------------------------------------------------
    def _body_fisher(self):
        self._fisher_image(self._f_lat, self._f_real)
        return {}

    def fisher_round(self, latents: torch.Tensor, reals: torch.Tensor, layer_noise=None):
        """The Fisher round with its per-image body (G forward, joint D pass, two backward passes, grad**2 accumulation)
        replayed from a CUDA graph; the exchange step and the mask update stay eager.  Explicit per-layer noise (parity
        tests) takes the eager path."""
        if layer_noise is not None:
            return super().fisher_round(latents, reals, layer_noise)
        mine = self._fisher_begin(latents.shape[0])
        self._ensure("fisher")                          # capture (with its warm-up executions) BEFORE clearing
        self.acc_g.zero()
        self.acc_d.zero()
        for j in mine:
            self._f_lat.copy_(latents[j:j + 1], non_blocking=True)
            self._f_real.copy_(reals[j:j + 1], non_blocking=True)
            self._run("fisher")
        self._fisher_end()

    # ---- state that a graph body mutates: snapshotted around warm-up + capture ---------------------------------
    def _mutable_state(self) -> List[torch.Tensor]:
        ts: List[torch.Tensor] = []
        for net in (self.g, self.d, self.g_ema, self.d_ema):
            ts += [p.data for p in net.parameters()]
        for opt in (self.g_optim, self.d_optim):
            if self.fused_optim:
                ts += [opt.steps] + list(opt.exp_avg.values()) + list(opt.exp_avg_sq.values())
        ts += [self.mean_path_length] + list(self.acc_g.acc) + list(self.acc_d.acc)
        return ts

    def _snapshot(self):
        snap = [(t, t.clone()) for t in self._mutable_state()]
        grads = [(p, p.grad) for net in (self.g, self.d, self.g_ema, self.d_ema) for p in net.parameters()]
        torch_adam = None
        if not self.fused_optim:           # torch.optim.Adam creates its state lazily: remember what existed
            torch_adam = [{id(p): {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                           for p, st in opt.state.items()} for opt in (self.g_optim, self.d_optim)]
        return snap, grads, torch_adam, torch.cuda.get_rng_state(self.device)

    @torch.no_grad()
    def _restore(self, state):
        snap, grads, torch_adam, rng = state
        for t, saved in snap:
            t.copy_(saved)
        for p, g in grads:
            p.grad = g
        if torch_adam is not None:
            for opt, before in zip((self.g_optim, self.d_optim), torch_adam):
                for p, st in opt.state.items():
                    old = before.get(id(p))
                    for k, v in st.items():
                        if torch.is_tensor(v):          # in place: the graph recorded these very tensors
                            v.copy_(old[k]) if old is not None else v.zero_()
        torch.cuda.set_rng_state(rng, self.device)

    def _ensure(self, key: str):
        """Capture graph ``key`` (no-op when it exists).  The two warm-up executions the capture needs are rolled back:
        parameters, optimiser state, EMA copies, path-length mean, Fisher accumulators and the RNG state are the same
        after this call as before it."""
        if key in self._graphs:
            return
        body = getattr(self, "_body_" + key)
        state = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):              # warm-up executions on a side stream (rolled back below)
                body()
        torch.cuda.current_stream().wait_stream(side)
        from . import _lib
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.lib().rick_launch_count()
        with torch.cuda.graph(graph):
            outs = body()
        self._graph_launches[key] = int(_lib.lib().rick_launch_count() - n0)   # launches recorded, not executed
        self._graphs[key], self._outs[key] = graph, outs
        self._restore(state)

    def prepare(self, fisher: bool = True):
        """Capture every graph now.  Leaves the training state untouched (see ``_ensure``)."""
        if fisher:
            self._fisher_begin(1)
            self._ensure("fisher")
        for key in ("d", "r1", "g", "path", "ema"):
            self._ensure(key)
        return self

    def _run(self, key: str):
        self._ensure(key)
        self._graphs[key].replay()
        self.replayed_launches += self._graph_launches[key]
        return self._outs[key]

    def _draw_inputs(self, key: str, draws):
        """Everything random one sub-step consumes, in the reference's order (train:397, 501, 548; model_probe_tune.py:556):
        mixing decision, latents, crossover index, then per-layer noise.  Throughput mode only takes the mixing decision
        and the index from the host stream; latents / noise are drawn on the device inside the graph."""
        cfg = self.cfg
        n = self.g.n_latent
        b = self._batch_of[key]
        if self.explicit_inputs:
            if draws is None:
                raise RuntimeError("GraphedRickAdapter(explicit_inputs=True).step needs a DrawStream")
            z = draws.mixing_latents(b, cfg.latent, cfg.mixing)
            inject = draws.randint(1, n - 1) if len(z) == 2 else n
            self._z[key][0].copy_(z[0], non_blocking=True)
            if len(z) == 2:
                self._z[key][1].copy_(z[1], non_blocking=True)
            for dst, src in zip(self._noise[key], draws.layer_noise(b, cfg.size)):
                dst.copy_(src, non_blocking=True)
        elif draws is not None:
            mix = cfg.mixing > 0 and draws.uniform() < cfg.mixing
            inject = draws.randint(1, n - 1) if mix else n
        else:
            mix = cfg.mixing > 0 and self._rng.random() < cfg.mixing
            inject = self._rng.randint(1, n - 1) if mix else n
        self._inject[key].fill_(inject)

    def step(self, i: int, real_img: torch.Tensor, draws=None, explicit_layer_noise: bool = False):
        """One iteration (same sub-steps and order as ``RickAdapter.step``).  ``explicit_layer_noise`` is implied by
        ``explicit_inputs`` and ignored otherwise (throughput mode draws per-layer noise on the device)."""
        cfg = self.cfg
        self._real.copy_(real_img, non_blocking=True)
        out: Dict[str, torch.Tensor] = {}
        self._draw_inputs("d", draws)
        out.update(self._run("d"))
        if i % cfg.d_reg_every == 0:
            out.update(self._run("r1"))
        self._draw_inputs("g", draws)
        out.update(self._run("g"))
        if i % cfg.g_reg_every == 0:
            self._draw_inputs("path", draws)
            if self.explicit_inputs:
                self._path_noise.copy_(draws.normal(*self._path_noise.shape), non_blocking=True)
            out.update(self._run("path"))
        self._run("ema")
        return {k: v.clone() for k, v in out.items()} if self.explicit_inputs else out

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
(164 -> 50 ms per round at 256 px).  Capturing a graph runs its body twice as warm-up; for the training sub-steps those
are real optimiser steps, for the Fisher body the accumulators are cleared afterwards.  Under torchrun (world_size > 1)
the DDP gradient all-reduce is captured inside the graphs (NCCL supports stream capture), so every rank replays the
same sequence of collectives.
"""
from __future__ import annotations

import random
from typing import Dict

import torch
from torch import autograd, optim

from . import dist as rdist
from .adapt import (AdaptConfig, RickAdapter, d_logistic_loss, d_pair, d_r1_loss, g_nonsaturating_loss,
                    g_path_regularize)
from .fused import FusedGenerator


class GraphedRickAdapter(RickAdapter):
    def __init__(self, cfg: AdaptConfig, generator, discriminator, g_ema, d_ema, fused_generator: bool = True,
                 fused_optim: bool = True):
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
        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self._outs: Dict[str, Dict[str, torch.Tensor]] = {}
        self._graph_launches: Dict[str, int] = {}      # rick_b200 kernel nodes recorded in each graph
        self.replayed_launches = 0                      # rick_b200 kernels executed through graph replays so far
        for p in list(generator.parameters()) + list(discriminator.parameters()):
            p.requires_grad_(True)

    # ---- graph bodies -----------------------------------------------------------------------------------
    def _latent(self, batch: int, key: str) -> torch.Tensor:
        cfg = self.cfg
        z = torch.randn(2, batch, cfg.latent, device=self.device)
        w1, w2 = self.g.style(z[0]), self.g.style(z[1])
        return torch.where(self._layer < self._inject[key], w1.unsqueeze(1), w2.unsqueeze(1))

    def _body_d(self):
        cfg = self.cfg
        with torch.no_grad():
            latent = self._latent(cfg.batch, "d")
            if self.fg is not None:
                self.fg.refresh()
                fake_img, _ = self.fg([latent], input_is_latent=True)
            else:
                fake_img, _ = self.g([latent], input_is_latent=True)
        fake_pred, real_pred = d_pair(self.d, fake_img, self._real)
        d_loss = d_logistic_loss(real_pred, fake_pred)
        self.d.zero_grad(set_to_none=True)
        autograd.backward(d_loss, inputs=self.d_train)
        self._sync_grads(self.d_train)
        self._optim_step("d", True, force_masks=True)
        return {"d": d_loss.detach(), "real_score": real_pred.mean().detach(), "fake_score": fake_pred.mean().detach()}

    def _body_r1(self):
        cfg = self.cfg
        real_r = self._real.detach().clone().requires_grad_(True)
        real_pred, _ = self.d(real_r)
        real_pred = real_pred.view(real_r.size(0), -1).mean(dim=1).unsqueeze(1)
        r1_loss = d_r1_loss(real_pred, real_r)
        self.d.zero_grad(set_to_none=True)
        autograd.backward(cfg.r1 / 2 * r1_loss * cfg.d_reg_every + 0 * real_pred[0], inputs=self.d_train)
        self._sync_grads(self.d_train)
        self._optim_step("d", True, force_masks=True)
        return {"r1": r1_loss.detach()}

    def _body_g(self):
        cfg = self.cfg
        latent = self._latent(cfg.batch, "g")
        fake_img, _ = self.g([latent], input_is_latent=True)
        fake_pred, _ = self.d(fake_img)
        g_loss = g_nonsaturating_loss(fake_pred)
        self.g.zero_grad(set_to_none=True)
        autograd.backward(g_loss, inputs=self.g_train)
        self._sync_grads(self.g_train)
        self._optim_step("g", True, force_masks=True)
        return {"g": g_loss.detach()}

    def _body_path(self):
        cfg = self.cfg
        pb = max(1, cfg.batch // cfg.path_batch_shrink)
        latent = self._latent(pb, "path")
        fake_img, latents = self.g([latent], input_is_latent=True, return_latents=True)
        path_loss, path_mean, path_lengths = g_path_regularize(fake_img, latents, self.mean_path_length,
                                                               torch.randn_like(fake_img))
        self.g.zero_grad(set_to_none=True)
        weighted = cfg.path_regularize * cfg.g_reg_every * path_loss
        if cfg.path_batch_shrink:
            weighted = weighted + 0 * fake_img[0, 0, 0, 0]
        autograd.backward(weighted, inputs=self.g_train)
        self._sync_grads(self.g_train)
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

    def _ensure(self, key: str):
        if key not in self._graphs:
            body = getattr(self, "_body_" + key)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):          # warm-up executions on a side stream (these are real training steps)
                    body()
            torch.cuda.current_stream().wait_stream(side)
            from . import _lib
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.lib().rick_launch_count()
            with torch.cuda.graph(graph):
                outs = body()
            self._graph_launches[key] = int(_lib.lib().rick_launch_count() - n0)   # launches recorded, not executed
            self._graphs[key], self._outs[key] = graph, outs

    def _run(self, key: str):
        self._ensure(key)
        self._graphs[key].replay()
        self.replayed_launches += self._graph_launches[key]
        return self._outs[key]

    def _set_inject(self, key: str):
        cfg = self.cfg
        n = self.g.n_latent
        v = random.randint(1, n - 1) if (cfg.mixing > 0 and random.random() < cfg.mixing) else n
        self._inject[key].fill_(v)

    def step(self, i: int, real_img: torch.Tensor, draws=None, explicit_layer_noise: bool = False):
        cfg = self.cfg
        self._real.copy_(real_img, non_blocking=True)
        out: Dict[str, torch.Tensor] = {}
        self._set_inject("d")
        out.update(self._run("d"))
        if i % cfg.d_reg_every == 0:
            out.update(self._run("r1"))
        self._set_inject("g")
        out.update(self._run("g"))
        if i % cfg.g_reg_every == 0:
            self._set_inject("path")
            out.update(self._run("path"))
        self._run("ema")
        return out

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of one RICK adaptation run (train_dynamic_update_prune.py:193-699).

The reference's ``train()`` cannot be imported (it needs lmdb, lpips, wandb paths and a CUDA device), so the loop
body is restated here over the functional oracle (model_oracle.py) and the NumPy mask step (rick_oracle.py), in the
reference's order: warm-up gating (202-211), Fisher round (214-393), D step (396-438), R1 every ``d_reg_every``
(462-493), G step (500-540), path-length regularisation every ``g_reg_every`` (546-589), EMA (697-698).
Optimisers are ``torch.optim.Adam`` with the reference's parameter subsets and betas (908-931).

Random numbers come from a caller-supplied draw stream (duck-typed: ``mixing_latents``, ``randint``, ``layer_noise``,
``normal``) so that the CUDA path can consume the identical sequence.  As in the reference, the D step differentiates
through G as well (fake images are not detached, 401-420); those gradients are discarded by ``zero_grad`` (516), so
this restatement computes only the D gradients there -- same parameter updates, less CPU time.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
from torch import optim

from . import model_oracle as mo
from . import rick_oracle as ro


class OracleAdapter:
    def __init__(self, cfg, g: mo.Params, d: mo.Params, g_ema: mo.Params, d_ema: mo.Params):
        """``cfg`` needs the attributes of rick_b200.adapt.AdaptConfig; the four dicts are state_dict-like and are
        updated in place (tensors become leaf tensors that require grad)."""
        self.cfg = cfg
        self.size = cfg.size
        self.g_names, self.d_names = mo.g_param_names(cfg.size, cfg.n_mlp), mo.d_param_names(cfg.size)
        for sd, names in ((g, self.g_names), (d, self.d_names), (g_ema, self.g_names), (d_ema, self.d_names)):
            for n in names:
                sd[n] = sd[n].detach().clone().requires_grad_(True)
        self.g, self.d, self.g_ema, self.d_ema = g, d, g_ema, d_ema
        g_ratio = cfg.g_reg_every / (cfg.g_reg_every + 1)
        d_ratio = cfg.d_reg_every / (cfg.d_reg_every + 1)
        self.g_train = [n for n in self.g_names if "convs" in n]
        self.d_train = [n for n in self.d_names if ("convs" in n and "convs.0" not in n) or "final" in n]
        self.g_optim = optim.Adam([g[n] for n in self.g_train], lr=cfg.lr * g_ratio,
                                  betas=(0 ** g_ratio, 0.99 ** g_ratio))
        self.d_optim = optim.Adam([d[n] for n in self.d_train], lr=cfg.lr * d_ratio,
                                  betas=(0 ** d_ratio, 0.99 ** d_ratio))
        self.n_convs = 2 * (int(np.log2(cfg.size)) - 2)
        self.d_blocks = range(1, int(np.log2(cfg.size)) - 1)
        self.mean_path_length = 0
        self.ema_decay = 0.5 ** (32 / (10 * 1000))
        self.freeze_g = self.freeze_d = self.zero_g = self.zero_d = None
        self.rounds = 0
        self.fisher_g: Dict[str, np.ndarray] = {}
        self.fisher_d: Dict[str, np.ndarray] = {}

    # ---------------------------------------------------------------------------------------- Fisher round
    def fisher_round(self, latents: torch.Tensor, reals: torch.Tensor, layer_noise=None):
        cfg = self.cfg
        fg: Dict[str, np.ndarray] = {}
        fd: Dict[str, np.ndarray] = {}
        for j in range(latents.shape[0]):
            noise = None if layer_noise is None else layer_noise[j]
            fake, _ = mo.g_forward(self.g_ema, [latents[j:j + 1]], self.size, noise=noise, n_mlp=cfg.n_mlp)
            fake_pred = mo.d_forward(self.d_ema, fake, self.size)
            real_pred = mo.d_forward(self.d_ema, reals[j:j + 1], self.size)
            g_loss = mo.g_nonsaturating_loss(fake_pred)
            d_loss = mo.d_logistic_loss(real_pred, fake_pred)
            eg = mo.estimate_fisher(g_loss, self.g_ema, self.g_names)
            ed = mo.estimate_fisher(d_loss, self.d_ema, self.d_names)
            ro.fisher_accumulate(fg, {k: v.numpy() for k, v in eg.items()})
            ro.fisher_accumulate(fd, {k: v.numpy() for k, v in ed.items()})
        ro.fisher_average(fg, cfg.num_fisher_img, cfg.batch)
        ro.fisher_average(fd, cfg.num_fisher_img, cfg.batch)
        self.fisher_g, self.fisher_d = fg, fd
        self.freeze_g, _, prune_g, self.lines_g = ro.decide_g(fg, cfg.fisher_quantile, cfg.prune_quantile, self.n_convs)
        self.freeze_d, _, prune_d, self.lines_d = ro.decide_d(fd, cfg.fisher_quantile, cfg.prune_quantile, self.d_blocks)
        if self.rounds == 0:
            self.zero_g, self.zero_d = prune_g, prune_d
        else:
            self.zero_g, self.zero_d = ro.zero_idx_merge(self.zero_g, prune_g), ro.zero_idx_merge(self.zero_d, prune_d)
        self.rounds += 1

    # ---------------------------------------------------------------------------------------- helpers
    @staticmethod
    def _zero_grad(sd, names):
        for n in names:
            sd[n].grad = None

    def _apply_masks(self, sd, names, freeze, zero):
        """train:427-437 / 521-539 on torch tensors (same indexing as the reference: dim 1 for 5-D weights)."""
        if freeze is None:
            return
        with torch.no_grad():
            for n in names:
                p = sd[n]
                if n in freeze and p.grad is not None:
                    if p.ndim != 5:
                        p.grad[freeze[n]] = 0
                    else:
                        p.grad[:, freeze[n]] = 0
                if n in zero:
                    if p.ndim != 5:
                        p[zero[n]] = 0
                        if p.grad is not None:
                            p.grad[zero[n]] = 0
                    else:
                        p[:, zero[n]] = 0
                        if p.grad is not None:
                            p.grad[:, zero[n]] = 0

    def _gate(self, i):
        warm = i < self.cfg.warmup_iter
        for n in self.d_names:
            self.d[n].requires_grad_((not warm) or ("final" in n))

    # ---------------------------------------------------------------------------------------- one iteration
    def step(self, i: int, real_img: torch.Tensor, draws, explicit_layer_noise: bool = True):
        cfg = self.cfg
        after = i >= cfg.warmup_iter
        self._gate(i)
        noise_of = (lambda b: draws.layer_noise(b, cfg.size)) if explicit_layer_noise else (lambda b: None)
        out = {}
        n_latent = mo.g_n_latent(cfg.size)

        z = draws.mixing_latents(cfg.batch, cfg.latent, cfg.mixing)
        inject = draws.randint(1, n_latent - 1) if len(z) == 2 else None
        with torch.no_grad():
            fake, _ = mo.g_forward(self.g, z, self.size, noise=noise_of(cfg.batch), inject_index=inject, n_mlp=cfg.n_mlp)
        fake_pred = mo.d_forward(self.d, fake, self.size)
        real_pred = mo.d_forward(self.d, real_img, self.size)
        d_loss = mo.d_logistic_loss(real_pred, fake_pred)
        out["d"] = d_loss.detach()
        self._zero_grad(self.d, self.d_names)
        d_loss.backward()
        if after:
            self._apply_masks(self.d, self.d_names, self.freeze_d, self.zero_d)
        self.d_optim.step()

        if i % cfg.d_reg_every == 0:
            real_r = real_img.detach().clone().requires_grad_(True)
            rp = mo.d_forward(self.d, real_r, self.size)
            rp = rp.view(real_r.size(0), -1).mean(dim=1).unsqueeze(1)
            r1 = mo.d_r1_loss(rp, real_r)
            self._zero_grad(self.d, self.d_names)
            (cfg.r1 / 2 * r1 * cfg.d_reg_every + 0 * rp[0]).backward()
            if after:
                self._apply_masks(self.d, self.d_names, self.freeze_d, self.zero_d)
            self.d_optim.step()
            out["r1"] = r1.detach()

        z = draws.mixing_latents(cfg.batch, cfg.latent, cfg.mixing)
        inject = draws.randint(1, n_latent - 1) if len(z) == 2 else None
        if after:
            fake, _ = mo.g_forward(self.g, z, self.size, noise=noise_of(cfg.batch), inject_index=inject, n_mlp=cfg.n_mlp)
            g_loss = mo.g_nonsaturating_loss(mo.d_forward(self.d, fake, self.size))
            self._zero_grad(self.g, self.g_names)
            torch.autograd.backward(g_loss, inputs=[self.g[n] for n in self.g_train])
            self._apply_masks(self.g, self.g_names, self.freeze_g, self.zero_g)
            self.g_optim.step()
        else:
            with torch.no_grad():
                fake, _ = mo.g_forward(self.g, z, self.size, noise=noise_of(cfg.batch), inject_index=inject,
                                       n_mlp=cfg.n_mlp)
                g_loss = mo.g_nonsaturating_loss(mo.d_forward(self.d, fake, self.size))
        out["g"] = g_loss.detach()

        if i % cfg.g_reg_every == 0 and after:
            pb = max(1, cfg.batch // cfg.path_batch_shrink)
            z = draws.mixing_latents(pb, cfg.latent, cfg.mixing)
            inject = draws.randint(1, n_latent - 1) if len(z) == 2 else None
            fake, lat = mo.g_forward(self.g, z, self.size, noise=noise_of(pb), inject_index=inject, return_latents=True,
                                     n_mlp=cfg.n_mlp)
            pl, self.mean_path_length, plen = mo.g_path_regularize(fake, lat, self.mean_path_length,
                                                                   noise=draws.normal(*fake.shape))
            self._zero_grad(self.g, self.g_names)
            w = cfg.path_regularize * cfg.g_reg_every * pl
            if cfg.path_batch_shrink:
                w = w + 0 * fake[0, 0, 0, 0]
            torch.autograd.backward(w, inputs=[self.g[n] for n in self.g_train])
            self._apply_masks(self.g, self.g_names, self.freeze_g, self.zero_g)
            self.g_optim.step()
            out["path"], out["path_length"] = pl.detach(), plen.mean().detach()

        with torch.no_grad():
            for n in self.g_names:
                self.g_ema[n].mul_(self.ema_decay).add_(self.g[n], alpha=1 - self.ema_decay)
            for n in self.d_names:
                self.d_ema[n].mul_(self.ema_decay).add_(self.d[n], alpha=1 - self.ema_decay)
        return out

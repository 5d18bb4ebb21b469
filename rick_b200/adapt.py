"""The RICK adaptation iteration and FID-style sample generation on top of the sm_100a hot path.

Host-side mirror of the reference's ``train()`` loop body (train_dynamic_update_prune.py:193-699) and of the sample
loop of ``Evaluator.compute_inception_score`` (gan_training/eval.py:31-46), restructured so that the device is never
idle waiting for the host:

  * Fisher tensors, FIM vectors, thresholds and masks stay on the GPU (rick_b200/rick.py); the reference round-trips
    ~235 MB per Fisher image through NumPy and re-uploads index arrays every iteration.
  * No ``.item()`` / ``empty_cache()`` inside the iteration (the reference has ten such syncs, e.g. train:425, 596-603).
  * The D step runs G under ``no_grad`` and the G step differentiates only w.r.t. G's parameters.  Both are
    result-preserving: the reference back-propagates into the other network and then discards those gradients with
    ``zero_grad()`` (train:401-420, 516).  Likewise the mapping network (``style.*``) is evaluated without autograd:
    it is not among the parameters handed to the optimiser (train:908-917 keeps names containing 'convs'), and the
    path-length penalty differentiates w.r.t. its OUTPUT (train:548-566).
  * Optimiser selection, betas, warm-up gating, regulariser schedule, EMA and the order in which random numbers are
    consumed are the reference's.

All randomness comes from a ``DrawStream`` so that a seeded CPU oracle run can consume the identical sequence
(tests/test_adapt_gpu.py); for throughput runs the stream draws on the device.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
from torch import autograd, optim
from torch.nn import functional as F

from . import dist as rdist
from . import rick


@dataclass
class AdaptConfig:
    size: int = 256
    batch: int = 2
    latent: int = 512
    n_mlp: int = 8
    channel_multiplier: int = 2
    lr: float = 0.002
    r1: float = 10.0
    path_regularize: float = 2.0
    path_batch_shrink: int = 2
    d_reg_every: int = 16
    g_reg_every: int = 4
    mixing: float = 0.9
    warmup_iter: int = 0
    fisher_freq: int = 50
    num_fisher_img: int = 5
    fisher_quantile: float = 40.0
    prune_quantile: float = 0.1


class DrawStream:
    """Every random quantity one adaptation run consumes, in the reference's order.

    ``cpu_seeded=True`` draws from a seeded CPU generator and moves the result to ``device`` (bit-identical across
    machines: this is what parity tests use on both the CUDA path and the CPU oracle).  ``cpu_seeded=False`` draws on
    the device (throughput runs)."""

    def __init__(self, seed: int, device, cpu_seeded: bool = True):
        self.device = torch.device(device)
        self.cpu_seeded = cpu_seeded
        self.gen = torch.Generator(device="cpu" if cpu_seeded else self.device).manual_seed(seed)

    def normal(self, *shape) -> torch.Tensor:
        if self.cpu_seeded:
            return torch.randn(*shape, generator=self.gen).to(self.device, non_blocking=True)
        return torch.randn(*shape, generator=self.gen, device=self.device)

    def uniform(self) -> float:
        g = self.gen if self.cpu_seeded else None
        return float(torch.rand(1, generator=g))

    def randint(self, lo: int, hi: int) -> int:
        """inclusive bounds, like ``random.randint`` (model_probe_tune.py:556)"""
        g = self.gen if self.cpu_seeded else None
        return int(torch.randint(lo, hi + 1, (1,), generator=g))

    # ---- composite draws ----
    def mixing_latents(self, batch: int, latent: int, prob: float) -> List[torch.Tensor]:
        """mixing_noise (train:130-135)"""
        if prob > 0 and self.uniform() < prob:
            z = self.normal(2, batch, latent)
            return [z[0], z[1]]
        return [self.normal(batch, latent)]

    def layer_noise(self, batch: int, size: int) -> List[torch.Tensor]:
        """per-layer NoiseInjection draws in forward order (model_probe_tune.py:294-296)"""
        log_size = int(math.log2(size))
        out = [self.normal(batch, 1, 4, 4)]
        for i in range(3, log_size + 1):
            out += [self.normal(batch, 1, 2 ** i, 2 ** i), self.normal(batch, 1, 2 ** i, 2 ** i)]
        return out


def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    (grad_real,) = autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_path_regularize(fake_img, latents, mean_path_length, noise, decay=0.01):
    noise = noise / math.sqrt(fake_img.shape[2] * fake_img.shape[3])
    (grad,) = autograd.grad(outputs=(fake_img * noise).sum(), inputs=latents, create_graph=True)
    path_lengths = torch.sqrt(grad.pow(2).sum(2).mean(1))
    path_mean = mean_path_length + decay * (path_lengths.mean() - mean_path_length)
    path_penalty = (path_lengths - path_mean).pow(2).mean()
    return path_penalty, path_mean.detach(), path_lengths


def d_pair(d, fake: torch.Tensor, real: torch.Tensor):
    """``d(fake)[0], d(real)[0]`` in ONE pass over the discriminator (half the launches of the D step, the weights
    scaled once).  The only cross-sample coupling in D is the minibatch standard deviation (model_probe_tune.py:749-756):
    a call on B samples splits them into ``group = min(B, 25)`` members x ``M = B / group`` groups, sample ``j*M + m``
    being member j of group m.  Stacking the two batches as [fake[jM:(j+1)M], real[jM:(j+1)M] for j in range(group)]
    and keeping ``group`` makes the 2M groups of the joint call exactly the M fake groups and the M real groups."""
    b = fake.shape[0]
    if real.shape[0] != b:
        return d(fake)[0], d(real)[0]
    group = min(b, d.stddev_group)
    if b % group != 0:
        return d(fake)[0], d(real)[0]
    m = b // group
    x = torch.stack([fake.reshape(group, m, *fake.shape[1:]), real.reshape(group, m, *real.shape[1:])], dim=1)
    pred, _ = d(x.reshape(2 * b, *fake.shape[1:]), stddev_group=group)
    pred = pred.reshape(group, 2, m, *pred.shape[1:])
    return pred[:, 0].reshape(b, *pred.shape[3:]), pred[:, 1].reshape(b, *pred.shape[3:])


class RickAdapter:
    """Owns the four networks, the two Adam optimisers, the Fisher accumulators and the filter masks."""

    def __init__(self, cfg: AdaptConfig, generator, discriminator, g_ema, d_ema, fused_adam: bool = True,
                 fused_generator: bool = False, fused_optim: Optional[bool] = None):
        """``fused_generator``: produce the D step's fake batch (a no-grad generator call) with the tcgen05 executor
        (rick_b200.fused.FusedGenerator) instead of the differentiable module path.
        ``fused_optim``: one rick_adam_mask_ema launch per optimiser step (masks + Adam + EMA, rick_b200.optim) instead
        of rick_mask_apply + torch.optim.Adam + the foreach EMA (default: on for CUDA parameters)."""
        self.cfg = cfg
        self.g, self.d, self.g_ema, self.d_ema = generator, discriminator, g_ema, d_ema
        self.device = next(generator.parameters()).device
        g_ratio = cfg.g_reg_every / (cfg.g_reg_every + 1)
        d_ratio = cfg.d_reg_every / (cfg.d_reg_every + 1)
        self.g_named = dict(generator.named_parameters())
        self.d_named = dict(discriminator.named_parameters())
        # trainable subsets, train:908-931
        self.g_train = [p for n, p in self.g_named.items() if "convs" in n]
        self.d_train = [p for n, p in self.d_named.items()
                        if ("convs" in n and "convs.0" not in n) or "final" in n]
        # DDP exchange step (world_size > 1).  The generator's modulation layers hang off one autograd node (their
        # gradients land together when backward ends): last bucket.  D's equalised-lr scaling runs as one node per
        # bucket, so that a bucket's gradients reach the leaves -- and its all-reduce starts -- while backward goes on.
        g_late = [p for n, p in self.g_named.items() if "convs" in n and "modulation" in n]
        nbk = int(os.environ.get("RICK_DDP_BUCKETS", "4"))
        self._gsync = {"g": rdist.GradSync(self.g_train, n_buckets=nbk, late=g_late),
                       "d": rdist.GradSync(self.d_train, n_buckets=nbk)}
        if self.device.type == "cuda":
            from . import stylegan2 as _sg
            _sg.declare_prescale_groups(generator, self.g_train)          # trainable subset | rest
            _sg.declare_prescale_groups(discriminator, *(self._gsync["d"].buckets if rdist.world_size() > 1
                                                         else [self.d_train]))
            _sg.declare_prescale_groups(g_ema, list(g_ema.parameters()))    # the Fisher round differentiates all of them
            _sg.declare_prescale_groups(d_ema, list(d_ema.parameters()))
        self.acc_g = rick.FisherAccumulator(g_ema.named_parameters())
        self.acc_d = rick.FisherAccumulator(d_ema.named_parameters())
        self.masks_g = rick.FilterMasks(rick.generator_layers(dict(g_ema.named_parameters())), self.device)
        self.masks_d = rick.FilterMasks(rick.discriminator_layers(dict(d_ema.named_parameters())), self.device)
        self.mean_path_length = torch.zeros((), device=self.device)
        self.ema_decay = 0.5 ** (32 / (10 * 1000))
        self.fused_optim = (self.device.type == "cuda") if fused_optim is None else bool(fused_optim)
        g_hyper = dict(lr=cfg.lr * g_ratio, betas=(0 ** g_ratio, 0.99 ** g_ratio))
        d_hyper = dict(lr=cfg.lr * d_ratio, betas=(0 ** d_ratio, 0.99 ** d_ratio))
        if self.fused_optim:
            from .optim import FusedMaskedAdam
            self.g_optim = FusedMaskedAdam(self.g_named, self.g_train, masks=self.masks_g, ema_decay=self.ema_decay,
                                           ema_named=dict(g_ema.named_parameters()), **g_hyper)
            self.d_optim = FusedMaskedAdam(self.d_named, self.d_train, masks=self.masks_d, ema_decay=self.ema_decay,
                                           ema_named=dict(d_ema.named_parameters()), **d_hyper)
        else:
            kw = dict(fused=True) if fused_adam and self.device.type == "cuda" else {}
            self.g_optim = optim.Adam(self.g_train, **g_hyper, **kw)
            self.d_optim = optim.Adam(self.d_train, **d_hyper, **kw)
        self._ema_pairs = None
        self.fg = None
        if fused_generator:
            from .fused import FusedGenerator
            if FusedGenerator.supports(generator):          # e.g. not the 64 / 32-channel layers of 512 / 1024 px
                self.fg = FusedGenerator(generator)

    # ------------------------------------------------------------------------------------------ Fisher round
    def fisher_round(self, latents: torch.Tensor, reals: torch.Tensor, layer_noise: Optional[list] = None):
        """train:214-393.  ``latents`` (num_fisher_img, latent) are the reference's ``_noise/000j.pt`` rows,
        ``reals`` (num_fisher_img, 3, H, W) the first image of each loader batch."""
        mine = self._fisher_begin(latents.shape[0])
        for j in mine:
            self._fisher_image(latents[j:j + 1], reals[j:j + 1], None if layer_noise is None else layer_noise[j])
        self._fisher_end()

    def _fisher_begin(self, n_images: int):
        for m in (self.g_ema, self.d_ema):
            for p in m.parameters():
                p.requires_grad_(True)
            m.eval()
        self.acc_g.zero()
        self.acc_d.zero()
        world = rdist.world_size()
        rank = torch.distributed.get_rank() if world > 1 else 0
        return rdist.shard_range(n_images, rank, world)            # Fisher images are sharded over ranks

    def _fisher_image(self, latent: torch.Tensor, real: torch.Tensor, noise=None):
        """grad**2 of one (latent, real image) pair into the accumulators (train:225-263).  The two discriminator calls
        of the reference share one pass (d_pair; at batch 1 each sample is its own minibatch-stddev group)."""
        fake, _ = self.g_ema([latent], noise=noise)
        fake_pred, real_pred = d_pair(self.d_ema, fake, real)
        g_loss = g_nonsaturating_loss(fake_pred)
        d_loss = d_logistic_loss(real_pred, fake_pred)
        # the generator loss reaches G's parameters THROUGH D: D's own weight gradients are not part of this pass
        from . import conv as _conv
        d_ptrs = set(getattr(self.d_ema, "_scaled_weight_ptrs", ())) | {p.data_ptr() for p in self.d_ema.parameters()}
        with _conv.skip_weight_grads(d_ptrs):
            g_grads = autograd.grad(g_loss, list(self.g_ema.parameters()), retain_graph=True)
        d_grads = autograd.grad(d_loss, list(self.d_ema.parameters()))
        self.acc_g.add(g_grads, first=False)
        self.acc_d.add(d_grads, first=False)

    def _fisher_end(self):
        cfg = self.cfg
        if rdist.world_size() > 1:                     # one exchange step: SUM of the grad^2 accumulators
            for acc in (self.acc_g, self.acc_d):
                rdist.allreduce_sum_(acc.acc)
        div = cfg.num_fisher_img * cfg.batch          # the reference's divisor (train:267), kept as is
        self.acc_g.average(div)
        self.acc_d.average(div)
        self.masks_g.update(self.acc_g.as_dict(), cfg.fisher_quantile, cfg.prune_quantile)
        self.masks_d.update(self.acc_d.as_dict(), cfg.fisher_quantile, cfg.prune_quantile)

    # ------------------------------------------------------------------------------------------ one iteration
    @torch.no_grad()
    def _latents(self, z: List[torch.Tensor], inject: Optional[int]) -> torch.Tensor:
        """(B, n_latent, D) latent of ``Generator.forward(z, inject_index=inject)`` (model_probe_tune.py:532-560)."""
        n = self.g.n_latent
        w = [self.g.map_latent(zz) for zz in z]
        if len(w) < 2:
            return w[0].unsqueeze(1).repeat(1, n, 1)
        return torch.cat([w[0].unsqueeze(1).repeat(1, inject, 1), w[1].unsqueeze(1).repeat(1, n - inject, 1)], 1)

    def _gate_warmup(self, i: int):
        self._warm = i < self.cfg.warmup_iter

    def _grad_mode(self, net: str):
        """``requires_grad(generator, net == "g"); requires_grad(discriminator, net == "d")`` (train:396-398, 497-499),
        narrowed to the parameters the optimisers own (train:908-931): a network's other parameters never need a
        gradient inside the loop, and switching the OTHER network off keeps the custom convolution / bias-act Functions
        from computing weight gradients nobody reads (the G step back-propagates through D for its data gradients only).
        During warm-up only D's ``final_*`` parameters train (train:202-211)."""
        warm = getattr(self, "_warm", False)
        train_g, train_d = {id(p) for p in self.g_train}, {id(p) for p in self.d_train}
        for n, p in self.g_named.items():
            p.requires_grad_(net == "g" and id(p) in train_g)
        for n, p in self.d_named.items():
            p.requires_grad_(net == "d" and id(p) in train_d and ((not warm) or "final" in n))

    def step(self, i: int, real_img: torch.Tensor, draws: DrawStream, explicit_layer_noise: bool = False
             ) -> Dict[str, torch.Tensor]:
        """One iteration of train:396-589 + EMA (697-698).  Returns loss tensors (no host sync)."""
        cfg = self.cfg
        after_warmup = i >= cfg.warmup_iter
        self._gate_warmup(i)
        noise_of = (lambda b: draws.layer_noise(b, cfg.size)) if explicit_layer_noise else (lambda b: None)
        out: Dict[str, torch.Tensor] = {}

        # ---- D step (train:396-438) ----
        z = draws.mixing_latents(cfg.batch, cfg.latent, cfg.mixing)
        inject = draws.randint(1, self.g.n_latent - 1) if len(z) == 2 else None
        with torch.no_grad():
            lat = self._latents(z, inject)
            gen = self.fg if self.fg is not None else self.g     # fg reads G's parameters in place: always current
            fake_img, _ = gen([lat], input_is_latent=True, noise=noise_of(cfg.batch))
        self._grad_mode("d")
        fake_pred, real_pred = d_pair(self.d, fake_img, real_img)
        d_loss = d_logistic_loss(real_pred, fake_pred)
        out["d"], out["real_score"], out["fake_score"] = d_loss.detach(), real_pred.mean().detach(), fake_pred.mean().detach()
        self.d.zero_grad(set_to_none=True)
        # only the trainable subset needs gradients (train:921-931 hands exactly these to the optimiser); asking for
        # them alone skips the from-RGB weight gradient and everything upstream of it
        self._backward("d", d_loss, [p for p in self.d_train if p.requires_grad])
        r1_iter = i % cfg.d_reg_every == 0
        path_iter = i % cfg.g_reg_every == 0 and after_warmup
        # with the fused optimiser the EMA rides on the LAST update of each network in this iteration (the weights do
        # not move again before train:697-698 would read them)
        self._optim_step("d", after_warmup, ema=not r1_iter)

        # ---- R1 (train:462-493) ----
        if r1_iter:
            real_r = real_img.detach().requires_grad_(True)
            real_pred, _ = self.d(real_r)
            real_pred = real_pred.view(real_r.size(0), -1).mean(dim=1).unsqueeze(1)
            r1_loss = d_r1_loss(real_pred, real_r)
            self.d.zero_grad(set_to_none=True)
            self._backward("d", cfg.r1 / 2 * r1_loss * cfg.d_reg_every + 0 * real_pred[0],
                           [p for p in self.d_train if p.requires_grad])
            self._optim_step("d", after_warmup, ema=True)
            out["r1"] = r1_loss.detach()

        # ---- G step (train:500-540) ----
        z = draws.mixing_latents(cfg.batch, cfg.latent, cfg.mixing)
        inject = draws.randint(1, self.g.n_latent - 1) if len(z) == 2 else None
        self._grad_mode("g")
        if after_warmup:
            fake_img, _ = self.g([self._latents(z, inject)], input_is_latent=True, noise=noise_of(cfg.batch))
            fake_pred, _ = self.d(fake_img)
            g_loss = g_nonsaturating_loss(fake_pred)
            self.g.zero_grad(set_to_none=True)
            self._backward("g", g_loss)
            self._optim_step("g", True, ema=not path_iter)
        else:
            with torch.no_grad():                       # warm-up: the loss is only logged (train:518-519)
                fake_img, _ = self.g([self._latents(z, inject)], input_is_latent=True, noise=noise_of(cfg.batch))
                g_loss = g_nonsaturating_loss(self.d(fake_img)[0])
        out["g"] = g_loss.detach()

        # ---- path-length regularisation (train:546-589) ----
        if path_iter:
            pb = max(1, cfg.batch // cfg.path_batch_shrink)
            z = draws.mixing_latents(pb, cfg.latent, cfg.mixing)
            inject = draws.randint(1, self.g.n_latent - 1) if len(z) == 2 else None
            lat = self._latents(z, inject).requires_grad_(True)
            fake_img, latents = self.g([lat], input_is_latent=True, return_latents=True, noise=noise_of(pb))
            path_loss, self.mean_path_length, path_lengths = g_path_regularize(
                fake_img, latents, self.mean_path_length, draws.normal(*fake_img.shape))
            self.g.zero_grad(set_to_none=True)
            weighted = cfg.path_regularize * cfg.g_reg_every * path_loss
            if cfg.path_batch_shrink:
                weighted = weighted + 0 * fake_img[0, 0, 0, 0]
            self._backward("g", weighted)
            self._optim_step("g", True, ema=True)
            out["path"], out["path_length"] = path_loss.detach(), path_lengths.mean().detach()

        # ---- EMA (train:697-698) ----
        if not self.fused_optim:
            self._ema()
        elif not after_warmup:
            self.g_optim.ema_only()                     # G took no optimiser step during warm-up
        return out

    def _optim_step(self, net: str, masked: bool, ema: bool = False, force_masks: bool = False):
        """Masks + optimiser step of network ``net`` ("g" / "d"); with the fused optimiser also its EMA when ``ema``."""
        opt = self.g_optim if net == "g" else self.d_optim
        if self.fused_optim:
            opt.step(apply_masks=masked, ema=ema)
            return
        if masked:
            (self.masks_g if net == "g" else self.masks_d).apply(self.g_named if net == "g" else self.d_named,
                                                                 force=force_masks)
        opt.step()

    def _backward(self, net: str, loss: torch.Tensor, inputs=None):
        """Backward pass of network ``net`` ("g" / "d") into its trainable parameters, with the DDP exchange step
        (average of the gradients over ranks; nothing in a single process) overlapped: rick_b200.dist.GradSync launches
        each bucket's all-reduce as soon as backward has produced it and the call returns once all have landed."""
        sync = self._gsync[net]
        sync.begin()
        autograd.backward(loss, inputs=(self.g_train if net == "g" else self.d_train) if inputs is None else inputs)
        sync.finish()

    @torch.no_grad()
    def _ema(self):
        if self._ema_pairs is None:
            ge, gs = dict(self.g_ema.named_parameters()), self.g_named
            de, ds = dict(self.d_ema.named_parameters()), self.d_named
            self._ema_pairs = ([ge[k] for k in ge] + [de[k] for k in de], [gs[k] for k in ge] + [ds[k] for k in de])
        dst, src = self._ema_pairs
        torch._foreach_mul_(dst, self.ema_decay)
        torch._foreach_add_(dst, src, alpha=1 - self.ema_decay)


@torch.no_grad()
def generate_samples(generator, n_samples: int, batch: int, latent: int = 512, rank: int = 0, world: int = 1,
                     seed: int = 0, to_host: bool = False, fused: bool = False):
    """Batch-sharded sample generation (gan_training/eval.py:31-46): rank r takes batches
    r, r+W, ...; ``z`` for batch k comes from ``manual_seed(seed + k)`` so any sharding produces the same images.
    Yields (batch_index, images)."""
    device = next(generator.parameters()).device
    n_batches = (n_samples + batch - 1) // batch
    gen = torch.Generator(device=device)
    if fused:                                   # tcgen05 executor (rick_b200.fused), same images up to TF32 rounding
        from .fused import FusedGenerator
        if FusedGenerator.supports(generator):
            generator = FusedGenerator(generator)
    for k in range(rank, n_batches, world):
        gen.manual_seed(seed + k)
        z = torch.randn(batch, latent, generator=gen, device=device)
        img, _ = generator([z])
        yield k, (img.cpu() if to_host else img)

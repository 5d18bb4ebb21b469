"""Fused optimiser step of the adaptation loop: filter masks + Adam + EMA in one launch per network
(rick_adam_mask_ema, rick_b200/csrc/optim.cu).

Stands where the reference has, per optimiser step, the index_put mask application (train_dynamic_update_prune.py:
427-437, 482-492, 521-539, 566-585), ``optim.Adam(...).step()`` (train:916-925) and -- once per iteration --
``accumulate(g_ema, g_module, accum)`` (train:68-73, 697-698).  ``state_dict`` / ``load_state_dict`` use the layout of
``torch.optim.Adam`` so the reference's ``g_optim`` / ``d_optim`` checkpoint entries (train:647-659) load unchanged.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from . import _lib
from .rick import filter_major


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _same_layout(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Same element order in memory (strides of size-1 dimensions carry no information)."""
    return a.shape == b.shape and all(sa == sb for n, sa, sb in zip(a.shape, a.stride(), b.stride()) if n > 1)


class FusedMaskedAdam:
    """Adam over ``train`` (a subset of ``named``), with the filter masks of ``masks`` applied to the gradient / weight
    inside the update and, on request, the EMA of EVERY parameter of ``named`` into ``ema_named`` in the same launch.

    Same hyper-parameter meaning as ``torch.optim.Adam(params, lr, betas, eps)`` without weight decay / amsgrad (the
    reference uses neither).  Step counts are per parameter, as in torch.optim.Adam (a parameter that is gated off
    during warm-up starts its bias correction when it first receives a gradient), and live on the device so a step can
    be recorded into a CUDA graph."""

    def __init__(self, named: Dict[str, torch.nn.Parameter], train: Iterable[torch.nn.Parameter], lr: float,
                 betas=(0.9, 0.999), eps: float = 1e-8, masks=None, ema_named: Optional[Dict[str, torch.nn.Parameter]] = None,
                 ema_decay: float = 0.0):
        self.named = dict(named)
        self.train: List[torch.nn.Parameter] = list(train)
        train_ids = {id(p) for p in self.train}
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.masks = masks
        self.ema_named = dict(ema_named) if ema_named is not None else None
        self.ema_decay = float(ema_decay)
        first = self.train[0]
        if first.device.type != "cuda":
            raise RuntimeError("FusedMaskedAdam runs on CUDA parameters only (there is no CPU fallback)")
        self.device = first.device
        for p in self.named.values():
            if p.dtype != torch.float32:
                raise RuntimeError("FusedMaskedAdam: float32 parameters only")
        self.steps = torch.zeros(len(self.train), dtype=torch.float32, device=self.device)
        self._slot = {id(p): i for i, p in enumerate(self.train)}
        self._inc_cache: Dict[tuple, torch.Tensor] = {}
        self.exp_avg = {id(p): torch.zeros_like(p, memory_format=torch.preserve_format) for p in self.train}
        self.exp_avg_sq = {id(p): torch.zeros_like(p, memory_format=torch.preserve_format) for p in self.train}
        self._is_train = {n: id(p) in train_ids for n, p in self.named.items()}
        # torch.optim.Adam-compatible view for checkpoints
        self.param_groups = [dict(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=0, amsgrad=False,
                                  params=list(range(len(self.train))))]

    # ------------------------------------------------------------------------------------------ the step
    def ema_only(self):
        """accumulate() alone (train:697-698) for an iteration in which this network took no optimiser step."""
        self.step(apply_masks=False, ema=True, update=False)

    @torch.no_grad()
    def step(self, apply_masks: bool = True, ema: bool = False, update: bool = True):
        """One optimiser step.  ``apply_masks``: use the current freeze / prune masks (no-op before the first Fisher
        round: they are all zero).  ``ema``: also fold every parameter into its EMA copy (last step of an iteration).
        ``update=False``: no Adam update at all (EMA only)."""
        state = self.masks.state if (self.masks is not None and apply_masks) else {}
        zero = self.masks.zero if (self.masks is not None and apply_masks) else {}
        P, G, M, V, E, S, Z, T, rows, inner = [], [], [], [], [], [], [], [], [], []
        stepped = [False] * len(self.train)
        step_base, step_size = self.steps.data_ptr(), self.steps.element_size()
        for name, p in self.named.items():
            g = p.grad if (update and self._is_train[name]) else None
            e = self.ema_named[name] if (ema and self.ema_named is not None) else None
            if g is None and e is None:
                continue
            st = state.get(name) if g is not None else None
            if not filter_major(p, st.numel() if st is not None else 1):
                raise RuntimeError(f"FusedMaskedAdam: {name} is not stored filter-major")
            for other, what in ((g, "gradient"), (e, "EMA copy")):
                if other is not None and not _same_layout(p, other):
                    raise RuntimeError(f"FusedMaskedAdam: {what} of {name} does not share the parameter's memory layout")
            r = st.numel() if st is not None else 1
            if g is not None:
                stepped[self._slot[id(p)]] = True
            P.append(p.data_ptr())
            G.append(g.data_ptr() if g is not None else None)
            M.append(self.exp_avg[id(p)].data_ptr() if g is not None else None)
            V.append(self.exp_avg_sq[id(p)].data_ptr() if g is not None else None)
            E.append(e.data_ptr() if e is not None else None)
            S.append(st.data_ptr() if st is not None else None)
            Z.append(zero[name].data_ptr() if st is not None else None)
            T.append(step_base + step_size * self._slot[id(p)] if g is not None else None)
            rows.append(r)
            inner.append(p.numel() // r)
        if any(stepped):                                     # step += 1 for exactly the tensors updated now
            key = tuple(stepped)
            inc = self._inc_cache.get(key)
            if inc is None:
                inc = self._inc_cache[key] = torch.tensor([float(b) for b in stepped], device=self.device)
            self.steps += inc
        if not P:
            return
        with torch.cuda.device(self.device):
            rc = _lib.lib().rick_adam_mask_ema(
                _lib.ptr_table(P), _lib.ptr_table(G), _lib.ptr_table(M), _lib.ptr_table(V), _lib.ptr_table(E),
                _lib.ptr_table(S), _lib.ptr_table(Z), _lib.ptr_table(T), _lib.i64_table(rows), _lib.i64_table(inner),
                len(P), self.lr, self.betas[0], self.betas[1], self.eps, self.ema_decay, _stream())
        _lib.check(rc, "rick_adam_mask_ema")

    def zero_grad(self, set_to_none: bool = True):
        for p in self.train:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    # ------------------------------------------------------------------------------------------ checkpoints
    def state_dict(self) -> dict:
        """``torch.optim.Adam.state_dict()`` layout: {"state": {i: {step, exp_avg, exp_avg_sq}}, "param_groups": [...]}."""
        state = {}
        steps = self.steps.cpu()
        for i, p in enumerate(self.train):
            if float(steps[i]) > 0:
                state[i] = {"step": steps[i].clone(), "exp_avg": self.exp_avg[id(p)].clone(),
                            "exp_avg_sq": self.exp_avg_sq[id(p)].clone()}
        return {"state": state, "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd: dict):
        groups = sd["param_groups"]
        n = sum(len(g["params"]) for g in groups)
        if n != len(self.train):
            raise ValueError(f"optimizer state has {n} parameters, this optimiser {len(self.train)}")
        g0 = groups[0]
        self.lr, self.betas, self.eps = float(g0["lr"]), tuple(float(b) for b in g0["betas"]), float(g0["eps"])
        self.param_groups = [dict(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=0, amsgrad=False,
                                  params=list(range(len(self.train))))]
        steps = torch.zeros(len(self.train), dtype=torch.float32)
        for i, p in enumerate(self.train):
            st = sd["state"].get(i)
            if st is None:
                self.exp_avg[id(p)].zero_(), self.exp_avg_sq[id(p)].zero_()
                continue
            self.exp_avg[id(p)].copy_(st["exp_avg"].reshape(p.shape))
            self.exp_avg_sq[id(p)].copy_(st["exp_avg_sq"].reshape(p.shape))
            steps[i] = float(st["step"])
        self.steps.copy_(steps)

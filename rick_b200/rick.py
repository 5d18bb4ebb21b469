"""RICK's Fisher -> quantile -> freeze/prune-mask step, device resident.

Host-side mirror of the in-line block of the reference's ``train()``
(train_dynamic_update_prune.py:214-393) and of the per-iteration mask application (427-437, 482-492, 521-539,
566-585).  The reference moves every squared gradient to the host (110 + 38 tensors per Fisher image), reduces
with NumPy, and re-uploads index arrays for ~200 ``index_put`` calls per iteration.  Here the Fisher tensors, the
per-filter FIM vectors, the thresholds and the masks never leave the GPU:

    FisherAccumulator.add()      rick_fisher_accum   (one multi-tensor launch per image and model)
    FisherAccumulator.average()  rick_fisher_divide
    FilterMasks.update()         rick_filter_fim (NumPy-exact pairwise float32 means) -> rick_percentile (radix select
                                 + float64 lerp == np.percentile 'linear') -> rick_decide (freeze / ft / prune bits and
                                 the cumulative prune union)
    FilterMasks.apply()          rick_mask_apply     (one launch per model per optimiser step)

Given the same Fisher tensors the resulting sets are bit-identical to the reference's NumPy index sets
(tests/test_rick_gpu.py, tests/golden/rick_masks_golden.npz).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------ accumulation

def filter_major(t: torch.Tensor, rows: int = 1) -> bool:
    """True when ``t`` is stored densely (a permutation of a contiguous block) with its filter dimension outermost, so
    that filter r occupies elements [r * inner, (r + 1) * inner) of its storage -- what the per-filter kernels assume.
    Holds for contiguous tensors, channels-last 4-D weights and the (1, Cout, Cin, k, k) modulated-conv weights kept
    as [Cout][k][k][Cin].  The filter dimension is dim 1 for 5-D tensors (train:526-537), else dim 0."""
    sizes, strides = list(t.shape), list(t.stride())
    dims = sorted((d for d in range(t.dim()) if sizes[d] > 1), key=lambda d: -strides[d])
    expect = t.numel()
    for d in dims:                                   # dense: each stride is the product of the sizes inside it
        expect //= sizes[d]
        if strides[d] != expect:
            return False
    if rows <= 1 or not dims:
        return True
    fd = 1 if t.dim() == 5 else 0
    return sizes[fd] == rows and dims[0] == fd


class FisherAccumulator:
    """Sum of squared gradients per parameter (train:252-263), kept on the device in float32."""

    def __init__(self, named_params: Iterable[Tuple[str, torch.Tensor]]):
        self.names: List[str] = []
        self.acc: List[torch.Tensor] = []
        for n, p in named_params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("FisherAccumulator: parameters must be float32 CUDA tensors")
            self.names.append(n)
            self.acc.append(torch.empty_like(p, memory_format=torch.contiguous_format))
        self._numel = _lib.i64_table([a.numel() for a in self.acc])
        self._acc_tab = _lib.ptr_table([a.data_ptr() for a in self.acc])
        self.count = 0

    def reset(self):
        self.count = 0

    def zero(self):
        """Explicitly clear the accumulators (for callers that always accumulate, e.g. a captured CUDA graph)."""
        torch._foreach_zero_(self.acc)
        self.count = 0

    def add(self, grads: Sequence[torch.Tensor], first: Optional[bool] = None):
        """acc (+)= grad ** 2 for every parameter; ``grads`` is what ``autograd.grad(loss, params)`` returned.
        ``first`` overrides "overwrite on the first call after reset()" (False = always accumulate)."""
        if len(grads) != len(self.acc):
            raise RuntimeError("FisherAccumulator.add: gradient list does not match the parameter list")
        keep = [g.detach().contiguous() for g in grads]       # alive until the launch is enqueued
        for g, a in zip(keep, self.acc):
            if g.shape != a.shape or g.dtype != torch.float32 or not g.is_cuda:
                raise RuntimeError("FisherAccumulator.add: gradient shape / dtype / device mismatch")
        with torch.cuda.device(self.acc[0].device):
            st = _lib.lib().rick_fisher_accum(self._acc_tab, _lib.ptr_table([g.data_ptr() for g in keep]), self._numel,
                                              len(keep), int(self.count == 0 if first is None else first), _stream())
        _lib.check(st, "rick_fisher_accum")
        self.count += 1

    def average(self, divisor: float):
        """``fisher /= num_fisher_img * batch`` (train:266-269)."""
        with torch.cuda.device(self.acc[0].device):
            st = _lib.lib().rick_fisher_divide(self._acc_tab, self._numel, len(self.acc), float(divisor), _stream())
        _lib.check(st, "rick_fisher_divide")

    def as_dict(self) -> Dict[str, torch.Tensor]:
        return dict(zip(self.names, self.acc))


# ------------------------------------------------------------------------------------------------ layer tables

@dataclass
class FilterLayer:
    weight: str                    # Fisher key whose rows are the filters
    bias: Optional[str]            # Fisher key averaged in, (mean(w) + b) / 2, or None
    rows: int
    length: int                    # elements per filter row
    group: str                     # layers pooled into one percentile
    closed_low: bool = False       # D skip comparisons (train:382-384)
    targets: List[str] = field(default_factory=list)   # parameters masked by this layer's decision
    offset: int = 0                # position inside the group's pooled vector


def generator_layers(g_state: Dict[str, torch.Tensor], n_convs: Optional[int] = None) -> List[FilterLayer]:
    """train:281-299 -- 'conv' group: convs.i.conv.weight rows; 'fc' group: modulation weight rows (+ bias)."""
    if n_convs is None:
        n_convs = len({k.split(".")[1] for k in g_state if k.startswith("convs.") and k.endswith(".conv.weight")})
    layers: List[FilterLayer] = []
    for i in range(n_convs):
        w = g_state[f"convs.{i}.conv.weight"]
        layers.append(FilterLayer(f"convs.{i}.conv.weight", None, w.shape[1], w[0, 0].numel(), "conv",
                                  targets=[f"convs.{i}.conv.weight"]))
    for i in range(n_convs):
        w = g_state[f"convs.{i}.conv.modulation.weight"]
        layers.append(FilterLayer(f"convs.{i}.conv.modulation.weight", f"convs.{i}.conv.modulation.bias", w.shape[0],
                                  w.shape[1], "fc", targets=[f"convs.{i}.conv.modulation.weight",
                                                             f"convs.{i}.conv.modulation.bias"]))
    return layers


def discriminator_layers(d_state: Dict[str, torch.Tensor], blocks: Optional[Sequence[int]] = None) -> List[FilterLayer]:
    """train:336-351 -- one group; conv1, conv2 (with their FusedLeakyReLU bias), skip, per ResBlock."""
    if blocks is None:
        blocks = sorted({int(k.split(".")[1]) for k in d_state if k.startswith("convs.") and ".conv1.0.weight" in k})
    layers: List[FilterLayer] = []
    for b in blocks:
        for wk, bk in ((f"convs.{b}.conv1.0.weight", f"convs.{b}.conv1.1.bias"),
                       (f"convs.{b}.conv2.1.weight", f"convs.{b}.conv2.2.bias"),
                       (f"convs.{b}.skip.1.weight", None)):
            w = d_state[wk]
            layers.append(FilterLayer(wk, bk, w.shape[0], w[0].numel(), "d", closed_low=bk is None,
                                      targets=[wk] + ([bk] if bk else [])))
    return layers


# ------------------------------------------------------------------------------------------------ masks

class FilterMasks:
    """Freeze / fine-tune / prune decisions of one model, as device byte vectors.

    ``state[key]`` holds per-filter bits (1 freeze, 2 prune, 4 fine-tune); ``zero[key]`` is the cumulative prune
    union (train:386-393).  Keys are parameter names, exactly the keys of the reference's idx_* dicts."""

    def __init__(self, layers: List[FilterLayer], device, numpy1_compare: bool = False):
        """``numpy1_compare``: compare float32 FIMs against the float64 thresholds the way NumPy < 2 does (threshold
        cast to float32 first -- the reference pins NumPy 1.23.1); default is NumPy >= 2 promotion (float64), which is
        what the golden masks of this repository were generated under (NumPy 2.3.5).  The two differ only for a filter
        whose FIM lies within half a float32 ulp of a threshold."""
        self.numpy1_compare = bool(numpy1_compare)
        self.layers = layers
        self.device = torch.device(device)
        self.groups: Dict[str, int] = {}
        for l in layers:
            l.offset = self.groups.get(l.group, 0)
            self.groups[l.group] = l.offset + l.rows
        self.fim = {g: torch.zeros(n, dtype=torch.float32, device=device) for g, n in self.groups.items()}
        self._state = {g: torch.zeros(n, dtype=torch.uint8, device=device) for g, n in self.groups.items()}
        self._zero = {g: torch.zeros(n, dtype=torch.uint8, device=device) for g, n in self.groups.items()}
        self.lines = {g: torch.zeros(2, dtype=torch.float64, device=device) for g in self.groups}
        self.state: Dict[str, torch.Tensor] = {}
        self.zero: Dict[str, torch.Tensor] = {}
        for l in layers:
            for t in l.targets:
                self.state[t] = self._state[l.group][l.offset:l.offset + l.rows]
                self.zero[t] = self._zero[l.group][l.offset:l.offset + l.rows]
        self.rounds = 0
        self._apply_cache = None

    def update(self, fisher: Dict[str, torch.Tensor], fisher_quantile: float, prune_quantile: float):
        """One Fisher round: FIM per filter, pooled percentiles, decisions, cumulative prune union."""
        lib = _lib.lib()
        s = _stream()
        with torch.cuda.device(self.device):
            for l in self.layers:
                w = fisher[l.weight]
                if not w.is_contiguous() or w.dtype != torch.float32:
                    raise RuntimeError(f"FilterMasks.update: Fisher tensor {l.weight} must be contiguous float32")
                b = fisher[l.bias] if l.bias else None
                fim = self.fim[l.group][l.offset:l.offset + l.rows]
                _lib.check(lib.rick_filter_fim(fim.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None,
                                               l.rows, l.length, s), "rick_filter_fim")
            q = (ctypes.c_double * 2)(float(fisher_quantile), float(prune_quantile))
            for g, n in self.groups.items():
                _lib.check(lib.rick_percentile(self.lines[g].data_ptr(), self.fim[g].data_ptr(), n, q, 2, s),
                           "rick_percentile")
            for l in self.layers:
                sl = slice(l.offset, l.offset + l.rows)
                _lib.check(lib.rick_decide(self._state[l.group][sl].data_ptr(), self._zero[l.group][sl].data_ptr(),
                                           self.fim[l.group][sl].data_ptr(), l.rows, self.lines[l.group].data_ptr(),
                                           int(l.closed_low) | (2 if self.numpy1_compare else 0), int(self.rounds == 0), s),
                           "rick_decide")
        self.rounds += 1

    def apply(self, named_params: Dict[str, torch.nn.Parameter], force: bool = False):
        """grad[freeze] = 0; param[zero] = 0; grad[zero] = 0 for every masked parameter -- one launch.
        Before the first Fisher round the masks are all-zero and the call is skipped unless ``force`` (CUDA-graph
        capture wants the launch recorded regardless)."""
        if self.rounds == 0 and not force:
            return
        params, grads, states, zeros, rows, inner = [], [], [], [], [], []
        for name, st in self.state.items():
            p = named_params[name]
            g = p.grad
            r = st.numel()
            for t in (p, g):
                # rows (= filters) must be the outermost, dense blocks of memory
                if t is not None and not filter_major(t, r):
                    raise RuntimeError(f"FilterMasks.apply: {name} (or its gradient) is not stored filter-major")
            params.append(p.data_ptr())
            grads.append(g.data_ptr() if g is not None else None)
            states.append(st.data_ptr())
            zeros.append(self.zero[name].data_ptr())
            rows.append(r)
            inner.append(p.numel() // r)
        with torch.cuda.device(self.device):
            st = _lib.lib().rick_mask_apply(_lib.ptr_table(params), _lib.ptr_table(grads), _lib.ptr_table(states),
                                            _lib.ptr_table(zeros), _lib.i64_table(rows), _lib.i64_table(inner),
                                            len(params), _stream())
        _lib.check(st, "rick_mask_apply")

    # ---- host views for tests / logging (synchronise) ----
    def index_sets(self):
        """(freeze, ft, prune, zero) dicts of int64 NumPy index arrays, the reference's idx_* / zero_filter_idx_*."""
        out = ({}, {}, {}, {})
        for name, st in self.state.items():
            s = st.cpu().numpy()
            out[0][name] = (s & 1).nonzero()[0]
            out[1][name] = (s & 4).nonzero()[0]
            out[2][name] = (s & 2).nonzero()[0]
            out[3][name] = self.zero[name].cpu().numpy().nonzero()[0]
        return out

"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of RICK's Fisher -> quantile -> mask step.

The reference keeps this logic in-line in ``train()``
(train_dynamic_update_prune.py:214-393 for the Fisher round, 427-437 / 482-492 /
521-539 / 566-585 for the per-iteration mask application).  It is restated here
as functions over ``{key: np.ndarray}`` Fisher dicts.  ``oracle/make_golden.py``
executes the reference's own lines 277-393 on the same seeded Fisher dicts and
stores the resulting index sets; ``tests/test_rick_oracle.py`` checks this file
against them, so the restatement is pinned to the reference's code, not to
my reading of it.

All arithmetic intentionally stays in NumPy with the reference's dtypes:
float32 Fisher arrays, float32 per-filter means (``ndarray.mean``: pairwise
float32 summation, then a float32 divide), a float64 pooled vector
(``np.concatenate(([], m))``), ``np.percentile`` with the default 'linear'
method, and float32-vs-float64 comparisons promoted to float64 (NumPy >= 2).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

Fisher = Dict[str, np.ndarray]
IndexSets = Dict[str, np.ndarray]


# ------------------------------------------------------------------ accumulation (train:252-269)

def fisher_accumulate(total: Fisher, one_image: Fisher) -> Fisher:
    """First image assigns, later images ``+=`` in float32 (train:252-263)."""
    for k, v in one_image.items():
        v = np.asarray(v, dtype=np.float32)
        if k not in total:
            total[k] = v.copy()
        else:
            total[k] += v
    return total


def fisher_average(total: Fisher, num_fisher_img: int, batch: int) -> Fisher:
    """In-place divide by ``num_fisher_img * batch`` (train:266-269)."""
    for k in total:
        total[k] /= (num_fisher_img * batch)
    return total


# ------------------------------------------------------------------ per-filter FIM (train:281-299, 336-351)

def g_conv_keys(n_convs: int = 12) -> List[str]:
    return [f"convs.{i}.conv.weight" for i in range(n_convs)]


def g_fc_keys(n_convs: int = 12) -> List[Tuple[str, str]]:
    return [(f"convs.{i}.conv.modulation.weight", f"convs.{i}.conv.modulation.bias") for i in range(n_convs)]


def d_layer_keys(blocks=range(1, 7)) -> List[Tuple[str, str | None]]:
    """(weight key, bias key or None) in the order the reference pools them (train:336-351)."""
    out: List[Tuple[str, str | None]] = []
    for b in blocks:
        out.append((f"convs.{b}.conv1.0.weight", f"convs.{b}.conv1.1.bias"))
        out.append((f"convs.{b}.conv2.1.weight", f"convs.{b}.conv2.2.bias"))
        out.append((f"convs.{b}.skip.1.weight", None))
    return out


def fim_g_conv(f: Fisher, key: str) -> np.ndarray:
    return f[key].mean(axis=(0, 2, 3, 4))                       # train:282, 309  -> (Cout,)


def fim_g_fc(f: Fisher, wkey: str, bkey: str) -> np.ndarray:
    return (f[wkey].mean(axis=1) + f[bkey]) / 2                 # train:291-293, 318-320 -> (Cin,)


def fim_d(f: Fisher, wkey: str, bkey: str | None) -> np.ndarray:
    m = f[wkey].mean(axis=(1, 2, 3))                            # train:339, 349
    return m if bkey is None else (m + f[bkey]) / 2             # train:340-341


def pool(vectors: List[np.ndarray]) -> np.ndarray:
    """``np.concatenate(([], v), axis=None)`` repeatedly: a float64 vector (train:283, 294, 342, 350)."""
    g = np.asarray([], dtype=np.float64)
    for v in vectors:
        g = np.concatenate((g, v), axis=None)
    return g


# ------------------------------------------------------------------ decisions (train:302-384)

# NumPy < 2 (the reference pins 1.23.1) compares a float32 array with a float64 scalar after casting the scalar to
# float32 (value-based casting); NumPy >= 2 promotes the array to float64.  Set to True to restate the former.
NUMPY1_COMPARE = False


def _three_way(fim: np.ndarray, cut, prune, closed_low: bool):
    """freeze / fine-tune / prune index sets.  ``closed_low`` is the D-skip variant (train:382-384)."""
    if NUMPY1_COMPARE and fim.dtype == np.float32:
        cut, prune = np.float32(cut), np.float32(prune)
    else:
        fim = fim.astype(np.float64, copy=False)
    freeze = np.where(fim > cut)[0]
    if closed_low:
        ft = np.where((fim >= prune) & (fim <= cut))[0]
        pr = np.where(fim < prune)[0]
    else:
        ft = np.where((fim > prune) & (fim <= cut))[0]
        pr = np.where(fim <= prune)[0]
    return freeze, ft, pr


def decide_g(f: Fisher, fisher_quantile: float, prune_quantile: float, n_convs: int = 12):
    conv_fims = {k: fim_g_conv(f, k) for k in g_conv_keys(n_convs)}
    fc_fims = {w: fim_g_fc(f, w, b) for w, b in g_fc_keys(n_convs)}
    pooled_conv = pool(list(conv_fims.values()))
    pooled_fc = pool(list(fc_fims.values()))
    lines = {
        "cut_conv": np.percentile(pooled_conv, q=fisher_quantile),
        "prune_conv": np.percentile(pooled_conv, q=prune_quantile),
        "cut_fc": np.percentile(pooled_fc, q=fisher_quantile),
        "prune_fc": np.percentile(pooled_fc, q=prune_quantile),
    }
    freeze: IndexSets = {}
    ft: IndexSets = {}
    prune: IndexSets = {}
    for k, fim in conv_fims.items():
        freeze[k], ft[k], prune[k] = _three_way(fim, lines["cut_conv"], lines["prune_conv"], False)
    for (w, b) in g_fc_keys(n_convs):
        fr, t, pr = _three_way(fc_fims[w], lines["cut_fc"], lines["prune_fc"], False)
        for k in (w, b):
            freeze[k], ft[k], prune[k] = fr, t, pr
    return freeze, ft, prune, lines


def decide_d(f: Fisher, fisher_quantile: float, prune_quantile: float, blocks=range(1, 7)):
    layers = d_layer_keys(blocks)
    fims = {w: fim_d(f, w, b) for w, b in layers}
    pooled = pool([fims[w] for w, _ in layers])
    lines = {"cut": np.percentile(pooled, q=fisher_quantile), "prune": np.percentile(pooled, q=prune_quantile)}
    freeze: IndexSets = {}
    ft: IndexSets = {}
    prune: IndexSets = {}
    for w, b in layers:
        fr, t, pr = _three_way(fims[w], lines["cut"], lines["prune"], closed_low=(b is None))
        freeze[w], ft[w], prune[w] = fr, t, pr
        if b is not None:
            freeze[b], ft[b], prune[b] = fr, t, pr
    return freeze, ft, prune, lines


def zero_idx_merge(old: IndexSets, new: IndexSets) -> IndexSets:
    """Cumulative union of prune sets (train:138-144)."""
    return {k: np.unique(np.concatenate((old[k], new[k]))) for k in old}


# ------------------------------------------------------------------ mask application (train:427-437, 521-539)

def apply_masks_numpy(params: Dict[str, np.ndarray], grads: Dict[str, np.ndarray], freeze: IndexSets,
                      zero: IndexSets) -> None:
    """grad[freeze] = 0; param[zero] = 0; grad[zero] = 0 -- on dim 1 for 5-D (G conv) tensors, else dim 0."""
    for name in params:
        five_d = params[name].ndim == 5
        if name in freeze:
            if five_d:
                grads[name][:, freeze[name]] = 0
            else:
                grads[name][freeze[name]] = 0
        if name in zero:
            if five_d:
                params[name][:, zero[name]] = 0
                grads[name][:, zero[name]] = 0
            else:
                params[name][zero[name]] = 0
                grads[name][zero[name]] = 0

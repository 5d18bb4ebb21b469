"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by make_golden.py, tests/ and bench.py.

Everything here is a pure function of its seed (torch CPU ``Generator`` / NumPy ``default_rng``), so the
GPU box regenerates bit-identical inputs for the committed golden outputs without the reference tree.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import model_oracle as mo


def g_state(size: int, seed: int, channel_multiplier: int = 2) -> Dict[str, torch.Tensor]:
    """Random-init G state with non-trivial biases / noise strengths (the constructors' zeros would hide bugs)."""
    g = torch.Generator().manual_seed(seed)
    p = mo.init_g_params(size, channel_multiplier=channel_multiplier, generator=g)
    for k in sorted(p):
        if k.endswith("noise.weight"):
            p[k] = torch.randn(1, generator=g) * 0.1
        elif k.endswith("activate.bias") or (k.endswith(".bias") and "modulation" not in k and "style" not in k):
            p[k] = p[k] + torch.randn(p[k].shape, generator=g) * 0.1
    return p


def d_state(size: int, seed: int, channel_multiplier: int = 2) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    p = mo.init_d_params(size, channel_multiplier=channel_multiplier, generator=g)
    for k in sorted(p):
        if k.endswith(".bias"):
            p[k] = p[k] + torch.randn(p[k].shape, generator=g) * 0.1
    return p


def latents(n: int, seed: int, dim: int = 512) -> torch.Tensor:
    return torch.randn(n, dim, generator=torch.Generator().manual_seed(seed))


def shots(n: int = 10, size: int = 256, seed: int = 0) -> torch.Tensor:
    """ten synthetic 'real' images in [-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    return torch.clamp(torch.randn(n, 3, size, size, generator=g) * 0.5, -1, 1)


def _fisher_like(rng: np.random.Generator, shape, filter_axis: int) -> np.ndarray:
    """Positive, heavy-tailed, per-filter-scaled values shaped like an averaged grad**2 tensor."""
    a = rng.standard_normal(shape, dtype=np.float32) ** 2
    sc_shape = [1] * len(shape)
    sc_shape[filter_axis] = shape[filter_axis]
    scale = np.exp(rng.standard_normal(sc_shape, dtype=np.float32) * 2.0 - 14.0).astype(np.float32)
    return (a * scale).astype(np.float32)


def fisher_g(seed: int, size: int = 256, channel_multiplier: int = 2) -> Dict[str, np.ndarray]:
    """Synthetic averaged Fisher dict with the G shapes the mask step reads (train:281-299)."""
    rng = np.random.default_rng(seed)
    ch = mo.CHANNELS(channel_multiplier)
    out: Dict[str, np.ndarray] = {}
    cin = ch[4]
    for i in range(3, int(math.log(size, 2)) + 1):
        cout = ch[2 ** i]
        for j, (a, b) in enumerate(((cin, cout), (cout, cout))):
            k = f"convs.{2 * (i - 3) + j}.conv"
            out[k + ".weight"] = _fisher_like(rng, (1, b, a, 3, 3), 1)
            out[k + ".modulation.weight"] = _fisher_like(rng, (a, 512), 0)
            out[k + ".modulation.bias"] = _fisher_like(rng, (a,), 0)
        cin = cout
    return out


def fisher_d(seed: int, size: int = 256, channel_multiplier: int = 2) -> Dict[str, np.ndarray]:
    """Synthetic averaged Fisher dict with the D shapes the mask step reads (train:336-351)."""
    rng = np.random.default_rng(seed)
    ch = mo.CHANNELS(channel_multiplier)
    out: Dict[str, np.ndarray] = {}
    cin = ch[size]
    for j, i in enumerate(range(int(math.log(size, 2)), 2, -1)):
        cout = ch[2 ** (i - 1)]
        k = f"convs.{j + 1}"
        out[f"{k}.conv1.0.weight"] = _fisher_like(rng, (cin, cin, 3, 3), 0)
        out[f"{k}.conv1.1.bias"] = _fisher_like(rng, (cin,), 0)
        out[f"{k}.conv2.1.weight"] = _fisher_like(rng, (cout, cin, 3, 3), 0)
        out[f"{k}.conv2.2.bias"] = _fisher_like(rng, (cout,), 0)
        out[f"{k}.skip.1.weight"] = _fisher_like(rng, (cout, cin, 1, 1), 0)
        cin = cout
    return out


# op-level cases shared by the golden generator and the parity tests:
# (name, N, C, H, W, kh, kw, up, down, pad0, pad1, taps_kind)
UPFIRDN_CASES = [
    ("blur_after_convT", 2, 8, 17, 17, 4, 4, 1, 1, 1, 1, "blur4"),       # ModulatedConv2d upsample blur (pad (1,1))
    ("rgb_upsample", 2, 3, 8, 8, 4, 4, 2, 1, 2, 1, "blur4"),             # ToRGB skip upsample
    ("d_blur_conv2", 2, 8, 16, 16, 4, 4, 1, 1, 2, 2, "blur"),            # D conv2 blur (pad (2,2))
    ("d_blur_skip", 2, 8, 16, 16, 4, 4, 1, 1, 1, 1, "blur"),             # D skip blur (pad (1,1))
    ("downsample", 2, 4, 16, 16, 4, 4, 1, 2, 1, 1, "blur"),              # Downsample module
    ("up2_odd", 1, 5, 7, 9, 4, 4, 2, 1, 2, 1, "rand"),                   # ragged sizes, random taps
    ("updown", 1, 3, 9, 11, 5, 5, 2, 2, -1, 3, "rand"),                  # negative pad (crop), up and down
    ("big_kernel", 1, 2, 20, 20, 12, 12, 2, 1, 6, 5, "rand"),            # non_leaking.py-sized taps -> generic path
    ("k3", 2, 3, 10, 6, 3, 3, 1, 1, 1, 1, "rand"),
    ("k1x1", 1, 2, 5, 5, 1, 1, 1, 1, 0, 0, "rand"),
    ("tiny", 1, 1, 1, 1, 4, 4, 2, 1, 2, 1, "blur4"),
    ("down2_pad0", 1, 3, 12, 12, 2, 2, 1, 2, 0, 0, "rand"),
]


def upfirdn_case_inputs(case, seed: int = 0):
    name, n, c, h, w, kh, kw, up, down, p0, p1, kind = case
    g = torch.Generator().manual_seed(seed + sum(map(ord, name)))
    x = torch.randn(n, c, h, w, generator=g)
    if kind == "rand":
        taps = torch.randn(kh, kw, generator=g)
    else:
        taps = mo.make_kernel([1, 3, 3, 1]) * (4 if kind == "blur4" else 1)
    return x, taps

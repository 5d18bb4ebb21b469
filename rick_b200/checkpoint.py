"""Checkpoints in the reference's on-disk format (train_dynamic_update_prune.py:645-659):

    torch.save({"g_ema": g_ema.state_dict(), "g": g_module.state_dict(), "d": d_module.state_dict(),
                "g_optim": g_optim.state_dict(), "d_optim": d_optim.state_dict()}, path)

and its loading side (train:872-879): ``generator <- ckpt["g"]``, ``g_ema <- ckpt["g_ema"]``, ``discriminator <- ckpt["d"]``,
``d_ema <- ckpt["d"]``, all ``strict=False``.  Parameter / buffer names and shapes of rick_b200.stylegan2 are the
reference's, so files move in both directions; tensors are written contiguous (the channels-last storage of D's conv
weights is a private detail of this package).
"""
from __future__ import annotations

import torch


def _plain(sd: dict) -> dict:
    return {k: (v.detach().contiguous().cpu() if torch.is_tensor(v) else v) for k, v in sd.items()}


def _plain_optim(sd: dict) -> dict:
    state = {i: {k: (v.detach().contiguous().cpu() if torch.is_tensor(v) else v) for k, v in st.items()}
             for i, st in sd["state"].items()}
    return {"state": state, "param_groups": sd["param_groups"]}


def save(path, adapter):
    """Write ``adapter``'s networks and optimisers (a rick_b200.adapt.RickAdapter) with the reference's five keys."""
    torch.save({"g_ema": _plain(adapter.g_ema.state_dict()), "g": _plain(adapter.g.state_dict()),
                "d": _plain(adapter.d.state_dict()), "g_optim": _plain_optim(adapter.g_optim.state_dict()),
                "d_optim": _plain_optim(adapter.d_optim.state_dict())}, path)


def read(path, map_location="cpu", trust_pickle: bool = False) -> dict:
    """``torch.load`` restricted to tensors / containers of primitives (``weights_only=True``): everything the
    five-key checkpoint holds.  ``trust_pickle=True`` is the explicit opt-in for legacy files that pickled other
    objects -- unpickling those can execute arbitrary code, so only use it on files you produced."""
    return torch.load(path, map_location=map_location, weights_only=not trust_pickle)


def load_networks(path, generator, discriminator, g_ema=None, d_ema=None, map_location="cpu",
                  trust_pickle: bool = False) -> dict:
    """The reference's start-up (train:872-879): G from ``"g"``, g_ema from ``"g_ema"``, D and d_ema from ``"d"``."""
    ckpt = read(path, map_location, trust_pickle)
    generator.load_state_dict(ckpt["g"], strict=False)
    if g_ema is not None:
        g_ema.load_state_dict(ckpt["g_ema"], strict=False)
    discriminator.load_state_dict(ckpt["d"], strict=False)
    if d_ema is not None:
        d_ema.load_state_dict(ckpt["d"], strict=False)
    return ckpt


def resume(path, adapter, map_location="cpu", trust_pickle: bool = False) -> dict:
    """:func:`load_networks` plus both optimiser states (the keys the reference writes "if you wish to resume")."""
    ckpt = load_networks(path, adapter.g, adapter.d, adapter.g_ema, adapter.d_ema, map_location, trust_pickle)
    if "g_optim" in ckpt:
        adapter.g_optim.load_state_dict(ckpt["g_optim"])
    if "d_optim" in ckpt:
        adapter.d_optim.load_state_dict(ckpt["d_optim"])
    return ckpt

"""TEST INFRASTRUCTURE ONLY -- load the REAL reference from /root/reference (authoring container only).

``/root/reference`` does not exist on the GPU box, so nothing that runs there may
import this module; it is used only by ``oracle/make_golden.py``, which writes the
``tests/golden/*`` fixtures that ``tests/test_oracle_golden.py`` checks the oracle against.

Why not ``import op`` / ``import gan_training.models``:
  * ``op/upfirdn2d.py`` and ``op/fused_act.py`` JIT-compile CUDA extensions at import
    (op/upfirdn2d.py:10-16, op/fused_act.py:10-16) and ``fused_act`` has no CPU branch;
  * ``gan_training/models/__init__.py:1-4`` imports a module that does not exist.
So: the reference's own ``upfirdn2d_native`` and ``FusedLeakyReLU`` definitions are
pulled out of their source files by AST (executed unmodified, nothing is copied
into this repo), wrapped in a shim ``op`` package, and ``model_probe_tune.py`` is
then imported by file path against that shim.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import textwrap
import types

import torch
from torch import nn
from torch.nn import functional as F

REF_ROOT = os.environ.get("RICK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "op", "upfirdn2d.py"))


def _extract(path: str, names, namespace: dict) -> None:
    """exec only the named top-level defs/classes of a reference source file."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, namespace)


_cache: dict = {}


def load():
    """Returns a namespace with: upfirdn2d_native, upfirdn2d, fused_leaky_relu, FusedLeakyReLU, model (module)."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    ns_up: dict = {"torch": torch, "F": F}
    _extract(os.path.join(REF_ROOT, "op", "upfirdn2d.py"), {"upfirdn2d_native"}, ns_up)
    upfirdn2d_native = ns_up["upfirdn2d_native"]

    def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
        # the CPU branch of the reference dispatcher (op/upfirdn2d.py:146-149)
        return upfirdn2d_native(input, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])

    def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
        # no CPU branch in the reference; arithmetic of op/fused_bias_act_kernel.cu:26-47, same form as
        # ScaledLeakyReLU (model_probe_tune.py:176-185) with the bias add in front.
        shape = [1, -1] + [1] * (input.ndim - 2)
        return F.leaky_relu(input + bias.view(shape), negative_slope) * scale

    ns_act: dict = {"torch": torch, "nn": nn, "fused_leaky_relu": fused_leaky_relu}
    _extract(os.path.join(REF_ROOT, "op", "fused_act.py"), {"FusedLeakyReLU"}, ns_act)

    shim = types.ModuleType("op")
    shim.upfirdn2d = upfirdn2d
    shim.fused_leaky_relu = fused_leaky_relu
    shim.FusedLeakyReLU = ns_act["FusedLeakyReLU"]
    saved = sys.modules.get("op")
    sys.modules["op"] = shim
    try:
        spec = importlib.util.spec_from_file_location(
            "_rick_reference_model", os.path.join(REF_ROOT, "gan_training", "models", "model_probe_tune.py"))
        model = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(model)
    finally:
        if saved is None:
            del sys.modules["op"]
        else:
            sys.modules["op"] = saved

    ns = types.SimpleNamespace(upfirdn2d_native=upfirdn2d_native, upfirdn2d=upfirdn2d,
                               fused_leaky_relu=fused_leaky_relu, FusedLeakyReLU=ns_act["FusedLeakyReLU"],
                               model=model)
    _cache["ns"] = ns
    return ns


def train_source_lines(first: int, last: int) -> str:
    """Dedented source of train_dynamic_update_prune.py lines [first, last] (1-based, inclusive)."""
    with open(os.path.join(REF_ROOT, "train_dynamic_update_prune.py")) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def train_functions(names):
    """The reference's own top-level helper functions from the training script (losses etc.)."""
    import math
    import random
    import numpy as np
    from torch import autograd
    ns: dict = {"torch": torch, "F": F, "autograd": autograd, "math": math, "np": np, "random": random}
    _extract(os.path.join(REF_ROOT, "train_dynamic_update_prune.py"), set(names), ns)
    return types.SimpleNamespace(**{n: ns[n] for n in names})

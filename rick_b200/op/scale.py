"""One-launch equalised-lr scaling of a whole network's weights (rick_scale_multi, rick_b200/csrc/optim.cu).

The reference multiplies every EqualConv2d / EqualLinear weight by its ``scale`` (and every EqualLinear bias by
``lr_mul``) inside each module's forward (model_probe_tune.py:124, 160-164): ~45 two-microsecond launches per network pass
and as many in backward.  ``scale_all`` does them in one launch each way; second derivatives fall back to per-tensor
torch ops (the map is linear, so that branch is only there to stay differentiable)."""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import torch

from .. import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dense(t: torch.Tensor) -> bool:
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def _launch(outs: Sequence[torch.Tensor], ins: Sequence[torch.Tensor], scales: Sequence[float]):
    n = len(ins)
    sc = (ctypes.c_float * n)(*[float(s) for s in scales])
    with torch.cuda.device(ins[0].device):
        rc = _lib.lib().rick_scale_multi(_lib.ptr_table([o.data_ptr() for o in outs]),
                                         _lib.ptr_table([i.data_ptr() for i in ins]), sc,
                                         _lib.i64_table([i.numel() for i in ins]), n, _stream())
    _lib.check(rc, "rick_scale_multi")


class _ScaleAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scales, *tensors):
        ctx.scales = scales
        outs = [torch.empty_like(t) for t in tensors]          # preserve_format: same strides as the parameter
        _launch(outs, tensors, scales)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        scales = ctx.scales
        if torch.is_grad_enabled():                            # create_graph: stay differentiable
            return (None,) + tuple(None if g is None else g * s for g, s in zip(grads, scales))
        res: List = [None] * len(grads)
        idx = [i for i, g in enumerate(grads) if g is not None and ctx.needs_input_grad[i + 1]]
        fast = [i for i in idx if _dense(grads[i]) and grads[i].dtype == torch.float32]
        if fast:
            outs = [torch.empty_like(grads[i]) for i in fast]
            _launch(outs, [grads[i] for i in fast], [scales[i] for i in fast])
            for i, o in zip(fast, outs):
                res[i] = o
        for i in idx:
            if res[i] is None:
                res[i] = grads[i] * scales[i]
        return (None,) + tuple(res)


def scale_all(tensors: Sequence[torch.Tensor], scales: Sequence[float]) -> List[torch.Tensor]:
    """[t * s for t, s in zip(tensors, scales)] for dense float32 CUDA tensors, differentiable to any order."""
    tensors = list(tensors)
    if not tensors:
        return []
    for t in tensors:
        if not t.is_cuda or t.dtype != torch.float32 or not _dense(t):
            return [t * s for t, s in zip(tensors, scales)]
    return list(_ScaleAll.apply(tuple(float(s) for s in scales), *tensors))

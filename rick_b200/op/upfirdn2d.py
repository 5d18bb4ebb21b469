"""``upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))`` -- same call surface as the reference's
op/upfirdn2d.py:145-156, executed by ``rick_upfirdn2d`` (rick_b200/csrc/upfirdn2d.cu) through the C ABI.

Autograd: the operator is linear, and its adjoint is the same operator with up <-> down swapped, the taps
flipped and the pads of op/upfirdn2d.py:111-114.  One self-recursive ``autograd.Function`` therefore provides
forward, backward and every higher derivative (the reference needs two nested Functions for fwd/bwd/double-bwd,
op/upfirdn2d.py:19-142).  The tap flip is a flag of the kernel, so no ``torch.flip`` launch per call.

CUDA only: a CPU tensor raises (the reference's CPU branch, ``upfirdn2d_native``, lives in ``oracle/`` as the checker).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib

_DTYPES = {torch.float32: _lib.RICK_F32, torch.bfloat16: _lib.RICK_BF16}


def _is_channels_last(x: torch.Tensor) -> bool:
    return x.dim() == 4 and x.shape[1] > 1 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)


def _run(x: torch.Tensor, taps: torch.Tensor, up, down, pad, flip: bool) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("rick_b200.op.upfirdn2d: input must be a CUDA tensor (no CPU fallback in this package)")
    if not taps.is_cuda:
        raise RuntimeError("rick_b200.op.upfirdn2d: kernel must be a CUDA tensor")
    if x.dtype not in _DTYPES:
        raise RuntimeError(f"rick_b200.op.upfirdn2d: unsupported dtype {x.dtype} (float32 / bfloat16)")
    if x.dim() != 4 or taps.dim() != 2:
        raise RuntimeError("rick_b200.op.upfirdn2d: expected input (N, C, H, W) and kernel (kh, kw)")
    up_x, up_y = up
    down_x, down_y = down
    px0, px1, py0, py1 = pad
    n, c, h, w = x.shape
    kh, kw = taps.shape
    lib = _lib.lib()
    oh = lib.rick_upfirdn2d_out_size(h, kh, up_y, down_y, py0, py1)
    ow = lib.rick_upfirdn2d_out_size(w, kw, up_x, down_x, px0, px1)
    if oh < 1 or ow < 1:
        raise RuntimeError(f"rick_b200.op.upfirdn2d: empty output ({oh} x {ow})")
    taps = taps.detach().to(torch.float32).contiguous()
    if (_is_channels_last(x) and x.dtype == torch.float32 and c % 4 == 0 and (kh, kw) == (4, 4)
            and up_x == up_y == down_x == down_y == 1 and px0 == py0 and px1 == py1):
        # channels-last 4x4 blur: the NHWC sliding-window kernel (128-bit accesses along C)
        from .. import conv_tc as _ct
        out_nhwc = _ct.blur_nhwc(x.permute(0, 2, 3, 1), taps, (px0, px1), flip=flip)
        return out_nhwc.permute(0, 3, 1, 2)           # logical NCHW view of channels-last memory
    if _is_channels_last(x):
        major, minor = n, c
        out = torch.empty((n, c, oh, ow), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    else:
        x = x.contiguous()
        major, minor = n * c, 1
        out = torch.empty((n, c, oh, ow), dtype=x.dtype, device=x.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        st = lib.rick_upfirdn2d(out.data_ptr(), x.data_ptr(), taps.data_ptr(), major, h, w, minor, kh, kw, up_x, up_y,
                                down_x, down_y, px0, px1, py0, py1, int(flip), _DTYPES[x.dtype],
                                torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rick_upfirdn2d")
    return out


class UpFirDn2d(Function):
    """y = upfirdn2d(x); backward re-enters this Function with the adjoint parameters."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad, flip):
        ctx.save_for_backward(kernel)
        ctx.cfg = (tuple(up), tuple(down), tuple(pad), bool(flip), tuple(input.shape[2:]))
        return _run(input, kernel, up, down, pad, flip)

    @staticmethod
    def backward(ctx, grad_output):
        (kernel,) = ctx.saved_tensors
        (up_x, up_y), (down_x, down_y), (px0, px1, py0, py1), flip, (in_h, in_w) = ctx.cfg
        kh, kw = kernel.shape
        out_h, out_w = grad_output.shape[2:]
        g_pad = (kw - px0 - 1, in_w * up_x - out_w * down_x + px0 - up_x + 1,
                 kh - py0 - 1, in_h * up_y - out_h * down_y + py0 - up_y + 1)
        grad_input = UpFirDn2d.apply(grad_output, kernel, (down_x, down_y), (up_x, up_y), g_pad, not flip)
        return grad_input, None, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]), False)


def upfirdn2d_adjoint(grad_output, kernel, up=1, down=1, pad=(0, 0), in_size=None):
    """The gradient of ``upfirdn2d(x, kernel, up, down, pad)`` w.r.t. ``x`` for an input of spatial size ``in_size``
    (H, W): the same operator with up <-> down swapped, flipped taps and the pads of op/upfirdn2d.py:111-114 -- what
    ``UpFirDn2d.backward`` launches, callable directly (benchmarks, callers that manage their own graphs)."""
    kh, kw = kernel.shape
    in_h, in_w = in_size
    out_h, out_w = grad_output.shape[2:]
    g_pad = (kw - pad[0] - 1, in_w * up - out_w * down + pad[0] - up + 1,
             kh - pad[0] - 1, in_h * up - out_h * down + pad[0] - up + 1)
    return UpFirDn2d.apply(grad_output, kernel, (down, down), (up, up), g_pad, True)


def upfirdn2d_xy(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """The native entry point's full parameter set (op/upfirdn2d.cpp:12-19)."""
    return UpFirDn2d.apply(input, kernel, (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1), False)

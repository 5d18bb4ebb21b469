"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two reference operators.

upfirdn2d   follows op/upfirdn2d.py:159-200 (``upfirdn2d_native``): zero-stuff by
            ``up``, pad (negative pad = crop), correlate with the flipped FIR
            taps, keep every ``down``-th sample.  Output size is
            ``(in*up + pad0 + pad1 - k) // down + 1`` (op/upfirdn2d.py:103-104).
bias_act    follows op/fused_bias_act_kernel.cu:26-47 (act=3 leaky-ReLU, grad
            modes 0/1/2) and the autograd wiring of op/fused_act.py:19-70.  The
            reference ships no CPU branch for this op, so this is a restatement
            of the kernel's arithmetic: ``x += b[c]; y = (ref>0 ? x : x*alpha) * scale``.

Everything here is plain torch on CPU and differentiable by torch autograd
(which is how the oracle gets first and second derivatives).
"""
from __future__ import annotations

import torch
from torch.nn import functional as F


def upfirdn2d_out_size(in_size: int, k: int, up: int, down: int, pad0: int, pad1: int) -> int:
    """op/upfirdn2d.py:103-104 / op/upfirdn2d_kernel.cu:237-240."""
    return (in_size * up + pad0 + pad1 - k) // down + 1


def upfirdn2d_xy(x: torch.Tensor, taps: torch.Tensor, up_x: int, up_y: int, down_x: int, down_y: int,
                 pad_x0: int, pad_x1: int, pad_y0: int, pad_y1: int) -> torch.Tensor:
    """General (separate x/y factors) form; x is (N, C, H, W), taps is (kh, kw)."""
    n, c, h, w = x.shape
    kh, kw = taps.shape
    planes = x.reshape(n * c, 1, h, w)

    # zero-stuffing: sample (iy, ix) lands on (iy*up_y, ix*up_x), trailing zeros kept
    stuffed = planes.new_zeros(n * c, 1, h * up_y, w * up_x)
    stuffed[:, :, ::up_y, ::up_x] = planes

    # positive pads add zeros, negative pads crop
    stuffed = F.pad(stuffed, [max(pad_x0, 0), max(pad_x1, 0), max(pad_y0, 0), max(pad_y1, 0)])
    hh, ww = stuffed.shape[2], stuffed.shape[3]
    stuffed = stuffed[:, :, max(-pad_y0, 0): hh - max(-pad_y1, 0), max(-pad_x0, 0): ww - max(-pad_x1, 0)]

    # true convolution == correlation with the flipped taps
    flipped = torch.flip(taps, [0, 1]).reshape(1, 1, kh, kw).to(x.dtype)
    full = F.conv2d(stuffed, flipped)
    out = full[:, :, ::down_y, ::down_x]

    oh = upfirdn2d_out_size(h, kh, up_y, down_y, pad_y0, pad_y1)
    ow = upfirdn2d_out_size(w, kw, up_x, down_x, pad_x0, pad_x1)
    return out.reshape(n, c, oh, ow)


def upfirdn2d(x: torch.Tensor, taps: torch.Tensor, up: int = 1, down: int = 1, pad=(0, 0)) -> torch.Tensor:
    """Same call signature as the reference's ``op.upfirdn2d`` (op/upfirdn2d.py:145)."""
    return upfirdn2d_xy(x, taps, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def bias_act(x: torch.Tensor, bias: torch.Tensor | None, ref: torch.Tensor | None,
             act: int, grad: int, alpha: float, scale: float) -> torch.Tensor:
    """Restatement of ``fused_bias_act(input, bias, refer, act, grad, alpha, scale)``.

    act 1 = linear, act 3 = leaky-ReLU; grad 0 = forward, 1 = first derivative
    gated by ``ref``, 2 = second derivative (identically zero).  Bias is
    indexed by dim 1 (op/fused_bias_act_kernel.cu:67-71: ``(i / step_b) % size_b``).
    """
    y = x
    if bias is not None and bias.numel() > 0:
        shape = [1, -1] + [1] * (x.ndim - 2)
        y = y + bias.reshape(shape)
    if grad == 2:
        y = torch.zeros_like(y)
    elif act == 3:
        gate = y if grad == 0 else ref
        y = torch.where(gate > 0, y, y * alpha)
    return y * scale


def fused_leaky_relu(x: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2,
                     scale: float = 2 ** 0.5) -> torch.Tensor:
    """``op.fused_leaky_relu`` (op/fused_act.py:106-107) as a differentiable composite."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    return F.leaky_relu(x + bias.reshape(shape), negative_slope) * scale


def fused_leaky_relu_backward(grad_out: torch.Tensor, out: torch.Tensor, negative_slope: float = 0.2,
                              scale: float = 2 ** 0.5):
    """op/fused_act.py:19-39: grad_input via kernel mode (act=3, grad=1, ref=out),
    grad_bias = grad_input summed over every dim but 1."""
    gi = bias_act(grad_out, None, out, 3, 1, negative_slope, scale)
    dims = [0] + list(range(2, gi.ndim))
    return gi, gi.sum(dims)

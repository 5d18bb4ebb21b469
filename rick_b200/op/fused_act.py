"""``fused_leaky_relu`` / ``FusedLeakyReLU`` -- same call surface as the reference's op/fused_act.py, executed by
``rick_bias_act`` / ``rick_bias_act_bwd`` (rick_b200/csrc/bias_act.cu) through the C ABI.

Derivative structure follows the operator's maths (and therefore matches op/fused_act.py:19-70):
  forward      out = lrelu(x + b[c]) * scale                      saves ``out`` (its sign is the gate)
  backward     grad_x = (out > 0 ? g : g * slope) * scale,  grad_b = sum_{n,hw} grad_x   -- ONE fused kernel pass
               (the reference runs the elementwise kernel and then a separate ATen ``sum``)
  double-bwd   gg_out = (out > 0 ? v : v * slope) * scale  with  v = gg_x + gg_b[c]
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function

from .. import _lib

_DTYPES = {torch.float32: _lib.RICK_F32, torch.bfloat16: _lib.RICK_BF16}


def _check(x: torch.Tensor, what: str):
    if not x.is_cuda:
        raise RuntimeError(f"rick_b200.op.{what}: input must be a CUDA tensor (no CPU fallback in this package)")
    if x.dtype not in _DTYPES:
        raise RuntimeError(f"rick_b200.op.{what}: unsupported dtype {x.dtype} (float32 / bfloat16)")


def _is_cl(x: torch.Tensor) -> bool:
    """4-D tensor stored channels-last (and not also plain-contiguous)."""
    return x.dim() == 4 and x.shape[1] > 1 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)


def bias_act(x, bias, ref, act, grad, alpha, scale):
    """Native-signature call: ``fused_bias_act(input, bias, refer, act, grad, alpha, scale)`` (op/fused_bias_act.cpp:11).
    Channels-last inputs are processed in place of their physical (N, H, W, C) layout (bias index = innermost)."""
    _check(x, "bias_act")
    cl = _is_cl(x)
    fmt = torch.channels_last if cl else torch.contiguous_format
    x = x.contiguous(memory_format=fmt)
    out = torch.empty_like(x, memory_format=fmt)
    if x.numel() == 0:
        return out
    b_ptr = r_ptr = None
    step_b = size_b = 1
    if bias is not None and bias.numel() > 0:
        bias = bias.to(x.dtype).contiguous()
        size_b = bias.numel()
        if x.dim() < 2 or x.shape[1] != size_b:
            raise RuntimeError(f"rick_b200.op.bias_act: bias has {size_b} entries but input dim 1 is "
                               f"{x.shape[1] if x.dim() > 1 else 'missing'}")
        step_b = 1
        if not cl:
            for d in x.shape[2:]:
                step_b *= d
        b_ptr = bias.data_ptr()
    if ref is not None and ref.numel() > 0:
        ref = ref.to(x.dtype).contiguous(memory_format=fmt)
        if ref.shape != x.shape:
            raise RuntimeError("rick_b200.op.bias_act: refer must have the input's shape")
        r_ptr = ref.data_ptr()
    with torch.cuda.device(x.device):
        st = _lib.lib().rick_bias_act(out.data_ptr(), x.data_ptr(), b_ptr, r_ptr, x.numel(), step_b, size_b, act, grad,
                                      float(alpha), float(scale), _DTYPES[x.dtype],
                                      torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rick_bias_act")
    return out


class FusedLeakyReLUFunctionBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        _check(grad_output, "fused_leaky_relu (backward)")
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        lib = _lib.lib()
        if _is_cl(out) and out.dtype == torch.float32 and out.shape[1] % 4 == 0:
            # channels-last: (N*H*W, C) matrix, bias gradient = column sums (rick_bias_act_bwd_nhwc)
            g = grad_output.contiguous(memory_format=torch.channels_last)
            c = g.shape[1]
            rows = g.numel() // c
            grad_input = torch.empty_like(g, memory_format=torch.channels_last)
            grad_bias = torch.empty(c, dtype=torch.float32, device=g.device)
            ws = torch.empty(max(int(lib.rick_bias_act_bwd_nhwc_workspace(rows, c)), 4), dtype=torch.uint8,
                             device=g.device)
            with torch.cuda.device(g.device):
                st = lib.rick_bias_act_bwd_nhwc(grad_input.data_ptr(), grad_bias.data_ptr(), ws.data_ptr(), g.data_ptr(),
                                                out.data_ptr(), rows, c, float(negative_slope), float(scale),
                                                torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "rick_bias_act_bwd_nhwc")
            return grad_input, grad_bias
        out = out.contiguous()
        g = grad_output.contiguous()
        n, c = g.shape[0], g.shape[1]
        hw = 1
        for d in g.shape[2:]:
            hw *= d
        grad_input = torch.empty_like(g)
        grad_bias = torch.empty(c, dtype=torch.float32, device=g.device)
        ws = torch.empty(max(int(lib.rick_bias_act_bwd_workspace(n, c, hw)), 4), dtype=torch.uint8, device=g.device)
        with torch.cuda.device(g.device):
            st = lib.rick_bias_act_bwd(grad_input.data_ptr(), grad_bias.data_ptr(), ws.data_ptr(), g.data_ptr(),
                                       out.data_ptr(), n, c, hw, float(negative_slope), float(scale), _DTYPES[g.dtype],
                                       torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "rick_bias_act_bwd")
        return grad_input, grad_bias.to(g.dtype)

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        (out,) = ctx.saved_tensors
        gradgrad_out = bias_act(gradgrad_input, gradgrad_bias, out, _lib.ACT_LRELU, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = bias_act(input, bias, None, _lib.ACT_LRELU, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


class FusedLeakyReLU_kml(nn.Module):
    """Exported by the reference (op/fused_act.py:85-103); unused by the hot path, kept for surface parity."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.b_vector = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        bias = self.bias + self.b_vector if self.b_vector.requires_grad else self.bias
        return fused_leaky_relu(input, bias, self.negative_slope, self.scale)

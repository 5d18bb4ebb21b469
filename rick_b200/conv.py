"""Convolution dispatch for the StyleGAN2 stack.

Two executors sit behind ``modulated_conv2d`` / ``conv2d``:

  * ``tc``     the hand-written tcgen05 / TMEM / TMA implicit-GEMM kernels of this package
               (rick_b200/csrc/conv_tc.cu via the C ABI) -- TF32 operands, fp32 accumulation in tensor memory,
               demodulation + noise + bias + leaky-ReLU fused into the epilogue.  Used wherever it is implemented
               (see ``tc_supported``); the status table lives in DESIGN.md.
  * ``cudnn``  ATen / cuDNN dense convolutions on the SHARED weight (``groups = 1``: the modulation is folded into
               the activations and the demodulation into the output, so this is a plain library conv, not the
               reference's grouped conv on materialised per-sample weights).  It is the differentiable path
               (first and second derivatives for R1 / path-length) until the tcgen05 dgrad / wgrad kernels land.

There is no CPU executor: inputs must be CUDA tensors.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch.nn import functional as F

from .op import styled as _styled

_FORCE = os.environ.get("RICK_CONV_BACKEND", "")   # "", "cudnn" or "tc" (tests use it to pin an executor)


# ---------------------------------------------------------------------------------------------------------------
# library convolution with explicit first- and second-order gradients
# ---------------------------------------------------------------------------------------------------------------
# R1 (train:462-493) and the path-length regulariser (train:546-589) differentiate THROUGH a backward pass.  Left to
# autograd's generic double-backward formula, the weight-gradient term of a convolution is evaluated as a forward
# convolution whose "filter" is a whole feature map (a (128, 1, 256, 256) input against a (128, 1, 256, 256) weight at
# 256 px): the library has no tensor-core kernel for that and one path-length iteration spent 10 of its 33 ms there
# (round-1 torch.profiler run, scripts/profile_path.py).  Written out, every term of the second derivative is again a
# forward conv, a data gradient or a weight gradient with ordinary shapes:
#     y  = conv(x, w)          gx = dgrad(g, w)          gw = wgrad(g, x)
#     d<ggx, gx>/dg = conv(ggx, w)      d<ggx, gx>/dw = wgrad(g, ggx)
#     d<ggw, gw>/dg = conv(x, ggw)      d<ggw, gw>/dx = dgrad(g, ggw)
# so the two Functions below call themselves / each other and stay differentiable to any order.
def _conv_args(stride: int, padding: int, transposed: bool):
    return [stride, stride], [padding, padding], [1, 1], transposed, [0, 0], 1


class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, padding, transposed):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, transposed)
        return torch.ops.aten.convolution(x, w, None, *_conv_args(stride, padding, transposed))

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not torch.is_grad_enabled():                    # plain backward (no create_graph): straight to the library
            gx, gw, _ = torch.ops.aten.convolution_backward(g, x, w, None, *_conv_args(*ctx.cfg),
                                                            [bool(need_x), bool(need_w), False])
            return gx, gw, None, None, None
        gx, gw = _ConvGrad.apply(g, x, w, need_x, need_w, *ctx.cfg)
        return gx, gw, None, None, None


class _ConvGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, x, w, need_x, need_w, stride, padding, transposed):
        ctx.save_for_backward(g, x, w)
        ctx.cfg = (stride, padding, transposed)
        gx, gw, _ = torch.ops.aten.convolution_backward(g, x, w, None, *_conv_args(stride, padding, transposed),
                                                        [bool(need_x), bool(need_w), False])
        ctx.have = (bool(need_x), bool(need_w))
        return gx, gw                                      # None for a gradient that was not asked for

    @staticmethod
    def backward(ctx, ggx, ggw):
        g, x, w = ctx.saved_tensors
        cfg = ctx.cfg
        have_x, have_w = ctx.have
        ggx = ggx if have_x else None
        ggw = ggw if have_w else None
        need_g, need_x, need_w = ctx.needs_input_grad[:3]
        d_g = d_x = d_w = None
        if need_g:
            if ggx is not None:
                d_g = _Conv.apply(ggx, w, *cfg)
            if ggw is not None:
                t = _Conv.apply(x, ggw, *cfg)
                d_g = t if d_g is None else d_g + t
        if need_w and ggx is not None:
            d_w = _ConvGrad.apply(g, ggx, w, False, True, *cfg)[1]
        if need_x and ggw is not None:
            d_x = _ConvGrad.apply(g, x, ggw, True, False, *cfg)[0]
        return d_g, d_x, d_w, None, None, None, None, None


def _lib_conv(x, w, stride: int = 1, padding: int = 0, transposed: bool = False):
    """``F.conv2d`` / ``F.conv_transpose2d`` (groups 1, no bias) with the gradient structure above."""
    return _Conv.apply(x, w, stride, padding, transposed)


def _require_cuda(x: torch.Tensor, what: str):
    if not x.is_cuda:
        raise RuntimeError(f"rick_b200.conv.{what}: input must be a CUDA tensor (no CPU fallback in this package)")


def conv2d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1, padding: int = 0):
    """EqualConv2d's dense convolution (model_probe_tune.py:121-130)."""
    _require_cuda(x, "conv2d")
    co, ci, kh, kw = weight.shape
    if kh == 1 and kw == 1 and ci <= 4 and stride == 1 and padding == 0:
        # D's from-RGB layer (3 -> C, 1x1).  As a library conv its weight gradient is a (C x 3) GEMM with
        # K = B*H*W = 131072 that cuBLAS runs on 4 CTAs (0.96 ms per call, round-1 ncu launch list); as ci broadcast
        # multiply-adds it is a handful of memory-bound passes and stays differentiable to any order.
        out = x[:, 0:1] * weight[:, 0].reshape(1, co, 1, 1)
        for c in range(1, ci):
            out = out + x[:, c:c + 1] * weight[:, c].reshape(1, co, 1, 1)
        if bias is not None:
            out = out + bias.reshape(1, co, 1, 1)
        if x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
            out = out.contiguous(memory_format=torch.channels_last)     # broadcasting does not keep the NHWC strides
        return out
    if ci % 4 != 0 and ci > 4:
        # D's final conv has 512 + 1 (minibatch-stddev) input channels; with a channel count that is not a multiple of
        # 4 the library falls back to a SIMT convolution (0.7 ms for a 4x4 map in the round-1 profile).  Zero-padding
        # the channel axis of input and weight keeps the result identical and the tensor-core kernels eligible.
        extra = 4 - ci % 4
        x = F.pad(x, (0, 0, 0, 0, 0, extra))
        weight = F.pad(weight, (0, 0, 0, 0, 0, extra))
    out = _lib_conv(x, weight, stride, padding)
    return out if bias is None else out + bias.reshape(1, -1, 1, 1)


def modulated_conv2d(x: torch.Tensor, w: torch.Tensor, s: torch.Tensor, demod: Optional[torch.Tensor],
                     upsample: bool = False, downsample: bool = False, padding: int = 1, blur=None,
                     epilogue=None) -> torch.Tensor:
    """out[b] = demod[b] * conv(x[b] * s[b], w)  -- ModulatedConv2d (model_probe_tune.py:243-284) without
    per-sample weights.  ``w`` is the shared (Cout, Cin, k, k) weight already multiplied by the equalised-lr scale,
    ``s`` the (B, Cin) style, ``demod`` the (B, Cout) demodulation or None."""
    _require_cuda(x, "modulated_conv2d")
    co, ci, kh, kw = w.shape
    if co <= 4 and kh == 1 and kw == 1 and not upsample and not downsample:
        # ToRGB: a 1x1 conv to 3 channels.  As a library conv it lands on a SIMT kernel (0.7 ms at 256 px in the
        # round-1 profile); as a batched matmul over the (pixels, Cin) view it is one pass over the activation.
        b, _, h, wd = x.shape
        wmod = w.reshape(1, co, ci) * s[:, None, :]                      # (B, 3, Cin)
        if demod is not None:
            wmod = wmod * demod[:, :, None]
        xf = x.permute(0, 2, 3, 1).reshape(b, h * wd, ci)                # free view when x is channels-last
        out = torch.bmm(xf, wmod.transpose(1, 2))                        # (B, HW, 3)
        return out.transpose(1, 2).reshape(b, co, h, wd)
    if epilogue is not None and demod is not None and not downsample and _styled.fused_ok(x) and w.shape[0] % 4 == 0:
        # StyledConv: modulate -> conv [-> blur] -> demod + noise + bias + leaky-ReLU, two fused ops around the conv
        noise, noise_weight, bias, slope, act_scale = epilogue
        xm = _styled.modulate(x, s)
        if upsample:
            out = blur(_lib_conv(xm, w.transpose(0, 1), 2, 0, True))
        else:
            out = _lib_conv(xm, w, 1, padding)
        if noise is None:
            noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
        return _styled.styled_epilogue(out, demod, noise, noise_weight, bias, slope, act_scale)
    xm = x * s[:, :, None, None]
    if upsample:
        out = _lib_conv(xm, w.transpose(0, 1), 2, 0, True)
        out = blur(out)
    elif downsample:
        out = _lib_conv(blur(xm), w, 2, 0)
    else:
        out = _lib_conv(xm, w, 1, padding)
    if demod is not None:
        out = out * demod[:, :, None, None]
    return out

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
    return F.conv2d(x, weight, bias=bias, stride=stride, padding=padding)


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
            out = blur(F.conv_transpose2d(xm, w.transpose(0, 1), stride=2, padding=0))
        else:
            out = F.conv2d(xm, w, padding=padding)
        if noise is None:
            noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
        return _styled.styled_epilogue(out, demod, noise, noise_weight, bias, slope, act_scale)
    xm = x * s[:, :, None, None]
    if upsample:
        out = F.conv_transpose2d(xm, w.transpose(0, 1), stride=2, padding=0)
        out = blur(out)
    elif downsample:
        out = F.conv2d(blur(xm), w, stride=2, padding=0)
    else:
        out = F.conv2d(xm, w, padding=padding)
    if demod is not None:
        out = out * demod[:, :, None, None]
    return out

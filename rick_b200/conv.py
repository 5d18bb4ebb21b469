"""Convolutions of the StyleGAN2 stack, forward AND backward, on the tcgen05 kernels of this package.

``_Conv`` / ``_ConvGrad`` are autograd Functions over three primitives

    fprop(x, w)        y  = conv(x, w)                      rick_conv_tc_w      (csrc/conv_tc.cu)
    dgrad(g, w)        gx = conv^T(g, w)                    rick_conv_tc_w, the weight read transposed in place
    wgrad(g, x)        gw = sum_pixels g (x) x              rick_conv_wgrad_tc  (csrc/conv_wgrad.cu)

for ordinary (stride 1 / 2) and stride-2 transposed convolutions -- every convolution ModulatedConv2d (in its algebraic
form: the modulation folded into the activations, the demodulation into the output, model_probe_tune.py:243-284) and
EqualConv2d (:122-128) run.  The reference gets these from cuDNN through autograd; here they are TF32 tcgen05 implicit
GEMMs on channels-last activations and channels-last weights, with no re-packed or transposed weight copies.

A primitive falls back to the ATen library call only when the kernels do not cover its shape (channel counts that are
not multiples of 32, e.g. D's 3-channel from-RGB layer; non-channels-last operands) or when TF32 math is switched off
(``torch.backends.cudnn.allow_tf32 = False``, which tests use to isolate the other kernels in fp32).
``RICK_CONV_BACKEND=cudnn`` forces the library, ``=tc`` makes an uncovered shape an error.  ``launch_stats`` counts both.

There is no CPU executor: inputs must be CUDA tensors.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch.nn import functional as F

from . import conv_tc as _ct
from .op import styled as _styled

_FORCE = os.environ.get("RICK_CONV_BACKEND", "")   # "", "cudnn" or "tc" (tests use it to pin an executor)
launch_stats = {"tc": 0, "library": 0}

# Weight tensors (by data pointer) whose gradient the backward pass in flight does not want.  ``ctx.needs_input_grad`` is fixed
# when the forward runs, so a backward restricted to OTHER tensors (the Fisher round differentiates the generator loss through
# D for G's parameters only, train:239-247) would still compute -- and autograd then drop -- every weight gradient of D.
_NO_WGRAD_PTRS: set = set()


class skip_weight_grads:
    """``with skip_weight_grads(ptrs):`` -- the convolution Functions skip the weight gradients of these weights."""

    def __init__(self, ptrs):
        self.ptrs = set(ptrs)

    def __enter__(self):
        self.prev = set(_NO_WGRAD_PTRS)
        _NO_WGRAD_PTRS.update(self.ptrs)
        return self

    def __exit__(self, *exc):
        _NO_WGRAD_PTRS.clear()
        _NO_WGRAD_PTRS.update(self.prev)
        return False


def _wants_wgrad(w: torch.Tensor) -> bool:
    return not _NO_WGRAD_PTRS or w.data_ptr() not in _NO_WGRAD_PTRS


# ---------------------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------------------
def _conv_args(stride: int, padding: int, transposed: bool):
    return [stride, stride], [padding, padding], [1, 1], transposed, [0, 0], 1


def _is_cl(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    """physical (B, H, W, C) view of a channels-last (B, C, H, W) tensor"""
    return t.permute(0, 2, 3, 1)


def _use_tc(x: torch.Tensor, w_shape, cfg, m_ch: int, k_ch: int) -> bool:
    """Shape / mode test shared by the three primitives.  ``m_ch`` / ``k_ch``: the channel counts that become GEMM-M
    (or N) and GEMM-K of the call."""
    stride, padding, transposed = cfg
    kh, kw = w_shape[2], w_shape[3]
    ok = (_FORCE != "cudnn" and torch.backends.cudnn.allow_tf32 and x.is_cuda and x.dtype == torch.float32
          and kh == kw and kh in (1, 3) and m_ch % 32 == 0 and k_ch % 32 == 0
          and ((not transposed and stride in (1, 2)) or (transposed and stride == 2 and padding == 0 and kh == 3))
          and x.numel() > 0 and x.shape[0] * x.shape[2] * x.shape[3] * max(m_ch, k_ch) < 2 ** 31)
    if not ok and _FORCE == "tc":
        raise RuntimeError(f"RICK_CONV_BACKEND=tc but the tcgen05 kernels do not cover conv {tuple(w_shape)} {cfg} on "
                           f"{tuple(x.shape)}")
    return ok


def _cl(t: torch.Tensor) -> torch.Tensor:
    return t if _is_cl(t) else t.contiguous(memory_format=torch.channels_last)


def _fprop(x, w, cfg, bias_act=None):
    """``bias_act`` = (bias, slope, scale): lrelu(conv + bias) * scale from the kernel's epilogue (tcgen05 path only)."""
    stride, padding, transposed = cfg
    co, ci, k, _ = w.shape
    if _use_tc(x, w.shape, cfg, co, ci):
        launch_stats["tc"] += 1
        x, w = _cl(x), _cl(w)
        b, _, h, wd = x.shape
        geom = (_ct.geom_conv_transpose_s2(b, h, wd, ci, co, k) if transposed
                else _ct.geom_conv(b, h, wd, ci, co, k, stride, padding))
        if bias_act is not None:
            bias, slope, scale = bias_act
            return _nhwc_out(_ct.conv_tc_nhwc(_nhwc(x), w, geom, bias=bias.contiguous(), act=True, alpha=slope, scale=scale))
        return _nhwc_out(_ct.conv_tc_nhwc(_nhwc(x), w, geom))
    if bias_act is not None:
        raise RuntimeError("rick_b200.conv: fused bias + activation needs the tcgen05 path")
    launch_stats["library"] += 1
    wl = w.transpose(0, 1) if transposed else w
    return torch.ops.aten.convolution(x, wl, None, *_conv_args(*cfg))


def _nhwc_out(t: torch.Tensor) -> torch.Tensor:
    """(B, H, W, C) kernel output -> logical (B, C, H, W), channels-last memory"""
    return t.permute(0, 3, 1, 2)


def _dgrad(g, w, x, cfg):
    """gradient w.r.t. the input ``x`` (only its shape is used by the tcgen05 path)"""
    stride, padding, transposed = cfg
    x_shape = x.shape
    co, ci, k, _ = w.shape
    if _use_tc(g, w.shape, cfg, ci, co):
        launch_stats["tc"] += 1
        g, w = _cl(g), _cl(w)
        b, _, h, wd = x_shape
        if transposed:      # data gradient of the transposed conv = an ordinary stride-2 conv over g
            geom = _ct.geom_conv(b, g.shape[2], g.shape[3], co, ci, k, 2, 0)
        else:
            geom = _ct.geom_conv_dgrad(b, h, wd, ci, co, k, stride, padding)
        return _nhwc_out(_ct.conv_tc_nhwc(_nhwc(g), w, geom, transpose_weight=True))
    launch_stats["library"] += 1
    wl = w.transpose(0, 1) if transposed else w
    gx, _, _ = torch.ops.aten.convolution_backward(g, x, wl, None, *_conv_args(*cfg), [True, False, False])
    return gx


def _wgrad(g, x, w, cfg):
    """gradient w.r.t. the (Cout, Cin, k, k) weight ``w`` (only its shape / strides are used)"""
    stride, padding, transposed = cfg
    co, ci, k, _ = w.shape
    if _use_tc(x, w.shape, cfg, co, ci) and _is_cl(w):
        launch_stats["tc"] += 1
        g, x = _cl(g), _cl(x)
        b, _, h, wd = x.shape
        geom = _ct.geom_wgrad(b, h, wd, ci, co, k, stride, padding, transposed)
        return _ct.conv_wgrad_tc(_nhwc(g), _nhwc(x), geom, like=w)
    launch_stats["library"] += 1
    wl = w.transpose(0, 1) if transposed else w
    _, gw, _ = torch.ops.aten.convolution_backward(g, x, wl, None, *_conv_args(*cfg), [False, True, False])
    return gw.transpose(0, 1) if transposed else gw


# ---------------------------------------------------------------------------------------------------------------
# autograd: first- and second-order structure
# ---------------------------------------------------------------------------------------------------------------
# R1 (train:462-493) and the path-length regulariser (train:546-589) differentiate THROUGH a backward pass.  Left to
# autograd's generic double-backward formula, the weight-gradient term of a convolution is evaluated as a forward
# convolution whose "filter" is a whole feature map; written out, every term of the second derivative is again one of
# the three primitives with ordinary shapes:
#     y  = fprop(x, w)          gx = dgrad(g, w)          gw = wgrad(g, x)
#     d<ggx, gx>/dg = fprop(ggx, w)      d<ggx, gx>/dw = wgrad(g, ggx)
#     d<ggw, gw>/dg = fprop(x, ggw)      d<ggw, gw>/dx = dgrad(g, ggw)
# so the two Functions below call themselves / each other and stay differentiable to any order.  ``w`` is always the
# layer's (Cout, Cin, k, k) weight, also for the transposed convolution (ModulatedConv2d's upsampling path).
class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, padding, transposed):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, transposed)
        return _fprop(x, w, ctx.cfg)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and _wants_wgrad(w)
        if not torch.is_grad_enabled():                    # plain backward (no create_graph): straight to the kernels
            gx = _dgrad(g, w, x, ctx.cfg) if need_x else None
            gw = _wgrad(g, x, w, ctx.cfg) if need_w else None
            return gx, gw, None, None, None
        gx, gw = _ConvGrad.apply(g, x, w, need_x, need_w, *ctx.cfg)
        return gx, gw, None, None, None


class _ConvGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, x, w, need_x, need_w, stride, padding, transposed):
        ctx.save_for_backward(g, x, w)
        ctx.cfg = (stride, padding, transposed)
        gx = _dgrad(g, w, x, ctx.cfg) if need_x else None
        gw = _wgrad(g, x, w, ctx.cfg) if need_w else None
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


class _ConvBiasAct(torch.autograd.Function):
    """``fused_leaky_relu(conv(x, w), bias)`` -- ConvLayer's EqualConv2d + FusedLeakyReLU pair (model_probe_tune.py:
    614-639) -- with bias and activation applied in the convolution kernel's epilogue: the activation map is written
    once instead of written, re-read and re-written by a separate bias-act launch.  Backward is composed of the two
    differentiable pieces it replaces (the bias-act backward Function on the saved OUTPUT, op/fused_act.py:19-48, and
    ``_ConvGrad``), so first and second derivatives are those of the unfused pair."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, padding, slope, scale):
        out = _fprop(x, w, (stride, padding, False), bias_act=(bias, slope, scale))
        ctx.save_for_backward(x, w, out)
        ctx.cfg = (stride, padding, False)
        ctx.act = (slope, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        from .op.fused_act import FusedLeakyReLUFunctionBackward
        x, w, out = ctx.saved_tensors
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        need_w = need_w and _wants_wgrad(w)
        g_pre, g_bias = FusedLeakyReLUFunctionBackward.apply(g, out, *ctx.act)
        gx = gw = None
        if need_x or need_w:
            if not torch.is_grad_enabled():
                gx = _dgrad(g_pre, w, x, ctx.cfg) if need_x else None
                gw = _wgrad(g_pre, x, w, ctx.cfg) if need_w else None
            else:
                gx, gw = _ConvGrad.apply(g_pre, x, w, need_x, need_w, *ctx.cfg)
        return gx, gw, (g_bias if need_b else None), None, None, None, None


def conv2d_bias_act(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, stride: int, padding: int,
                    slope: float, scale: float):
    """EqualConv2d (no bias of its own) followed by FusedLeakyReLU, as one launch when the tcgen05 kernel covers the
    shape; None otherwise (the caller then runs the two modules)."""
    if not (x.is_cuda and x.dim() == 4 and bias.dtype == torch.float32):
        return None
    if not _use_tc(x, weight.shape, (stride, padding, False), weight.shape[0], weight.shape[1]) or _FORCE == "tc-unfused":
        return None
    return _ConvBiasAct.apply(x, weight, bias, stride, padding, slope, scale)


def _lib_conv(x, w, stride: int = 1, padding: int = 0, transposed: bool = False):
    """``F.conv2d(x, w)`` or, with ``transposed``, ``F.conv_transpose2d(x, w.transpose(0, 1))`` (groups 1, no bias;
    ``w`` is (Cout, Cin, k, k) either way) with the gradient structure above."""
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
    if ci % 32 != 0 and ci > 32:
        # D's final conv has 512 + 1 (minibatch-stddev) input channels.  Zero-padding the channel axis of input and
        # weight to the next multiple of 32 (one GEMM-K block) keeps the result identical and the layer on the tcgen05
        # kernels (as a library conv an odd channel count fell back to a SIMT convolution: 0.7 ms for a 4x4 map).
        extra = 32 - ci % 32
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
            out = blur(_lib_conv(xm, w, 2, 0, True))
        else:
            out = _lib_conv(xm, w, 1, padding)
        if noise is None:
            noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
        return _styled.styled_epilogue(out, demod, noise, noise_weight, bias, slope, act_scale)
    xm = x * s[:, :, None, None]
    if upsample:
        out = _lib_conv(xm, w, 2, 0, True)
        out = blur(out)
    elif downsample:
        out = _lib_conv(blur(xm), w, 2, 0)
    else:
        out = _lib_conv(xm, w, 1, padding)
    if demod is not None:
        out = out * demod[:, :, None, None]
    return out

"""Fused stand-ins for short chains of ATen launches around the convolutions (rick_b200/csrc/glue_ops.cu).

    weight_sqsum(w)     (Cout, Cin, k, k) -> (Cout, Cin): sum over the taps of w**2, the table ModulatedConv2d's
                        demodulation is computed from (model_probe_tune.py:249-251), in one pass for any weight layout

Every op is differentiable to any order: the backward formulas are written with differentiable torch ops.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _even_taps(w: torch.Tensor) -> bool:
    kh, kw = w.shape[2], w.shape[3]
    return kh * kw == 1 or (w.stride(2) == kw * w.stride(3))


class _WeightSqSum(Function):
    @staticmethod
    def forward(ctx, w):
        co, ci, kh, kw = w.shape
        out = torch.empty(co, ci, dtype=torch.float32, device=w.device)
        with torch.cuda.device(w.device):
            _lib.check(_lib.lib().rick_weight_sqsum(out.data_ptr(), w.data_ptr(), co, ci, kh * kw, w.stride(0), w.stride(1),
                                                    w.stride(3) if kh * kw > 1 else 1, _stream()), "rick_weight_sqsum")
        ctx.save_for_backward(w)
        return out

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        return (2.0 * g)[:, :, None, None] * w           # keeps w's memory layout; differentiable again


def weight_sqsum(w: torch.Tensor) -> torch.Tensor:
    """``w.pow(2).sum([2, 3])`` for a (Cout, Cin, k, k) weight."""
    if w.is_cuda and w.dtype == torch.float32 and w.dim() == 4 and _even_taps(w):
        return _WeightSqSum.apply(w)
    return w.pow(2).sum([2, 3])


# ---------------------------------------------------------------------------------------------------------------
# many small-batch EqualLinear layers in one launch
# ---------------------------------------------------------------------------------------------------------------
import ctypes
from typing import List, Optional, Sequence


def _f32_table(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def _i32_table(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def _linear_multi_launch(latent, idx, weights, biases, w_scales, b_scales, act, alpha, act_scale, pixelnorm):
    b, n_lat, d = latent.shape
    outs = [torch.empty(b, w.shape[0], dtype=torch.float32, device=latent.device) for w in weights]
    base, row = latent.data_ptr(), n_lat * d
    with torch.cuda.device(latent.device):
        rc = _lib.lib().rick_linear_multi(
            _lib.ptr_table([o.data_ptr() for o in outs]), _lib.ptr_table([w.data_ptr() for w in weights]),
            _lib.ptr_table([None if bb is None else bb.data_ptr() for bb in biases]),
            _lib.ptr_table([base + 4 * d * i for i in idx]), _lib.i64_table([row] * len(idx)),
            _i32_table([w.shape[0] for w in weights]), _f32_table(w_scales), _f32_table(b_scales), len(weights), b, d,
            int(act), float(alpha), float(act_scale), int(pixelnorm), _stream())
    _lib.check(rc, "rick_linear_multi")
    return outs


class _LinearMulti(Function):
    """y_l = w_scale_l * latent[:, idx_l] @ W_l^T + b_scale_l * bias_l for a list of layers, one launch forward and one
    for all weight / bias gradients."""

    @staticmethod
    def forward(ctx, latent, idx, w_scales, b_scales, *wb):
        weights, biases = list(wb[0::2]), list(wb[1::2])
        ctx.idx, ctx.w_scales, ctx.b_scales = tuple(idx), tuple(w_scales), tuple(b_scales)
        ctx.n = len(weights)
        ctx.has_bias = [bb is not None for bb in biases]
        ctx.save_for_backward(latent, *weights)
        return tuple(_linear_multi_launch(latent, idx, weights, biases, w_scales, b_scales, False, 0.2, 1.0, False))

    @staticmethod
    def backward(ctx, *gys):
        latent, *weights = ctx.saved_tensors
        n, idx, ws, bs = ctx.n, ctx.idx, ctx.w_scales, ctx.b_scales
        need_lat = ctx.needs_input_grad[0]
        grads: List[Optional[torch.Tensor]] = [None] * (2 * n)
        g_lat = None
        if torch.is_grad_enabled():                      # create_graph (path-length): differentiable formulas
            per_idx = {}
            for l in range(n):
                gy = gys[l]
                if gy is None:
                    continue
                x = latent[:, idx[l]]
                if ctx.needs_input_grad[4 + 2 * l]:
                    grads[2 * l] = ws[l] * (gy.t() @ x)
                if ctx.has_bias[l] and ctx.needs_input_grad[5 + 2 * l]:
                    grads[2 * l + 1] = bs[l] * gy.sum(0)
                if need_lat:
                    t = ws[l] * (gy @ weights[l])
                    per_idx[idx[l]] = t if idx[l] not in per_idx else per_idx[idx[l]] + t
            if need_lat:
                zero = latent.new_zeros(latent.shape[0], latent.shape[2])
                g_lat = torch.stack([per_idx.get(i, zero) for i in range(latent.shape[1])], 1)
            return (g_lat, None, None, None) + tuple(grads)
        live = [l for l in range(n) if gys[l] is not None]
        if live:
            gw = {l: (torch.empty_like(weights[l]) if ctx.needs_input_grad[4 + 2 * l] else None) for l in live}
            gb = {l: (torch.empty(weights[l].shape[0], dtype=torch.float32, device=latent.device)
                      if ctx.has_bias[l] and ctx.needs_input_grad[5 + 2 * l] else None) for l in live}
            gyc = {l: gys[l].contiguous() for l in live}
            b, n_lat, d = latent.shape
            base, row = latent.data_ptr(), n_lat * d
            with torch.cuda.device(latent.device):
                rc = _lib.lib().rick_linear_multi_wgrad(
                    _lib.ptr_table([None if gw[l] is None else gw[l].data_ptr() for l in live]),
                    _lib.ptr_table([None if gb[l] is None else gb[l].data_ptr() for l in live]),
                    _lib.ptr_table([gyc[l].data_ptr() for l in live]), _lib.ptr_table([base + 4 * d * idx[l] for l in live]),
                    _lib.i64_table([row] * len(live)), _i32_table([weights[l].shape[0] for l in live]),
                    _f32_table([ws[l] for l in live]), _f32_table([bs[l] for l in live]), len(live), b, d, _stream())
            _lib.check(rc, "rick_linear_multi_wgrad")
            for l in live:
                grads[2 * l], grads[2 * l + 1] = gw[l], gb[l]
            if need_lat:                                  # only the path-length / Fisher passes differentiate the latents
                g_lat = torch.zeros_like(latent)
                for l in live:
                    g_lat[:, idx[l]].addmm_(gyc[l], weights[l], alpha=ws[l])
        return (g_lat, None, None, None) + tuple(grads)


def linear_multi(latent: torch.Tensor, idx: Sequence[int], weights: Sequence[torch.Tensor],
                 biases: Sequence[Optional[torch.Tensor]], w_scales: Sequence[float], b_scales: Sequence[float]):
    """[w_scales[l] * latent[:, idx[l]] @ weights[l].T + b_scales[l] * biases[l] for l ...] -- the modulation layers of
    all ModulatedConv2d modules of a generator in one launch.  ``latent`` is (B, n_latent, D) contiguous fp32."""
    wb = []
    for w, bb in zip(weights, biases):
        wb += [w, bb]
    return list(_LinearMulti.apply(latent, tuple(idx), tuple(w_scales), tuple(b_scales), *wb))


def linear_multi_ok(latent: torch.Tensor, weights: Sequence[torch.Tensor]) -> bool:
    return (latent.is_cuda and latent.dtype == torch.float32 and latent.dim() == 3 and latent.is_contiguous()
            and latent.shape[0] <= 8 and latent.shape[2] % 4 == 0
            and all(w.is_contiguous() and w.dtype == torch.float32 and w.shape[1] == latent.shape[2] for w in weights))


@torch.no_grad()
def mapping_network(z: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], w_scale: float,
                    b_scale: float, negative_slope: float = 0.2, act_scale: float = 2 ** 0.5) -> torch.Tensor:
    """PixelNorm + the chain of EqualLinear(activation='fused_lrelu') layers (model_probe_tune.py:383-392), forward only:
    one launch per layer (the normalisation rides on the first)."""
    x = z.contiguous().unsqueeze(1)                       # (B, 1, D)
    for i, (w, bb) in enumerate(zip(weights, biases)):
        (y,) = _linear_multi_launch(x, [0], [w], [bb], [w_scale], [b_scale], True, negative_slope, act_scale, i == 0)
        x = y.unsqueeze(1)
    return x.squeeze(1)


# ---------------------------------------------------------------------------------------------------------------
# the discriminator's from-RGB layer (1x1 conv from <= 4 channels + bias + leaky-ReLU) in one pass each way
# ---------------------------------------------------------------------------------------------------------------
class _FromRGB(Function):
    @staticmethod
    def forward(ctx, img, weight, bias, w_scale, alpha, act_scale):
        b, cin, h, w = img.shape
        cout = weight.shape[0]
        img_c = img.contiguous()
        w2 = weight.reshape(cout, cin).contiguous()
        y = torch.empty((b, h, w, cout), dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().rick_from_rgb_fwd(y.data_ptr(), img_c.data_ptr(), w2.data_ptr(),
                                                    None if bias is None else bias.data_ptr(), b, h * w, cin, cout,
                                                    float(w_scale), 1, float(alpha), float(act_scale), _stream()),
                       "rick_from_rgb_fwd")
        ctx.save_for_backward(img_c, weight, y)
        ctx.cfg = (float(w_scale), float(alpha), float(act_scale), bias is not None)
        return y.permute(0, 3, 1, 2)                       # logical NCHW, channels-last memory

    @staticmethod
    def backward(ctx, g):
        img, weight, y = ctx.saved_tensors
        w_scale, alpha, act_scale, has_bias = ctx.cfg
        need_img, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and has_bias
        from .. import conv as _conv
        if not _conv._wants_wgrad(weight):                 # a backward restricted to other tensors (Fisher round, G's loss)
            need_w = need_b = False
        b, cin, h, w = img.shape
        cout = weight.shape[0]
        w2 = weight.reshape(cout, cin)
        if torch.is_grad_enabled() or need_w or need_b:
            # create_graph (R1 differentiates d D / d image again) or parameter gradients (Fisher round only: the layer
            # is not among the trained parameters, train:921-931): differentiable formulas
            t = torch.where(y > 0, g.permute(0, 2, 3, 1), g.permute(0, 2, 3, 1) * alpha) * act_scale      # (B, H, W, Co)
            g_img = torch.einsum("bhwo,oc->bchw", t, w2) * w_scale if need_img else None
            g_w = (torch.einsum("bhwo,bchw->oc", t, img) * w_scale).reshape(weight.shape) if need_w else None
            g_b = t.sum((0, 1, 2)) if need_b else None
            return g_img, g_w, g_b, None, None, None
        g_img = None
        if need_img:
            gc = g.permute(0, 2, 3, 1).contiguous()         # a view when g is channels-last
            g_img = torch.empty_like(img)
            with torch.cuda.device(img.device):
                _lib.check(_lib.lib().rick_from_rgb_bwd_data(g_img.data_ptr(), gc.data_ptr(), y.data_ptr(),
                                                             w2.contiguous().data_ptr(), b, h * w, cin, cout, w_scale, 1,
                                                             alpha, act_scale, _stream()), "rick_from_rgb_bwd_data")
        return g_img, None, None, None, None, None


def from_rgb_ok(img: torch.Tensor, weight: torch.Tensor) -> bool:
    return (img.is_cuda and img.dtype == torch.float32 and img.dim() == 4 and weight.dim() == 4 and weight.shape[1] <= 4
            and weight.shape[2] == weight.shape[3] == 1 and weight.shape[0] % 4 == 0 and img.shape[1] == weight.shape[1])


def from_rgb(img, weight, bias, w_scale: float, negative_slope: float = 0.2, act_scale: float = 2 ** 0.5):
    """``fused_leaky_relu(conv2d(img, weight * w_scale), bias)`` for a 1x1 convolution from <= 4 channels; the result is a
    logical (B, Cout, H, W) tensor in channels-last memory."""
    return _FromRGB.apply(img, weight, bias, w_scale, negative_slope, act_scale)


# ---------------------------------------------------------------------------------------------------------------
# demodulation table of a ModulatedConv2d and the residual merge of a ResBlock
# ---------------------------------------------------------------------------------------------------------------
def _demod_composite(s, wsq, scale2: float, eps: float, s_scale: float):
    """The same quantities as module code (model_probe_tune.py:246-251, algebraic form)."""
    return torch.rsqrt(torch.nn.functional.linear(s.pow(2), wsq) * scale2 + eps), s * s_scale


class _Demod(Function):
    """(demod, s * s_scale) of one modulated convolution from its style ``s`` (B, Cin) and tap-summed squared weight
    ``wsq`` (Cout, Cin): one launch forward, two for the first derivatives.  When the backward pass is itself being
    differentiated (path-length regularisation, train:104-118) the gradient is re-derived with differentiable torch ops,
    so higher derivatives are those of the module code."""

    @staticmethod
    def forward(ctx, s, wsq, scale2, eps, s_scale):
        b, cin = s.shape
        cout = wsq.shape[0]
        demod = torch.empty(b, cout, dtype=torch.float32, device=s.device)
        s_out = torch.empty_like(s)
        with torch.cuda.device(s.device):
            _lib.check(_lib.lib().rick_demod_fwd(demod.data_ptr(), s_out.data_ptr(), s.data_ptr(), wsq.data_ptr(), b, cin,
                                                 cout, float(scale2), float(eps), float(s_scale), _stream()), "rick_demod_fwd")
        ctx.save_for_backward(s, wsq, demod)
        ctx.cfg = (float(scale2), float(eps), float(s_scale))
        return demod, s_out

    @staticmethod
    def backward(ctx, g_demod, g_sout):
        s, wsq, demod = ctx.saved_tensors
        scale2, eps, s_scale = ctx.cfg
        need_s, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if torch.is_grad_enabled():                      # double backward requested: differentiable composite
            with torch.enable_grad():
                s_ = s.detach().requires_grad_(True) if not s.requires_grad else s
                w_ = wsq.detach().requires_grad_(True) if not wsq.requires_grad else wsq
                d, so = _demod_composite(s_, w_, scale2, eps, s_scale)
                outs, gouts = [], []
                if g_demod is not None:
                    outs.append(d), gouts.append(g_demod)
                if g_sout is not None:
                    outs.append(so), gouts.append(g_sout)
                gs, gw = torch.autograd.grad(outs, (s_, w_), gouts, create_graph=True, allow_unused=True)
            return (gs if need_s else None), (gw if need_w else None), None, None, None
        b, cin = s.shape
        cout = wsq.shape[0]
        if g_demod is None:
            return (g_sout * s_scale if need_s and g_sout is not None else None), None, None, None, None
        g_demod = g_demod.contiguous()
        g_sout = None if g_sout is None else g_sout.contiguous()
        gs = torch.empty_like(s) if need_s else None
        gw = torch.empty_like(wsq) if need_w else None
        with torch.cuda.device(s.device):
            _lib.check(_lib.lib().rick_demod_bwd(None if gs is None else gs.data_ptr(), None if gw is None else gw.data_ptr(),
                                                 g_demod.data_ptr(), None if g_sout is None else g_sout.data_ptr(),
                                                 demod.data_ptr(), s.data_ptr(), wsq.data_ptr(), b, cin, cout, scale2,
                                                 s_scale, _stream()), "rick_demod_bwd")
        return gs, gw, None, None, None


import os as _os

_FUSED_GLUE = _os.environ.get("RICK_FUSED_GLUE", "1") != "0"     # A/B switch for the demod / residual-merge kernels


def demod_ok(s: torch.Tensor, wsq: torch.Tensor) -> bool:
    return (_FUSED_GLUE and s.is_cuda and s.dtype == torch.float32 and s.dim() == 2 and s.shape[0] <= 8 and s.is_contiguous()
            and wsq.dtype == torch.float32 and wsq.is_contiguous() and wsq.shape[1] == s.shape[1]
            and s.shape[0] * wsq.shape[0] * 4 <= 40 * 1024)


def demod(s: torch.Tensor, wsq: torch.Tensor, scale2: float, eps: float, s_scale: float):
    """(rsqrt(scale2 * s^2 @ wsq^T + eps), s * s_scale)."""
    if demod_ok(s, wsq):
        return _Demod.apply(s, wsq, scale2, eps, s_scale)
    return _demod_composite(s, wsq, scale2, eps, s_scale)


class _AddScale(Function):
    @staticmethod
    def forward(ctx, a, b, scale):
        out = torch.empty_like(a)
        with torch.cuda.device(a.device):
            _lib.check(_lib.lib().rick_add_scale(out.data_ptr(), a.data_ptr(), b.data_ptr(), float(scale), a.numel(),
                                                 _stream()), "rick_add_scale")
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, g):
        gs = g * ctx.scale                               # one pass, shared by both branches; differentiable again
        return gs, gs, None


def add_scale(a: torch.Tensor, b: torch.Tensor, scale: float) -> torch.Tensor:
    """``(a + b) * scale`` (ResBlock's residual merge, model_probe_tune.py:655-660) as one pass when both operands are
    dense float32 CUDA tensors with identical strides."""
    if (_FUSED_GLUE and a.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape
            and a.stride() == b.stride() and a.numel() % 4 == 0 and a.numel() > 0
            and (a.is_contiguous() or a.is_contiguous(memory_format=torch.channels_last))
            and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0):
        return _AddScale.apply(a, b, scale)
    return (a + b) * scale

"""Fused training-time ops around the differentiated modulated convolution (channels-last fp32):

    modulate(x, s)                                          y = x * s[b, c]
    styled_epilogue(a, demod, noise, noise_weight, bias)    y = lrelu(a * demod[b,c] + noise_weight * noise + bias[c]) * scale

i.e. ModulatedConv2d's modulation / demodulation in the algebraic form (model_probe_tune.py:246-251), NoiseInjection
(:293-298) and FusedLeakyReLU (op/fused_act.py) -- each as ONE kernel forward and ONE kernel backward (element-wise part
plus all broadcast-gradient reductions), instead of the ~10 broadcast / reduce kernels per layer autograd would launch.

Second derivatives (path-length regularisation differentiates through G twice): when ``backward`` runs with grad mode
enabled (``create_graph=True``) it is itself a differentiable fused op (``_ModulateBwd``, the bias-act backward Function)
whose derivatives are again these kernels, so autograd can differentiate to any order without leaving them.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def fused_ok(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0 and x.numel() > 0


def _cl(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous(memory_format=torch.channels_last)


def _ws(b, hw, c, device):
    n = max(int(_lib.lib().rick_colsum_workspace(b, hw, c)), 4)
    return torch.empty(n, dtype=torch.uint8, device=device)


def _modulate_bwd_launch(gy, x, s):
    b, c, h, w = x.shape
    gx = torch.empty_like(x, memory_format=torch.channels_last)
    gs = torch.empty(b, c, dtype=torch.float32, device=x.device)
    ws = _ws(b, h * w, c, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().rick_modulate_bwd_nhwc(gx.data_ptr(), gs.data_ptr(), ws.data_ptr(), gy.data_ptr(),
                                                     x.data_ptr(), s.data_ptr(), b, h * w, c, _stream()),
                   "rick_modulate_bwd_nhwc")
    return gx, gs


class _ModulateBwd(Function):
    """(gy * s[b, c], sum_hw gy * x) -- the backward of ``modulate`` -- as a differentiable op of its own, ONE launch.
    Its derivatives are again modulate / modulate-backward calls, so the path-length regulariser's second backward runs on
    the fused kernels to any order (round 2: the composite torch formulas it replaces were 2.6 ms of the 11 ms path-length
    sub-step -- spatial sums of channels-last tensors are ATen's slow case, 78 us for a 33 MB tensor)."""

    @staticmethod
    def forward(ctx, gy, x, s):
        ctx.save_for_backward(gy, x, s)
        return _modulate_bwd_launch(gy, x, s)

    @staticmethod
    def backward(ctx, ggx, ggs):
        gy, x, s = ctx.saved_tensors
        need_gy, need_x, need_s = ctx.needs_input_grad
        d_gy = d_x = d_s = None
        if ggx is not None and (need_gy or need_s):
            t1, d_s = _ModulateBwd.apply(_cl(ggx), gy, s)          # ggx * s and sum_hw ggx * gy in one launch
            d_gy = t1
        if ggs is not None:
            ggs = ggs.contiguous()
            if need_gy:
                t2 = _Modulate.apply(x, ggs)
                d_gy = t2 if d_gy is None else d_gy + t2
            if need_x:
                d_x = _Modulate.apply(gy, ggs)
        return (d_gy if need_gy else None), d_x, (d_s if need_s else None)


class _Modulate(Function):
    @staticmethod
    def forward(ctx, x, s):
        # x is channels-last and s contiguous (the wrapper converts OUTSIDE the Function, differentiably): the saved
        # tensors must be the inputs themselves or the double-backward branch would treat them as constants
        b, c, h, w = x.shape
        y = torch.empty_like(x, memory_format=torch.channels_last)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().rick_modulate_nhwc(y.data_ptr(), x.data_ptr(), s.data_ptr(), b, h * w, c, _stream()),
                       "rick_modulate_nhwc")
        ctx.save_for_backward(x, s)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, s = ctx.saved_tensors
        if torch.is_grad_enabled():                       # double backward requested: the differentiable fused op
            return _ModulateBwd.apply(_cl(gy), x, s)
        return _modulate_bwd_launch(_cl(gy), x, s)


class _StyledEpilogue(Function):
    @staticmethod
    def forward(ctx, a, demod, noise, noise_weight, bias, alpha, scale):
        # a channels-last, demod / noise (B, HW) / bias contiguous: converted by the wrapper (see _Modulate.forward)
        b, c, h, w = a.shape
        y = torch.empty_like(a, memory_format=torch.channels_last)
        with torch.cuda.device(a.device):
            _lib.check(_lib.lib().rick_styled_epilogue_nhwc(y.data_ptr(), a.data_ptr(), demod.data_ptr(), noise.data_ptr(),
                                                            noise_weight.data_ptr(), bias.data_ptr(), b, h * w, c,
                                                            float(alpha), float(scale), _stream()),
                       "rick_styled_epilogue_nhwc")
        ctx.save_for_backward(a, demod, noise, noise_weight, y)
        ctx.alpha, ctx.scale = alpha, scale
        return y

    @staticmethod
    def backward(ctx, gy):
        a, demod, noise, noise_weight, y = ctx.saved_tensors
        b, c, h, w = a.shape
        if torch.is_grad_enabled():
            # double backward requested: the same quantities from two differentiable fused ops -- the activation gradient
            # (t, sum t) from the bias-act backward (op/fused_act.py:19-48) and (t * demod, sum_hw t * a) from the
            # modulate backward; only the scalar noise-weight gradient stays a torch expression
            from .fused_act import FusedLeakyReLUFunctionBackward
            t, gb = FusedLeakyReLUFunctionBackward.apply(_cl(gy), y, ctx.alpha, ctx.scale)
            ga, gd = _ModulateBwd.apply(t, a, demod)
            gnw = (t.sum(1).reshape(b, h * w) * noise).sum().reshape(1)
            return ga, gd, None, gnw, gb, None, None
        gy = _cl(gy)
        ga = torch.empty_like(a, memory_format=torch.channels_last)
        gd = torch.empty(b, c, dtype=torch.float32, device=a.device)
        gb = torch.empty(c, dtype=torch.float32, device=a.device)
        gnw = torch.empty(1, dtype=torch.float32, device=a.device)
        ws = _ws(b, h * w, c, a.device)
        with torch.cuda.device(a.device):
            _lib.check(_lib.lib().rick_styled_epilogue_bwd_nhwc(ga.data_ptr(), gd.data_ptr(), gb.data_ptr(), gnw.data_ptr(),
                                                                ws.data_ptr(), gy.data_ptr(), y.data_ptr(), a.data_ptr(),
                                                                demod.data_ptr(), noise.data_ptr(), b, h * w, c,
                                                                float(ctx.alpha), float(ctx.scale), _stream()),
                       "rick_styled_epilogue_bwd_nhwc")
        return ga, gd, None, gnw, gb, None, None


def modulate(x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """x * s[:, :, None, None] (channels-last result)."""
    if fused_ok(x):
        return _Modulate.apply(_cl(x), s.contiguous())
    return x * s[:, :, None, None]


def styled_epilogue(a, demod, noise, noise_weight, bias, negative_slope=0.2, scale=2 ** 0.5):
    """lrelu(a * demod[b,c] + noise_weight * noise + bias[c]) * scale; ``noise`` is (B or 1, 1, H, W)."""
    if fused_ok(a):
        b, c, h, w = a.shape
        noise = noise.detach().expand(b, 1, h, w).reshape(b, h * w).contiguous()      # noise is data, not a parameter
        return _StyledEpilogue.apply(_cl(a), demod.contiguous(), noise, noise_weight, bias.contiguous(), negative_slope,
                                     scale)
    from .fused_act import fused_leaky_relu
    return fused_leaky_relu(a * demod[:, :, None, None] + noise_weight * noise, bias, negative_slope, scale)

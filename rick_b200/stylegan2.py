"""StyleGAN2 generator / discriminator modules with the reference's names, constructor signatures, return tuples
and state_dict keys (gan_training/models/model_probe_tune.py), so reference checkpoints load unchanged and the
Fisher step's string keys (train_dynamic_update_prune.py:282-351) resolve.

What is different underneath:
  * ``upfirdn2d`` / ``fused_leaky_relu`` are the sm_100a kernels of this package (rick_b200/op).
  * ``ModulatedConv2d`` never materialises per-sample weights.  It uses the algebraic form
        out[b] = demod[b, :, None, None] * conv(x[b] * s[b, :, None, None], scale * W)
        demod[b, co] = rsqrt( sum_ci s[b, ci]^2 * sum_{kh,kw} (scale * W[co, ci])^2 + 1e-8 )
    (identical to model_probe_tune.py:246-251 up to float rounding), so the batch folds into one dense convolution
    on the shared weight instead of ``groups = batch`` convolutions on B x 9.4 MB of weights.
  * ``Discriminator.forward`` does not evaluate conv1 / conv2 of every block twice; the reference does so only to
    fill a ``feat`` list the trainer throws away (model_probe_tune.py:740-744, train:407-410).  ``feat`` is still
    returned, filled from the single evaluation.
The dense convolutions themselves go through ``rick_b200.conv`` (tcgen05 implicit GEMM where available, see
DESIGN.md for the per-path status).
"""
from __future__ import annotations

import math
import os
import random
from typing import List, Optional

import torch
from torch import autograd, nn
from torch.nn import functional as F

from . import conv as _conv
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d
from .op import glue as _glue
from .op.glue import weight_sqsum
from .op.scale import scale_all


# Activation memory format inside G / D.  Channels-last keeps the library convolutions in their native NHWC kernels
# (no nchw<->nhwc transposes around every conv: ~15 % of the iteration in the round-1 launch list) and routes the
# custom ops to their NHWC kernels.  Tensor SHAPES stay (N, C, H, W) either way.
_CHANNELS_LAST = os.environ.get("RICK_CHANNELS_LAST", "1") != "0"


_PRESCALE = os.environ.get("RICK_PRESCALE", "1") != "0"           # one multi-tensor launch for all equalised-lr multipliers
_FUSED_STYLED = os.environ.get("RICK_FUSED_STYLED", "1") != "0"   # fused modulate / demod-noise-bias-act ops in StyledConv
_FUSED_LINEARS = os.environ.get("RICK_FUSED_LINEARS", "1") != "0"  # mapping network / all modulation layers: one launch each
_FUSED_CONV_ACT = os.environ.get("RICK_FUSED_CONV_ACT", "1") != "0"  # D: EqualConv2d + FusedLeakyReLU as one launch


def set_fused_styled(flag: bool) -> None:
    global _FUSED_STYLED
    _FUSED_STYLED = bool(flag)


def set_channels_last(flag: bool) -> None:
    global _CHANNELS_LAST
    _CHANNELS_LAST = bool(flag)


def _fmt(x: torch.Tensor) -> torch.Tensor:
    if _CHANNELS_LAST and x.is_cuda and x.dim() == 4:
        return x.contiguous(memory_format=torch.channels_last)
    return x


def make_kernel(k):
    k = torch.as_tensor(k, dtype=torch.float32)
    if k.dim() == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def _fir_pads(taps: int, factor: int, kernel_size: int, mode: str):
    """Padding arithmetic of Upsample / Downsample / Blur-around-conv (model_probe_tune.py:48-53, 69-74, 209-223, 608-612)."""
    if mode == "up":        # Upsample module
        p = taps - factor
        return (p + 1) // 2 + factor - 1, p // 2
    if mode == "down":      # Downsample module
        p = taps - factor
        return (p + 1) // 2, p // 2
    if mode == "conv_up":   # blur after a stride-2 transposed conv
        p = (taps - factor) - (kernel_size - 1)
        return (p + 1) // 2 + factor - 1, p // 2 + 1
    p = (taps - factor) + (kernel_size - 1)   # "conv_down": blur before a stride-2 conv
    return (p + 1) // 2, p // 2


class PixelNorm(nn.Module):
    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


class Upsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel) * (factor ** 2))
        self.pad = _fir_pads(self.kernel.shape[0], factor, 1, "up")

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel))
        self.pad = _fir_pads(self.kernel.shape[0], factor, 1, "down")

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


def declare_prescale_groups(net: nn.Module, *param_groups):
    """Turn on one-launch equalised-lr scaling for ``net`` (rick_b200.op.scale.scale_all).  Each group of parameters is
    scaled by its own autograd node, so a backward restricted to one group (``autograd.backward(loss, inputs=group)``,
    as the adaptation loop does for the trainable subset) still skips every gradient of the others -- with a single
    node autograd would have to produce ALL weight gradients (e.g. the K = 65536 weight gradient of each ToRGB) before
    it could run.  Parameters in no group form a last group of their own."""
    plan = []
    for m in net.modules():
        if isinstance(m, EqualConv2d):
            plan.append((m, "_ws", m.weight, m.scale))
        elif isinstance(m, EqualLinear):
            plan.append((m, "_ws", m.weight, m.scale))
            if m.bias is not None:
                plan.append((m, "_bs", m.bias, m.lr_mul))
    groups, taken = [], set()
    for params in param_groups:
        ids = {id(p) for p in params}
        grp = [e for e in plan if id(e[2]) in ids and id(e[2]) not in taken]
        taken |= {id(e[2]) for e in grp}
        if grp:
            groups.append(grp)
    rest = [e for e in plan if id(e[2]) not in taken]
    if rest:
        groups.append(rest)
    object.__setattr__(net, "_prescale_groups", groups)


class _Prescaled:
    """``with _Prescaled(net):`` -- every EqualConv2d / EqualLinear of ``net`` finds its equalised-lr-scaled weight
    (and bias) ready for the pass: one multi-tensor launch per declared group instead of one element-wise launch per
    module, forward and backward.  Same values, same gradients.  A no-op until declare_prescale_groups(net, ...)."""

    def __init__(self, net: nn.Module):
        self.groups = getattr(net, "_prescale_groups", None) if _PRESCALE else None
        self.ptrs = set()                                  # data pointers of the scaled weights of this pass
        object.__setattr__(net, "_scaled_weight_ptrs", self.ptrs)

    def __enter__(self):
        if self.groups and self.groups[0][0][2].is_cuda:
            for grp in self.groups:
                outs = scale_all([p for _, _, p, _ in grp], [s for *_, s in grp])
                for (m, slot, _, _), o in zip(grp, outs):
                    setattr(m, slot, o)
                    self.ptrs.add(o.data_ptr())
        return self

    def __exit__(self, *exc):
        for grp in self.groups or ():
            for m, slot, _, _ in grp:
                setattr(m, slot, None)
        return False


class EqualConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    _ws = None                                             # scaled weight handed in by _Prescaled for one pass

    def forward(self, input):
        w = self._ws if self._ws is not None else self.weight * self.scale
        return _conv.conv2d(input, w, self.bias, stride=self.stride, padding=self.padding)

    def __repr__(self):
        o, i, k, _ = self.weight.shape
        return f"{self.__class__.__name__}({i}, {o}, {k}, stride={self.stride}, padding={self.padding})"


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    _ws = None                                             # scaled weight / bias handed in by _Prescaled
    _bs = None

    def forward(self, input):
        w = self._ws if self._ws is not None else self.weight * self.scale
        b = self._bs if self._bs is not None else (self.bias * self.lr_mul if self.bias is not None else None)
        if self.activation:
            return fused_leaky_relu(F.linear(input, w), b)
        return F.linear(input, w, bias=b)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class ScaledLeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return F.leaky_relu(input, negative_slope=self.negative_slope) * math.sqrt(2)


class ModulatedConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            self.blur = Blur(blur_kernel, pad=_fir_pads(len(blur_kernel), 2, kernel_size, "conv_up"), upsample_factor=2)
        if downsample:
            self.blur = Blur(blur_kernel, pad=_fir_pads(len(blur_kernel), 2, kernel_size, "conv_down"))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        w = torch.randn(1, out_channel, in_channel, kernel_size, kernel_size)
        if _CHANNELS_LAST:
            # same shape / state_dict entry as the reference; MEMORY is [Cout][k][k][Cin] so that the tcgen05 kernels
            # read it in place as a K-major (forward) or MN-major (data gradient) operand and write its gradient there
            w = w.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        self.weight = nn.Parameter(w)
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")

    def forward(self, input, style, epilogue=None, s=None):
        """``epilogue`` = (noise, noise_weight, bias, slope, scale): StyledConv hands its NoiseInjection + FusedLeakyReLU
        down so that demodulation, noise, bias and activation run as one fused op behind the convolution.
        ``s``: this layer's modulation ``self.modulation(style)`` when the generator computed all of them in one launch."""
        if s is None:
            s = self.modulation(style)                               # (B, Cin)
        # conv(x * s, scale * W) == conv(x * (scale * s), W): the equalised-lr scale rides on the (B, Cin) style, so
        # the 2.4 M-float weight is not rescaled (one multiply kernel + one in backward per layer and call)
        w = self.weight.squeeze(0)                                   # (Cout, Cin, k, k), shared by the batch
        demod = None
        if self.demodulate:
            wsq = weight_sqsum(w)                                    # (Cout, Cin), one pass over the weight
            # (B, Cout) demodulation and the scaled style in one launch (pow / mm / mul / add / rsqrt / mul as module code)
            demod, s = _glue.demod(s.contiguous(), wsq, self.scale ** 2, self.eps, self.scale)
        else:
            s = s * self.scale
        return _conv.modulated_conv2d(input, w, s, demod, upsample=self.upsample, downsample=self.downsample,
                                      padding=self.padding, blur=getattr(self, "blur", None), epilogue=epilogue)


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None, s=None):
        if _FUSED_STYLED and _conv._styled.fused_ok(input) and self.conv.demodulate and self.conv.out_channel % 4 == 0:
            return self.conv(input, style, epilogue=(noise, self.noise.weight, self.activate.bias,
                                                     self.activate.negative_slope, self.activate.scale), s=s)
        out = self.conv(input, style, s=s)
        out = self.noise(out, noise=noise)
        return self.activate(out)


class ToRGB(nn.Module):
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None, s=None):
        out = self.conv(input, style, s=s) + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return out


def _estimate_fisher(module: nn.Module, loglikelihood):
    """Shared body of Generator/Discriminator.estimate_fisher (model_probe_tune.py:481-504, 706-729)."""
    names = [n for n, p in module.named_parameters()]
    grads = autograd.grad(loglikelihood, list(module.parameters()), retain_graph=True)
    info = {n: g.detach() ** 2 for n, g, p in zip(names, grads, module.parameters()) if p.requires_grad}
    return grads, info


_CHANNELS = lambda cm: {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm,  # noqa: E731
                        512: 32 * cm, 1024: 16 * cm}


class Generator(nn.Module):
    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        self.style = nn.Sequential(PixelNorm(), *[EqualLinear(style_dim, style_dim, lr_mul=lr_mlp,
                                                              activation="fused_lrelu") for _ in range(n_mlp)])
        self.channels = _CHANNELS(channel_multiplier)
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        in_channel = self.channels[4]
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2

    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            noises += [torch.randn(1, 1, 2 ** i, 2 ** i, device=device) for _ in range(2)]
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def map_latent(self, z):
        """``self.style(z)``.  Without autograd (every call of the adaptation loop: the mapping network is not among the
        trained parameters, train:908-917) the PixelNorm + 8 x EqualLinear chain runs as one launch per layer."""
        if (_FUSED_LINEARS and z.is_cuda and z.dtype == torch.float32 and z.dim() == 2 and z.shape[0] <= 8
                and not torch.is_grad_enabled()):
            lin = [m for m in self.style if isinstance(m, EqualLinear)]
            if all(m.activation and m.bias is not None and m.weight.is_contiguous() for m in lin):
                return _glue.mapping_network(z, [m.weight for m in lin], [m.bias for m in lin], lin[0].scale, lin[0].lr_mul)
        return self.style(z)

    def _modulated(self):
        """(ModulatedConv2d, latent index) in forward order: conv1, to_rgb1, then (up conv, conv, to_rgb) per block
        (model_probe_tune.py:567-582)."""
        plan = [(self.conv1.conv, 0), (self.to_rgb1.conv, 1)]
        i = 1
        for up_conv, conv, rgb in zip(self.convs[::2], self.convs[1::2], self.to_rgbs):
            plan += [(up_conv.conv, i), (conv.conv, i + 1), (rgb.conv, i + 2)]
            i += 2
        return plan

    def all_modulations(self, latent):
        """Every layer's ``modulation(latent[:, i])`` from TWO launches -- the StyledConv layers and the ToRGB layers --
        (and one each for their weight gradients), or None when the fused path does not apply.  Returned in the order
        of ``_modulated()``.  Two autograd nodes, not one: the adaptation loop trains the StyledConv modulations but not
        the ToRGB ones (train:908-917), and with a single node a backward restricted to the trained subset would still
        have to produce d loss / d s of every ToRGB -- a (3 x Cin) x 65536-pixel weight-gradient GEMM per layer that
        the restricted backward otherwise prunes (1.9 ms per iteration when it was not)."""
        plan = self._modulated()
        mods = [m.modulation for m, _ in plan]
        if not (_FUSED_LINEARS and latent.dim() == 3 and _glue.linear_multi_ok(latent, [m.weight for m in mods])
                and all(m.bias is not None and not m.activation for m in mods)):
            return None
        out = [None] * len(plan)
        for want_rgb in (False, True):
            sel = [k for k, (m, _) in enumerate(plan) if (m.out_channel == 3 and m.kernel_size == 1) == want_rgb]
            res = _glue.linear_multi(latent, [plan[k][1] for k in sel], [mods[k].weight for k in sel],
                                     [mods[k].bias for k in sel], [mods[k].scale for k in sel],
                                     [mods[k].lr_mul for k in sel])
            for k, r in zip(sel, res):
                out[k] = r
        return out

    def estimate_fisher(self, loglikelihood):
        return _estimate_fisher(self, loglikelihood)

    def forward(self, *args, **kwargs):
        if _FUSED_LINEARS and self.input.input.is_cuda:       # every EqualLinear of G is served by the fused launches
            return self._forward(*args, **kwargs)
        with _Prescaled(self):
            return self._forward(*args, **kwargs)

    def _forward(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                 input_is_latent=False, noise=None, randomize_noise=True, return_feats=False):
        if not input_is_latent:
            styles = [self.map_latent(s) for s in styles]
        if noise is None:
            noise = ([None] * self.num_layers if randomize_noise
                     else [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)])
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1) if styles[0].ndim < 3 else styles[0]
        else:
            if inject_index is None:
                inject_index = random.randint(1, self.n_latent - 1)
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)

        feats: List[torch.Tensor] = []
        latent = latent.contiguous()
        mods = self.all_modulations(latent) if latent.is_cuda else None
        sm = iter(mods) if mods is not None else iter(())
        nxt = (lambda: next(sm)) if mods is not None else (lambda: None)
        out = _fmt(self.input(latent))
        out = self.conv1(out, latent[:, 0], noise=noise[0], s=nxt())
        feats.append(out)
        skip = self.to_rgb1(out, latent[:, 1], s=nxt())
        i = 1
        for up_conv, conv, n1, n2, rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2], self.to_rgbs):
            out = up_conv(out, latent[:, i], noise=n1, s=nxt())
            feats.append(out)
            out = conv(out, latent[:, i + 1], noise=n2, s=nxt())
            feats.append(out)
            skip = rgb(out, latent[:, i + 2], skip, s=nxt())
            i += 2
        image = skip
        if return_latents:
            return image, latent
        if return_feats:
            return image, feats
        return image, None


class ConvLayer(nn.Sequential):
    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1], bias=True,
                 activate=True):
        layers = []
        if downsample:
            layers.append(Blur(blur_kernel, pad=_fir_pads(len(blur_kernel), 2, kernel_size, "conv_down")))
            stride, self.padding = 2, 0
        else:
            stride, self.padding = 1, kernel_size // 2
        layers.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=stride,
                                  bias=bias and not activate))
        if activate:
            layers.append(FusedLeakyReLU(out_channel) if bias else ScaledLeakyReLU(0.2))
        super().__init__(*layers)

    def forward(self, input):
        """Same modules, same order; an EqualConv2d directly followed by FusedLeakyReLU runs as ONE convolution launch
        whose epilogue adds the bias and applies the activation (rick_b200.conv.conv2d_bias_act)."""
        mods = list(self)
        x = input
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if (_FUSED_CONV_ACT and isinstance(m, EqualConv2d) and m.bias is None and isinstance(nxt, FusedLeakyReLU)
                    and x.is_cuda and m.weight.shape[1] % 32 == 0):
                w = m._ws if m._ws is not None else m.weight * m.scale
                y = _conv.conv2d_bias_act(x, w, nxt.bias, m.stride, m.padding, nxt.negative_slope, nxt.scale)
                if y is not None:
                    x = y
                    i += 2
                    continue
            x = m(x)
            i += 1
        return x


class ResBlock(nn.Module):
    def __init__(self, in_channel, out_channel, blur_kernel=[1, 3, 3, 1], downsample=True):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, 3)
        self.conv2 = ConvLayer(in_channel, out_channel, 3, downsample=downsample)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=downsample, activate=False, bias=False)

    def forward(self, input, feats: Optional[list] = None):
        h1 = self.conv1(input)
        h2 = self.conv2(h1)
        if feats is not None:
            feats += [h1, h2]
        return _glue.add_scale(h2, self.skip(input), 1 / math.sqrt(2))


class Discriminator(nn.Module):
    def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        channels = _CHANNELS(channel_multiplier)
        convs = [ConvLayer(3, channels[size], 1)]
        log_size = int(math.log(size, 2))
        in_channel = channels[size]
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(in_channel, out_channel, blur_kernel))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.stddev_group = 25
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, channels[4], 3)
        self.final_linear = nn.Sequential(EqualLinear(channels[4] * 4 * 4, channels[4], activation="fused_lrelu"),
                                          EqualLinear(channels[4], 1))
        if _CHANNELS_LAST:
            # keep the 4-D conv weights in channels-last memory (same shapes / state_dict): the library convolutions then
            # take weight * scale as is instead of re-laying it out on every call (~190 copy kernels per iteration)
            self.to(memory_format=torch.channels_last)

    def estimate_fisher(self, loglikelihood):
        return _estimate_fisher(self, loglikelihood)

    def forward(self, *args, **kwargs):
        with _Prescaled(self):
            return self._forward(*args, **kwargs)

    def _from_rgb(self, inp):
        """``self.convs[0]``: ConvLayer(3, C, 1) = 1x1 EqualConv2d + FusedLeakyReLU, as one fused pass when possible."""
        layer = self.convs[0]
        conv, act = layer[0], layer[1] if len(layer) > 1 else None
        if (_CHANNELS_LAST and isinstance(conv, EqualConv2d) and isinstance(act, FusedLeakyReLU) and conv.bias is None
                and conv.stride == 1 and conv.padding == 0 and _glue.from_rgb_ok(inp, conv.weight)):
            return _glue.from_rgb(inp, conv.weight, act.bias, conv.scale, act.negative_slope, act.scale)
        return layer(_fmt(inp))

    def _forward(self, inp, ind=None, real=False, stddev_group=None):
        """``stddev_group``: override of ``min(batch, self.stddev_group)`` for the minibatch-stddev grouping -- used by
        rick_b200.adapt.d_pair to score two batches in one pass with each batch's own statistics."""
        feat: list = []
        out = self._from_rgb(inp)
        feat.append(out)
        for block in list(self.convs)[1:]:
            out = block(out, feat)          # conv1 / conv2 evaluated ONCE (see module docstring)
        batch, channel, height, width = out.shape
        group = min(batch, self.stddev_group) if stddev_group is None else stddev_group
        stddev = out.view(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        stddev = torch.sqrt(stddev.var(0, unbiased=False) + 1e-8)
        stddev = stddev.mean([2, 3, 4], keepdims=True).squeeze(2)
        stddev = stddev.repeat(group, 1, height, width)
        out = torch.cat([out, stddev], 1)
        out = self.final_conv(out)
        feat.append(out)
        out = self.final_linear(out.reshape(batch, -1))      # logical (C, H, W) order whatever the memory format
        return out, feat

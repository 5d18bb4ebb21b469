"""Inference executor of the StyleGAN2 generator on the tcgen05 path (FID-style sample generation,
gan_training/eval.py:31-46, and every ``no_grad`` generator call of the adaptation loop).

Same parameters, same state_dict, same call signature and return value as ``Generator.forward``
(model_probe_tune.py:509-592); what changes is the execution plan.  Activations stay channels-last fp32 in HBM and
each StyledConv is one or two kernels:

    3x3 StyledConv        rick_conv_tc        implicit GEMM (TF32 tcgen05, fp32 TMEM accumulate) with demodulation, noise,
                                              bias and leaky-ReLU in the epilogue; the epilogue also emits the next
                                              layer's pre-modulated input (y * s_next) so no separate modulation pass exists
    upsampling StyledConv rick_conv_tc        stride-2 transposed conv as 4 polyphase sub-convolutions -> (2H+1)x(2W+1)
                          rick_blur_nhwc      4x4 blur + demod + noise + bias + leaky-ReLU (+ * s_next) in one pass
    ToRGB                 rick_to_rgb_nhwc    3-channel modulated 1x1 conv + bias + upsampled skip, one read of y
    skip upsample         rick_upfirdn2d      on the 3-channel NCHW image

Per-layer styles / demodulation coefficients are (B, C) matrices computed up front (tiny cuBLAS calls).
The convolution weights are read IN PLACE: ModulatedConv2d stores them [Cout][k][k][Cin], which the kernel consumes
through a strided tensor map, so the executor holds views of the parameters and is always current -- no re-packing
after an optimiser step (``refresh()`` is kept as a no-op for callers written against the packing version).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.nn import functional as F

from . import conv_tc as ct
from .op import upfirdn2d
from .op import glue as _glue
from .op.glue import weight_sqsum


class _ConvPlan:
    def __init__(self, sconv):
        mc = sconv.conv
        self.w = mc.weight.detach().squeeze(0)                   # (Cout, Cin, 3, 3) VIEW of the parameter's memory
        self.wscale = mc.scale                                   # equalised-lr scale: rides on the style (see below)
        self.cin, self.cout = mc.in_channel, mc.out_channel
        self.upsample = mc.upsample
        self.blur_taps = mc.blur.kernel.detach().contiguous() if mc.upsample else None
        self.blur_pad = mc.blur.pad if mc.upsample else None
        self.mod = mc.modulation
        self.noise_w = sconv.noise.weight.detach()
        self.bias = sconv.activate.bias.detach()
        self.alpha, self.scale = sconv.activate.negative_slope, sconv.activate.scale


class _RgbPlan:
    def __init__(self, rgb):
        mc = rgb.conv
        self.w = mc.weight.detach()[0, :, :, 0, 0]                                  # (3, Cin) view
        self.wscale = mc.scale
        self.mod = mc.modulation
        self.bias = rgb.bias.detach().reshape(3).contiguous()
        self.up = rgb.upsample if hasattr(rgb, "upsample") else None


class FusedGenerator:
    """``FusedGenerator(G)(styles, ...)`` == ``G(styles, ...)`` under ``no_grad`` (TF32 tolerance)."""

    def __init__(self, generator):
        self.g = generator
        self.refresh()

    @staticmethod
    def supports(generator) -> bool:
        """True when every StyledConv of ``generator`` fits the tcgen05 kernel (Cin % 32 == 0, Cout % 128 == 0): all
        layers up to 256 px at channel_multiplier 2.  The 512 / 1024 px layers (64 / 32 channels) do not; callers then
        keep the module path."""
        convs = [generator.conv1] + list(generator.convs)
        return all(ct.supported(c.conv.in_channel, c.conv.out_channel) for c in convs)

    def refresh(self):
        """(Re)build the per-layer plans.  The plans hold VIEWS of the parameters, so this is only needed when a
        parameter tensor was replaced (not when its values changed)."""
        g = self.g
        self.convs: List[_ConvPlan] = [_ConvPlan(g.conv1)] + [_ConvPlan(c) for c in g.convs]
        self.rgbs: List[_RgbPlan] = [_RgbPlan(g.to_rgb1)] + [_RgbPlan(r) for r in g.to_rgbs]
        for p in self.convs:
            if not ct.supported(p.cin, p.cout):
                raise RuntimeError(f"FusedGenerator: layer {p.cin}->{p.cout} is outside the tcgen05 kernel's shapes")

    @staticmethod
    def _weight(p):
        """the layer's weight as the kernel wants it: in place when stored channels-last, else packed per call"""
        w = p.w
        return w if w.is_contiguous(memory_format=torch.channels_last) else ct.pack_weight(w)

    @torch.no_grad()
    def __call__(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                 input_is_latent=False, noise=None, randomize_noise=True):
        g = self.g
        latent = self._latent(styles, inject_index, truncation, truncation_latent, input_is_latent)
        b = latent.shape[0]
        dev = latent.device
        if noise is None:
            noise = ([None] * g.num_layers if randomize_noise
                     else [getattr(g.noises, f"noise_{i}") for i in range(g.num_layers)])

        # latent index per StyledConv / ToRGB (model_probe_tune.py:567-582)
        conv_idx = [0] + [i for blk in range(len(g.to_rgbs)) for i in (1 + 2 * blk, 2 + 2 * blk)]
        rgb_idx = [1] + [3 + 2 * blk for blk in range(len(g.to_rgbs))]
        # every layer's modulation from one launch (order of Generator._modulated(): conv1, rgb1, then up / conv / rgb)
        latent = latent.contiguous()
        mods = g.all_modulations(latent)
        if mods is not None:
            raw_conv = [mods[0]] + [mods[2 + 3 * blk + j] for blk in range(len(g.to_rgbs)) for j in (0, 1)]
            raw_rgb = [mods[1]] + [mods[4 + 3 * blk] for blk in range(len(g.to_rgbs))]
        else:
            raw_conv = [p.mod(latent[:, i]) for p, i in zip(self.convs, conv_idx)]
            raw_rgb = [p.mod(latent[:, i]) for p, i in zip(self.rgbs, rgb_idx)]
        # conv(x * s, scale * W) == conv(x * (scale * s), W): the equalised-lr scale rides on the (B, Cin) style
        # (scaled style, demodulation) per layer: one launch each (rick_demod_fwd)
        sd = [_glue.demod(r.contiguous(), weight_sqsum(p.w), p.wscale ** 2, 1e-8, p.wscale) for p, r in zip(self.convs, raw_conv)]
        s = [x[1].contiguous() for x in sd]                                                           # (B, Cin)
        demod = [x[0].contiguous() for x in sd]                                                       # (B, Cout)
        s_rgb = [r * p.wscale for p, r in zip(self.rgbs, raw_rgb)]

        def layer_noise(li, h, w):
            n = noise[li]
            if n is None:
                return torch.randn(b, h, w, device=dev)
            return n.expand(b, 1, h, w).reshape(b, h, w).contiguous()

        def rgb(plan, s_r, y, skip):
            wmod = (plan.w[None] * s_r[:, None, :]).contiguous()                                   # (B, 3, Cin)
            if skip is not None:
                skip = upfirdn2d(skip, plan.up.kernel, up=plan.up.factor, down=1, pad=plan.up.pad)
            return ct.to_rgb_nhwc(y, wmod, plan.bias, skip)

        # 4x4 constant input, pre-modulated for conv1
        x0 = g.input.input.detach().permute(0, 2, 3, 1)                                            # (1, 4, 4, C)
        xm = (x0 * s[0][:, None, None, :]).contiguous()
        p0 = self.convs[0]
        has_next = len(self.convs) > 1
        res = ct.conv_tc_nhwc(xm, self._weight(p0), ct.geom_conv(b, 4, 4, p0.cin, p0.cout, 3, 1, 1), demod=demod[0],
                              noise=layer_noise(0, 4, 4), noise_weight=p0.noise_w, bias=p0.bias, act=True,
                              alpha=p0.alpha, scale=p0.scale, s_next=s[1] if has_next else None, want_out2=has_next)
        y, ym = res if has_next else (res, None)
        image = rgb(self.rgbs[0], s_rgb[0], y, None)
        h = 4
        for blk in range(len(g.to_rgbs)):
            up, cv = self.convs[1 + 2 * blk], self.convs[2 + 2 * blk]
            li = 1 + 2 * blk
            # upsampling StyledConv: transposed conv (raw accumulators) -> blur with the fused epilogue; only the copy
            # pre-modulated for the following conv is written
            raw = ct.conv_tc_nhwc(ym, self._weight(up), ct.geom_conv_transpose_s2(b, h, h, up.cin, up.cout))
            h *= 2
            ym = ct.blur_nhwc(raw, up.blur_taps, up.blur_pad, demod=demod[li], noise=layer_noise(li, h, h),
                              noise_weight=up.noise_w, bias=up.bias, act=True, alpha=up.alpha, scale=up.scale,
                              s_next=s[li + 1])
            last = blk == len(g.to_rgbs) - 1
            res = ct.conv_tc_nhwc(ym, self._weight(cv), ct.geom_conv(b, h, h, cv.cin, cv.cout, 3, 1, 1), demod=demod[li + 1],
                                  noise=layer_noise(li + 1, h, h), noise_weight=cv.noise_w, bias=cv.bias, act=True,
                                  alpha=cv.alpha, scale=cv.scale, s_next=None if last else s[li + 2],
                                  want_out2=not last)
            y, ym = (res, None) if last else res
            image = rgb(self.rgbs[1 + blk], s_rgb[1 + blk], y, image)
        return (image, latent) if return_latents else (image, None)

    def _latent(self, styles, inject_index, truncation, truncation_latent, input_is_latent):
        import random
        g = self.g
        if not input_is_latent:
            styles = [g.map_latent(s) for s in styles]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            return styles[0].unsqueeze(1).repeat(1, g.n_latent, 1) if styles[0].ndim < 3 else styles[0]
        if inject_index is None:
            inject_index = random.randint(1, g.n_latent - 1)
        return torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                          styles[1].unsqueeze(1).repeat(1, g.n_latent - inject_index, 1)], 1)

"""TEST INFRASTRUCTURE ONLY -- functional CPU restatement of the StyleGAN2 G / D stack.

The reference keeps this path in ``nn.Module`` classes
(gan_training/models/model_probe_tune.py).  The oracle restates the same
arithmetic, in the same operation order, as pure functions over a
``{state_dict key: tensor}`` mapping so that it can run on any box without the
reference tree, and so that torch autograd on CPU supplies first and second
derivatives (Fisher gradients, R1, path-length).  Citations give the reference
lines each function follows.  Validated against the real modules by
``oracle/make_golden.py`` (max-abs 0 on CPU for G and D forward).
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence

import torch
from torch.nn import functional as F

from . import ops_oracle as ops

Params = Dict[str, torch.Tensor]

CHANNELS = lambda cm: {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm,  # noqa: E731
                       256: 64 * cm, 512: 32 * cm, 1024: 16 * cm}                       # model_probe_tune.py:400-410


def make_kernel(k: Sequence[float]) -> torch.Tensor:
    """model_probe_tune.py:29-37."""
    t = torch.tensor(k, dtype=torch.float32)
    if t.ndim == 1:
        t = t[None, :] * t[:, None]
    return t / t.sum()


# ----------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------

def equal_linear(p: Params, key: str, x: torch.Tensor, lr_mul: float = 1.0, activation: bool = False) -> torch.Tensor:
    """EqualLinear.forward, model_probe_tune.py:158-168."""
    w = p[key + ".weight"]
    b = p[key + ".bias"]
    scale = (1 / math.sqrt(w.shape[1])) * lr_mul
    if activation:
        return ops.fused_leaky_relu(F.linear(x, w * scale), b * lr_mul)
    return F.linear(x, w * scale, bias=b * lr_mul)


def equal_conv2d(p: Params, key: str, x: torch.Tensor, stride: int, padding: int) -> torch.Tensor:
    """EqualConv2d.forward, model_probe_tune.py:121-130."""
    w = p[key + ".weight"]
    scale = 1 / math.sqrt(w.shape[1] * w.shape[2] ** 2)
    return F.conv2d(x, w * scale, bias=p.get(key + ".bias"), stride=stride, padding=padding)


def modulated_conv2d(p: Params, key: str, x: torch.Tensor, style: torch.Tensor, demodulate: bool = True,
                     upsample: bool = False) -> torch.Tensor:
    """ModulatedConv2d.forward, model_probe_tune.py:243-284 (downsample branch unused by G, omitted)."""
    w = p[key + ".weight"]                                  # (1, Cout, Cin, k, k)
    _, cout, cin, k, _ = w.shape
    b, _, h, wd = x.shape
    s = equal_linear(p, key + ".modulation", style).view(b, 1, cin, 1, 1)
    weight = (1 / math.sqrt(cin * k * k)) * w * s            # scale * W * style   (247)
    if demodulate:
        demod = torch.rsqrt(weight.pow(2).sum([2, 3, 4]) + 1e-8)
        weight = weight * demod.view(b, cout, 1, 1, 1)
    if upsample:
        xin = x.view(1, b * cin, h, wd)
        wt = weight.transpose(1, 2).reshape(b * cin, cout, k, k)
        out = F.conv_transpose2d(xin, wt, padding=0, stride=2, groups=b)
        out = out.view(b, cout, out.shape[2], out.shape[3])
        # Blur(pad=(pad0,pad1), upsample_factor=2)                               (209-215, 268)
        taps = p[key + ".blur.kernel"]
        pp = (taps.shape[0] - 2) - (k - 1)
        return ops.upfirdn2d(out, taps, pad=((pp + 1) // 2 + 1, pp // 2 + 1))
    xin = x.view(1, b * cin, h, wd)
    out = F.conv2d(xin, weight.view(b * cout, cin, k, k), padding=k // 2, groups=b)
    return out.view(b, cout, out.shape[2], out.shape[3])


def styled_conv(p: Params, key: str, x: torch.Tensor, style: torch.Tensor, noise: Optional[torch.Tensor],
                upsample: bool = False) -> torch.Tensor:
    """StyledConv.forward, model_probe_tune.py:342-348 (+ NoiseInjection 293-298)."""
    out = modulated_conv2d(p, key + ".conv", x, style, demodulate=True, upsample=upsample)
    if noise is None:
        noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
    out = out + p[key + ".noise.weight"] * noise
    return ops.fused_leaky_relu(out, p[key + ".activate.bias"])


def to_rgb(p: Params, key: str, x: torch.Tensor, style: torch.Tensor, skip: Optional[torch.Tensor]) -> torch.Tensor:
    """ToRGB.forward, model_probe_tune.py:361-370 (+ Upsample 40-58)."""
    out = modulated_conv2d(p, key + ".conv", x, style, demodulate=False) + p[key + ".bias"]
    if skip is not None:
        taps = p[key + ".upsample.kernel"]
        pp = taps.shape[0] - 2
        out = out + ops.upfirdn2d(skip, taps, up=2, down=1, pad=((pp + 1) // 2 + 1, pp // 2))
    return out


# ----------------------------------------------------------------------------------------------
# generator
# ----------------------------------------------------------------------------------------------

def g_n_latent(size: int) -> int:
    return int(math.log(size, 2)) * 2 - 2


def g_num_layers(size: int) -> int:
    return (int(math.log(size, 2)) - 2) * 2 + 1


def g_style(p: Params, z: torch.Tensor, n_mlp: int = 8, lr_mlp: float = 0.01) -> torch.Tensor:
    """PixelNorm + n_mlp EqualLinear(fused_lrelu), model_probe_tune.py:25-26, 389-398."""
    x = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(n_mlp):
        x = equal_linear(p, f"style.{i + 1}", x, lr_mul=lr_mlp, activation=True)
    return x


def g_forward(p: Params, styles: List[torch.Tensor], size: int = 256, noise=None, randomize_noise: bool = True,
              inject_index: Optional[int] = None, input_is_latent: bool = False, return_latents: bool = False,
              n_mlp: int = 8):
    """Generator.forward, model_probe_tune.py:509-592 (truncation path omitted: the trainer never uses it)."""
    n_latent, num_layers = g_n_latent(size), g_num_layers(size)
    if not input_is_latent:
        styles = [g_style(p, s, n_mlp) for s in styles]
    if noise is None:
        noise = [None] * num_layers if randomize_noise else [p[f"noises.noise_{i}"] for i in range(num_layers)]
    if len(styles) < 2:
        latent = styles[0].unsqueeze(1).repeat(1, n_latent, 1) if styles[0].ndim < 3 else styles[0]
    else:
        if inject_index is None:
            inject_index = random.randint(1, n_latent - 1)     # python RNG, as the reference (556)
        latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                            styles[1].unsqueeze(1).repeat(1, n_latent - inject_index, 1)], 1)

    out = p["input.input"].repeat(latent.shape[0], 1, 1, 1)
    out = styled_conv(p, "conv1", out, latent[:, 0], noise[0])
    skip = to_rgb(p, "to_rgb1", out, latent[:, 1], None)
    i = 1
    for blk in range((num_layers - 1) // 2):
        out = styled_conv(p, f"convs.{2 * blk}", out, latent[:, i], noise[1 + 2 * blk], upsample=True)
        out = styled_conv(p, f"convs.{2 * blk + 1}", out, latent[:, i + 1], noise[2 + 2 * blk])
        skip = to_rgb(p, f"to_rgbs.{blk}", out, latent[:, i + 2], skip)
        i += 2
    return (skip, latent) if return_latents else (skip, None)


# ----------------------------------------------------------------------------------------------
# discriminator
# ----------------------------------------------------------------------------------------------

def _blur(p: Params, key: str, x: torch.Tensor, ksize: int) -> torch.Tensor:
    """Blur in front of a stride-2 ConvLayer, model_probe_tune.py:608-614."""
    taps = p[key + ".kernel"]
    pp = (taps.shape[0] - 2) + (ksize - 1)
    return ops.upfirdn2d(x, taps, pad=((pp + 1) // 2, pp // 2))


def d_resblock(p: Params, key: str, x: torch.Tensor) -> torch.Tensor:
    """ResBlock.forward, model_probe_tune.py:654-660."""
    h = equal_conv2d(p, key + ".conv1.0", x, 1, 1)
    h = ops.fused_leaky_relu(h, p[key + ".conv1.1.bias"])
    h = _blur(p, key + ".conv2.0", h, 3)
    h = equal_conv2d(p, key + ".conv2.1", h, 2, 0)
    h = ops.fused_leaky_relu(h, p[key + ".conv2.2.bias"])
    s = _blur(p, key + ".skip.0", x, 1)
    s = equal_conv2d(p, key + ".skip.1", s, 2, 0)
    return (h + s) / math.sqrt(2)


def d_forward(p: Params, img: torch.Tensor, size: int = 256) -> torch.Tensor:
    """Discriminator.forward, model_probe_tune.py:732-764.

    The reference evaluates conv1/conv2 of every block a second time to fill a
    ``feat`` list that the trainer discards (740-744, train:407-410); those
    extra evaluations do not influence the logits and are not restated."""
    n_blocks = int(math.log(size, 2)) - 2
    out = equal_conv2d(p, "convs.0.0", img, 1, 0)
    out = ops.fused_leaky_relu(out, p["convs.0.1.bias"])
    for b in range(1, n_blocks + 1):
        out = d_resblock(p, f"convs.{b}", out)
    batch, channel, height, width = out.shape
    group = min(batch, 25)
    sd = out.view(group, -1, 1, channel, height, width)
    sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8)
    sd = sd.mean([2, 3, 4], keepdims=True).squeeze(2)
    sd = sd.repeat(group, 1, height, width)
    out = torch.cat([out, sd], 1)
    out = equal_conv2d(p, "final_conv.0", out, 1, 1)
    out = ops.fused_leaky_relu(out, p["final_conv.1.bias"])
    out = out.view(batch, -1)
    out = equal_linear(p, "final_linear.0", out, activation=True)
    return equal_linear(p, "final_linear.1", out)


# ----------------------------------------------------------------------------------------------
# losses (train_dynamic_update_prune.py:82-118)
# ----------------------------------------------------------------------------------------------

def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    (grad_real,) = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_path_regularize(fake_img, latents, mean_path_length, noise=None, decay=0.01):
    if noise is None:
        noise = torch.randn_like(fake_img)
    noise = noise / math.sqrt(fake_img.shape[2] * fake_img.shape[3])
    (grad,) = torch.autograd.grad(outputs=(fake_img * noise).sum(), inputs=latents, create_graph=True)
    path_lengths = torch.sqrt(grad.pow(2).sum(2).mean(1))
    path_mean = mean_path_length + decay * (path_lengths.mean() - mean_path_length)
    path_penalty = (path_lengths - path_mean).pow(2).mean()
    return path_penalty, path_mean.detach(), path_lengths


# ----------------------------------------------------------------------------------------------
# parameter bookkeeping
# ----------------------------------------------------------------------------------------------

def g_param_names(size: int = 256, n_mlp: int = 8) -> List[str]:
    """Parameter names in ``Generator.named_parameters()`` order (registration order in __init__)."""
    names = []
    for i in range(n_mlp):
        names += [f"style.{i + 1}.weight", f"style.{i + 1}.bias"]
    names.append("input.input")

    def sconv(k):
        return [f"{k}.conv.weight", f"{k}.conv.modulation.weight", f"{k}.conv.modulation.bias",
                f"{k}.noise.weight", f"{k}.activate.bias"]

    def rgb(k):
        return [f"{k}.bias", f"{k}.conv.weight", f"{k}.conv.modulation.weight", f"{k}.conv.modulation.bias"]

    names += sconv("conv1") + rgb("to_rgb1")
    nblk = int(math.log(size, 2)) - 2
    for i in range(2 * nblk):
        names += sconv(f"convs.{i}")
    for i in range(nblk):
        names += rgb(f"to_rgbs.{i}")
    return names


def d_param_names(size: int = 256) -> List[str]:
    """Parameter names in ``Discriminator.named_parameters()`` order."""
    names = ["convs.0.0.weight", "convs.0.1.bias"]
    for b in range(1, int(math.log(size, 2)) - 1):
        names += [f"convs.{b}.conv1.0.weight", f"convs.{b}.conv1.1.bias", f"convs.{b}.conv2.1.weight",
                  f"convs.{b}.conv2.2.bias", f"convs.{b}.skip.1.weight"]
    names += ["final_conv.0.weight", "final_conv.1.bias", "final_linear.0.weight", "final_linear.0.bias",
              "final_linear.1.weight", "final_linear.1.bias"]
    return names


def init_g_params(size: int = 256, style_dim: int = 512, n_mlp: int = 8, channel_multiplier: int = 2,
                  lr_mlp: float = 0.01, generator: Optional[torch.Generator] = None) -> Params:
    """Random-init G parameters + buffers with the reference's shapes and init distributions
    (model_probe_tune.py:145-148, 229-233, 291, 305, 359, 431).  The draw ORDER differs from the
    module constructors; parity tests therefore copy one state_dict into both sides."""
    ch = CHANNELS(channel_multiplier)
    rn = lambda *s: torch.randn(*s, generator=generator)  # noqa: E731
    p: Params = {}
    for i in range(n_mlp):
        p[f"style.{i + 1}.weight"] = rn(style_dim, style_dim) / lr_mlp
        p[f"style.{i + 1}.bias"] = torch.zeros(style_dim)
    p["input.input"] = rn(1, ch[4], 4, 4)
    blur = make_kernel([1, 3, 3, 1])

    def sconv(k, cin, cout, up):
        p[f"{k}.conv.weight"] = rn(1, cout, cin, 3, 3)
        if up:
            p[f"{k}.conv.blur.kernel"] = blur * 4
        p[f"{k}.conv.modulation.weight"] = rn(cin, style_dim)
        p[f"{k}.conv.modulation.bias"] = torch.ones(cin)
        p[f"{k}.noise.weight"] = torch.zeros(1)
        p[f"{k}.activate.bias"] = torch.zeros(cout)

    def rgb(k, cin, up):
        p[f"{k}.bias"] = torch.zeros(1, 3, 1, 1)
        if up:
            p[f"{k}.upsample.kernel"] = blur * 4
        p[f"{k}.conv.weight"] = rn(1, 3, cin, 1, 1)
        p[f"{k}.conv.modulation.weight"] = rn(cin, style_dim)
        p[f"{k}.conv.modulation.bias"] = torch.ones(cin)

    sconv("conv1", ch[4], ch[4], False)
    rgb("to_rgb1", ch[4], False)
    log_size = int(math.log(size, 2))
    cin = ch[4]
    for li in range(g_num_layers(size)):
        res = (li + 5) // 2
        p[f"noises.noise_{li}"] = rn(1, 1, 2 ** res, 2 ** res)
    for i in range(3, log_size + 1):
        cout = ch[2 ** i]
        sconv(f"convs.{2 * (i - 3)}", cin, cout, True)
        sconv(f"convs.{2 * (i - 3) + 1}", cout, cout, False)
        rgb(f"to_rgbs.{i - 3}", cout, True)
        cin = cout
    return p


def init_d_params(size: int = 256, channel_multiplier: int = 2,
                  generator: Optional[torch.Generator] = None) -> Params:
    """Random-init D parameters + buffers (model_probe_tune.py:107-116, 664-701)."""
    ch = CHANNELS(channel_multiplier)
    rn = lambda *s: torch.randn(*s, generator=generator)  # noqa: E731
    blur = make_kernel([1, 3, 3, 1])
    p: Params = {"convs.0.0.weight": rn(ch[size], 3, 1, 1), "convs.0.1.bias": torch.zeros(ch[size])}
    cin = ch[size]
    log_size = int(math.log(size, 2))
    for j, i in enumerate(range(log_size, 2, -1)):
        cout = ch[2 ** (i - 1)]
        k = f"convs.{j + 1}"
        p[f"{k}.conv1.0.weight"] = rn(cin, cin, 3, 3)
        p[f"{k}.conv1.1.bias"] = torch.zeros(cin)
        p[f"{k}.conv2.0.kernel"] = blur.clone()
        p[f"{k}.conv2.1.weight"] = rn(cout, cin, 3, 3)
        p[f"{k}.conv2.2.bias"] = torch.zeros(cout)
        p[f"{k}.skip.0.kernel"] = blur.clone()
        p[f"{k}.skip.1.weight"] = rn(cout, cin, 1, 1)
        cin = cout
    p["final_conv.0.weight"] = rn(ch[4], cin + 1, 3, 3)
    p["final_conv.1.bias"] = torch.zeros(ch[4])
    p["final_linear.0.weight"] = rn(ch[4], ch[4] * 16)
    p["final_linear.0.bias"] = torch.zeros(ch[4])
    p["final_linear.1.weight"] = rn(1, ch[4])
    p["final_linear.1.bias"] = torch.zeros(1)
    return p


def estimate_fisher(loss: torch.Tensor, p: Params, names: List[str]) -> Dict[str, torch.Tensor]:
    """Generator/Discriminator.estimate_fisher, model_probe_tune.py:481-504 / 706-729:
    ``autograd.grad(loss, parameters, retain_graph=True)`` then element-wise square."""
    grads = torch.autograd.grad(loss, [p[n] for n in names], retain_graph=True)
    return {n: g.detach() ** 2 for n, g in zip(names, grads)}

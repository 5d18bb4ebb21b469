"""Experiment (GPU box): how much of the tcgen05 generator's image error is TF32 *truncation* of the MMA operands?

tcgen05.mma.kind::tf32 reads the upper 19 bits of each fp32 operand (truncation: a systematic shrink of every product),
whereas rounding the operands to nearest (cvt.rna.tf32.f32) is unbiased.  Variants of FusedGenerator on the 256 px golden
image: operands as stored / weights rounded to nearest / activations rounded to nearest / both / outputs rescaled by the
expected truncation bias.  Prints the max-abs error of each against the committed reference image.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import synth
from rick_b200 import conv_tc as ct
from rick_b200 import fused
from rick_b200 import stylegan2 as sg


def rna(x: torch.Tensor) -> torch.Tensor:
    """round-to-nearest (ties away) onto the TF32 grid: 10 explicit mantissa bits"""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "g256_golden.npz"))
lat = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fisher_latents.npy"))).cuda()
G = sg.Generator(256, 512, 8)
G.load_state_dict(synth.g_state(256, 1))
G = G.cuda()
want = torch.as_tensor(gold["img_sub4"]).double()


def err(img):
    return (img[:, :, ::4, ::4].cpu().double() - want).abs().max().item()


with torch.no_grad():
    print("module path (library convs):", err(G([lat[:2]], randomize_noise=False)[0]))
    fg = fused.FusedGenerator(G)
    print("tcgen05 executor, operands as stored (truncation):", err(fg([lat[:2]], randomize_noise=False)[0]))

    for p in fg.convs:
        p.wt = rna(p.wt)
    print("  + weights rounded to nearest:", err(fg([lat[:2]], randomize_noise=False)[0]))

    orig = ct.conv_tc_nhwc

    def conv_rounded_x(xm, wt, geom, **kw):
        return orig(rna(xm), wt, geom, **kw)
    ct.conv_tc_nhwc = conv_rounded_x
    print("  + weights and activations rounded to nearest:", err(fg([lat[:2]], randomize_noise=False)[0]))

    fg.refresh()
    print("  activations rounded only:", err(fg([lat[:2]], randomize_noise=False)[0]))
    ct.conv_tc_nhwc = orig

    # bias compensation: E[relative truncation error] per operand = 0.5 * 2^-10 * E[1/m], m log-uniform in [1, 2)
    bias = 0.5 * 2.0 ** -10 * (0.5 / np.log(2.0))

    def conv_comp(xm, wt, geom, **kw):
        out = orig(xm, wt, geom, **kw)
        return out
    for factor, name in ((1 + 2 * bias, "both operands"), (1 + bias, "one operand")):
        fg.refresh()
        for p in fg.convs:
            p.wt = p.wt * factor
        print(f"  truncation, weights pre-scaled by {factor:.7f} ({name} compensation):",
              err(fg([lat[:2]], randomize_noise=False)[0]))

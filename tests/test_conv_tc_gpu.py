"""tcgen05 implicit-GEMM convolution (through the C ABI) against fp32 library convolutions.
The kernel consumes fp32 operands as TF32 (10-bit mantissa, fp32 accumulate): tolerance 2e-3 of the output scale."""
import math

import pytest
import torch
from torch.nn import functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-3


@pytest.fixture(scope="module", autouse=True)
def _no_tf32_in_reference():
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = prev


def _err(got, want):
    return ((got.double() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-30)).item()


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).cuda()


CONV_CASES = [
    # B, H, W, Cin, Cout, k, stride, pad
    (1, 16, 16, 32, 128, 1, 1, 0),        # plain GEMM: one tap, one k-block -> isolates descriptors / TMEM / epilogue
    (1, 16, 16, 128, 128, 1, 1, 0),       # K loop over 4 stages
    (2, 16, 16, 512, 512, 1, 1, 0),       # pipeline wrap-around, 4 cout tiles
    (2, 16, 16, 64, 128, 3, 1, 1),        # 3x3 taps with zero padding from TMA out-of-bounds fill
    (2, 4, 4, 512, 512, 3, 1, 1),         # G conv1: tile spans both samples
    (2, 8, 8, 512, 512, 3, 1, 1),
    (2, 32, 32, 512, 512, 3, 1, 1),
    (2, 64, 64, 256, 256, 3, 1, 1),       # multi-tile per image
    (1, 20, 12, 64, 128, 3, 1, 1),        # ragged: partial tiles in both directions
    (3, 5, 7, 32, 128, 3, 1, 1),
    (2, 128, 128, 128, 128, 3, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=["_".join(map(str, c)) for c in CONV_CASES])
def test_conv_tc_matches_fp32_conv(case):
    from rick_b200 import conv_tc as ct
    b, h, w, cin, cout, k, stride, pad = case
    x = _rand(b, cin, h, w, seed=1)
    wgt = _rand(cout, cin, k, k, seed=2) / math.sqrt(cin * k * k)
    want = F.conv2d(x, wgt, stride=stride, padding=pad)
    geom = ct.geom_conv(b, h, w, cin, cout, k, stride, pad)
    got = ct.conv_tc_nhwc(x.permute(0, 2, 3, 1).contiguous(), ct.pack_weight(wgt), geom)
    got = got.permute(0, 3, 1, 2)
    assert got.shape == want.shape
    assert _err(got, want) < TOL


@pytest.mark.parametrize("case", [(2, 4, 4, 512, 512), (2, 16, 16, 512, 512), (2, 64, 64, 512, 256), (1, 9, 5, 64, 128),
                                  (2, 128, 128, 256, 128)], ids=lambda c: "_".join(map(str, c)))
def test_conv_transpose_polyphase(case):
    from rick_b200 import conv_tc as ct
    b, h, w, cin, cout = case
    x = _rand(b, cin, h, w, seed=3)
    wgt = _rand(cout, cin, 3, 3, seed=4) / math.sqrt(cin * 9)
    want = F.conv_transpose2d(x, wgt.transpose(0, 1), stride=2, padding=0)
    geom = ct.geom_conv_transpose_s2(b, h, w, cin, cout)
    got = ct.conv_tc_nhwc(x.permute(0, 2, 3, 1).contiguous(), ct.pack_weight(wgt), geom).permute(0, 3, 1, 2)
    assert got.shape == want.shape == (b, cout, 2 * h + 1, 2 * w + 1)
    assert _err(got, want) < TOL


def test_fused_epilogue_styled_conv():
    """demod * acc + noise_w * noise + bias -> leaky-ReLU * sqrt(2), and the pre-modulated copy for the next layer."""
    from rick_b200 import conv_tc as ct
    b, h, w, cin, cout = 2, 32, 32, 256, 256
    x, wgt = _rand(b, cin, h, w, seed=5), _rand(cout, cin, 3, 3, seed=6) / math.sqrt(cin * 9)
    s, demod = _rand(b, cin, seed=7), _rand(b, cout, seed=8).abs() + 0.5
    noise, nw, bias, s_next = _rand(b, 1, h, w, seed=9), _rand(1, seed=10), _rand(cout, seed=11), _rand(b, cout, seed=12)
    xm = x * s[:, :, None, None]
    y = F.conv2d(xm, wgt, padding=1) * demod[:, :, None, None] + nw * noise + bias[None, :, None, None]
    want = F.leaky_relu(y, 0.2) * math.sqrt(2)
    want2 = want * s_next[:, :, None, None]
    geom = ct.geom_conv(b, h, w, cin, cout, 3, 1, 1)
    got, got2 = ct.conv_tc_nhwc(xm.permute(0, 2, 3, 1).contiguous(), ct.pack_weight(wgt), geom, demod=demod,
                                noise=noise.reshape(b, h, w).contiguous(), noise_weight=nw, bias=bias, act=True,
                                s_next=s_next, want_out2=True)
    assert _err(got.permute(0, 3, 1, 2), want) < TOL
    assert _err(got2.permute(0, 3, 1, 2), want2) < TOL


def test_strided_conv_d_path():
    """3x3 stride-2 pad-0 (D conv2 after its blur) and 1x1 stride-2 (D skip) through TMA traversal strides."""
    from rick_b200 import conv_tc as ct
    for (b, h, w, cin, cout, k) in [(2, 33, 33, 128, 128, 3), (2, 17, 17, 512, 512, 3), (2, 32, 32, 128, 256, 1)]:
        x, wgt = _rand(b, cin, h, w, seed=13), _rand(cout, cin, k, k, seed=14) / math.sqrt(cin * k * k)
        want = F.conv2d(x, wgt, stride=2)
        geom = ct.geom_conv(b, h, w, cin, cout, k, 2, 0)
        got = ct.conv_tc_nhwc(x.permute(0, 2, 3, 1).contiguous(), ct.pack_weight(wgt), geom).permute(0, 3, 1, 2)
        assert got.shape == want.shape
        assert _err(got, want) < TOL, (b, h, w, cin, cout, k)


def test_unsupported_shapes_are_reported():
    from rick_b200 import conv_tc as ct
    x = _rand(1, 8, 8, 24)
    with pytest.raises(RuntimeError, match="unsupported"):
        ct.conv_tc_nhwc(x, _rand(1, 128, 24), ct.geom_conv(1, 8, 8, 24, 128, 1))


def test_blur_nhwc_fused_epilogue():
    from rick_b200 import conv_tc as ct
    from oracle import ops_oracle as ops
    b, c, h = 2, 64, 33
    x = _rand(b, c, h, h, seed=20)
    t = torch.tensor([1., 3., 3., 1.])
    taps = (torch.outer(t, t) / 16).cuda()
    demod, noise, nw, bias, sn = _rand(b, c, seed=21), _rand(b, 1, h - 1, h - 1, seed=22), _rand(1, seed=23), \
        _rand(c, seed=24), _rand(b, c, seed=25)
    blurred = ops.upfirdn2d(x.cpu(), taps.cpu(), pad=(1, 1)).cuda()
    want = F.leaky_relu(blurred * demod[:, :, None, None] + nw * noise + bias[None, :, None, None], 0.2) * math.sqrt(2)
    xn = x.permute(0, 2, 3, 1).contiguous()
    got, got2 = ct.blur_nhwc(xn, taps, (1, 1), demod=demod, noise=noise.reshape(b, h - 1, h - 1).contiguous(),
                             noise_weight=nw, bias=bias, act=True, s_next=sn, want_out2=True)
    assert _err(got.permute(0, 3, 1, 2), want) < 1e-5
    assert _err(got2.permute(0, 3, 1, 2), want * sn[:, :, None, None]) < 1e-5
    only_mod = ct.blur_nhwc(xn, taps, (1, 1), demod=demod, s_next=sn)
    assert _err(only_mod.permute(0, 3, 1, 2), blurred * (demod * sn)[:, :, None, None]) < 1e-5
    plain = ct.blur_nhwc(xn, taps, (2, 2))
    assert _err(plain.permute(0, 3, 1, 2), ops.upfirdn2d(x.cpu(), taps.cpu(), pad=(2, 2)).cuda()) < 1e-5


@pytest.mark.parametrize("b,c,h,w", [(2, 128, 16, 16), (3, 512, 4, 4), (2, 64, 8, 6), (1, 36, 5, 3), (2, 256, 33, 1)])
def test_to_rgb_nhwc(b, c, h, w):
    """4-pixel butterfly kernel (H*W % 4 == 0) and the one-pixel kernel (odd pixel counts), ragged channel counts."""
    from rick_b200 import conv_tc as ct
    y, wmod, bias, skip = _rand(b, h, w, c, seed=30), _rand(b, 3, c, seed=31), _rand(3, seed=32), _rand(b, 3, h, w, seed=33)
    want = torch.einsum("bhwc,boc->bohw", y.double(), wmod.double()).float() + bias[None, :, None, None]
    assert _err(ct.to_rgb_nhwc(y, wmod, bias, None), want) < 1e-5
    assert _err(ct.to_rgb_nhwc(y, wmod, bias, skip), want + skip) < 1e-5


@pytest.mark.parametrize("size", [32, 256])
def test_fused_generator_matches_module_and_golden(size, golden):
    import numpy as np
    from oracle import synth
    from rick_b200 import stylegan2 as sg
    from rick_b200.fused import FusedGenerator
    seed = 11 if size == 32 else 1
    G = sg.Generator(size, 512, 8)
    G.load_state_dict(synth.g_state(size, seed))
    G = G.cuda()
    F_ = FusedGenerator(G)
    if size == 32:
        z = [synth.latents(2, 21).cuda()]
        ref = torch.from_numpy(golden("model32_golden.npz")["img"])
        sub = lambda t: t
    else:
        import os
        from conftest import ROOT
        z = [torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "fisher_latents.npy")))[:2].cuda()]
        ref = torch.from_numpy(golden("g256_golden.npz")["img_sub4"])
        sub = lambda t: t[:, :, ::4, ::4]
    with torch.no_grad():
        img_mod, _ = G(z, randomize_noise=False)
    img, none = F_(z, randomize_noise=False)
    assert none is None and img.shape == img_mod.shape
    scale = max(1.0, ref.abs().max().item())
    assert (sub(img).cpu() - ref).abs().max().item() / scale < 1e-2          # vs the reference's own output
    assert (img - img_mod).abs().max().item() / scale < 1e-2                 # vs the library-conv module path
    # style mixing and random noise paths run
    z2 = [z[0], z[0].flip(0)]
    a, lat = F_(z2, inject_index=3, randomize_noise=False, return_latents=True)
    with torch.no_grad():
        bref, _ = G(z2, inject_index=3, randomize_noise=False)
    assert lat.shape == (2, G.n_latent, 512)
    assert (a - bref).abs().max().item() / scale < 1e-2
    c, _ = F_(z, randomize_noise=True)
    assert torch.isfinite(c).all()

"""The differentiated convolutions on the tcgen05 kernels: forward on in-place channels-last weights, data gradient
(the same weight memory read as an MN-major operand), weight gradient (both operands MN-major, split over pixels,
deterministic fold) -- each against fp32 library results, then the autograd Functions of rick_b200.conv to second order.

TF32 operands (truncated to 10 mantissa bits, bias-compensated in the epilogue; fp32 accumulate): tolerance 2e-3 of the
result's scale, the bound VERDICT r1 set for dgrad / wgrad at every G / D shape."""
import math

import pytest
import torch
from torch.nn import functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _err(got, want):
    return ((got.double() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-30)).item()


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).cuda()


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _fp32(fn):
    """run ``fn`` with TF32 off: the fp32 library reference"""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        return fn()
    finally:
        torch.backends.cudnn.allow_tf32 = prev


# B, H, W, Cin, Cout, k, stride, pad, transposed -- the shapes G and D run at 256 px (batch 2 / joint batch 4), their
# 1024 px small-channel layers, and ragged ones
SHAPES = [
    (2, 4, 4, 512, 512, 3, 1, 1, False),
    (2, 8, 8, 512, 512, 3, 1, 1, False),
    (2, 32, 32, 512, 512, 3, 1, 1, False),
    (4, 64, 64, 512, 512, 3, 1, 1, False),
    (2, 128, 128, 256, 256, 3, 1, 1, False),
    (2, 256, 256, 128, 128, 3, 1, 1, False),
    (4, 129, 129, 256, 512, 3, 2, 0, False),      # D conv2 after its blur
    (4, 257, 257, 128, 256, 3, 2, 0, False),
    (4, 9, 9, 512, 512, 3, 2, 0, False),
    (4, 127, 127, 128, 256, 1, 2, 0, False),      # D skip: 1x1 stride 2 on the blurred map
    (2, 7, 7, 512, 512, 1, 2, 0, False),
    (4, 4, 4, 544, 512, 3, 1, 1, False),          # D final conv, 513 channels padded to 544 (three Cin blocks of 192)
    (2, 4, 4, 512, 512, 3, 2, 0, True),           # G upsampling convs (transposed, polyphase)
    (2, 32, 32, 512, 512, 3, 2, 0, True),
    (2, 64, 64, 512, 256, 3, 2, 0, True),
    (2, 128, 128, 256, 128, 3, 2, 0, True),
    (1, 64, 64, 128, 64, 3, 2, 0, True),          # 512 / 1024 px generators: 64- and 32-channel layers (partial M tiles)
    (1, 96, 96, 64, 64, 3, 1, 1, False),
    (1, 64, 64, 64, 32, 3, 2, 0, True),
    (1, 100, 100, 32, 32, 3, 1, 1, False),
    (3, 5, 7, 32, 96, 3, 1, 1, False),            # ragged everything
    (1, 11, 6, 96, 32, 3, 2, 0, False),
    (1, 20, 12, 64, 128, 1, 1, 0, False),
]
IDS = ["_".join(str(int(v)) for v in c) for c in SHAPES]


def _problem(case, seed=0):
    b, h, w, cin, cout, k, stride, pad, transposed = case
    x = _cl(_rand(b, cin, h, w, seed=seed + 1))
    wgt = _cl(_rand(cout, cin, k, k, seed=seed + 2) / math.sqrt(cin * k * k))
    return x, wgt


def _ref(x, wgt, case):
    b, h, w, cin, cout, k, stride, pad, transposed = case
    if transposed:
        return F.conv_transpose2d(x, wgt.transpose(0, 1), stride=2)
    return F.conv2d(x, wgt, stride=stride, padding=pad)


@pytest.mark.parametrize("case", SHAPES, ids=IDS)
def test_fprop_dgrad_wgrad_vs_fp32(case):
    from rick_b200 import conv
    b, h, w, cin, cout, k, stride, pad, transposed = case
    cfg = (stride, pad, transposed)
    x, wgt = _problem(case)
    xr, wr = x.clone().requires_grad_(True), wgt.clone().requires_grad_(True)
    want = _fp32(lambda: _ref(xr, wr, case))
    go = _cl(_rand(*want.shape, seed=9))
    want_gx, want_gw = _fp32(lambda: torch.autograd.grad(want, [xr, wr], go))
    before = dict(conv.launch_stats)
    got = conv._fprop(x, wgt, cfg)
    gx = conv._dgrad(go, wgt, x, cfg)
    gw = conv._wgrad(go, x, wgt, cfg)
    assert conv.launch_stats["tc"] - before["tc"] == 3 and conv.launch_stats["library"] == before["library"], \
        "all three primitives must run on the tcgen05 kernels for this shape"
    assert got.shape == want.shape and gx.shape == x.shape and gw.shape == wgt.shape
    assert gw.stride() == wgt.stride(), "the weight gradient arrives in the parameter's own layout"
    assert _err(got, want) < TOL, "fprop"
    assert _err(gx, want_gx) < TOL, "dgrad"
    assert _err(gw, want_gw) < TOL, "wgrad"


def test_wgrad_is_deterministic_and_scales():
    from rick_b200 import conv_tc as ct
    b, h, w, cin, cout = 4, 64, 64, 128, 128
    x, g = _rand(b, h, w, cin, seed=1), _rand(b, h, w, cout, seed=2)
    like = _cl(torch.empty(cout, cin, 3, 3, device="cuda"))
    geom = ct.geom_wgrad(b, h, w, cin, cout, 3, 1, 1)
    a = ct.conv_wgrad_tc(g, x, geom, like)
    for _ in range(3):
        assert torch.equal(ct.conv_wgrad_tc(g, x, geom, like), a)
    torch.testing.assert_close(ct.conv_wgrad_tc(g, x, geom, like, scale=0.25), a * 0.25, rtol=1e-6, atol=0)


def test_wgrad_writes_any_weight_layout():
    """the fold kernel writes with the caller's strides: a contiguous (Cout, Cin, k, k) destination works as well"""
    from rick_b200 import conv_tc as ct
    b, h, w, cin, cout = 2, 16, 16, 64, 128
    x, g = _rand(b, h, w, cin, seed=3), _rand(b, h, w, cout, seed=4)
    geom = ct.geom_wgrad(b, h, w, cin, cout, 3, 1, 1)
    cl = ct.conv_wgrad_tc(g, x, geom, _cl(torch.empty(cout, cin, 3, 3, device="cuda")))
    plain = ct.conv_wgrad_tc(g, x, geom, torch.empty(cout, cin, 3, 3, device="cuda"))
    assert plain.is_contiguous() and torch.equal(plain, cl)


@pytest.mark.parametrize("case", [(2, 16, 16, 64, 128, 3, 1, 1, False), (2, 17, 17, 64, 128, 3, 2, 0, False),
                                  (2, 8, 8, 128, 64, 3, 2, 0, True), (2, 9, 9, 32, 64, 1, 2, 0, False)],
                         ids=lambda c: "_".join(str(int(v)) for v in c))
def test_autograd_first_and_second_order_vs_fp32(case):
    """rick_b200.conv._lib_conv (the Function every module convolution goes through) against torch's own convolution:
    value, first derivatives, and second derivatives of a gradient-penalty-shaped scalar (what R1 / path-length need)."""
    from rick_b200 import conv
    b, h, w, cin, cout, k, stride, pad, transposed = case

    def run(fn):
        x, wgt = _problem(case, seed=20)
        x.requires_grad_(True), wgt.requires_grad_(True)
        y = fn(x, wgt)
        go = _cl(_rand(*y.shape, seed=29))
        (gx,) = torch.autograd.grad((y * go).sum(), x, create_graph=True)
        penalty = gx.pow(2).sum() + (y * y).mean()
        d_w, d_x = torch.autograd.grad(penalty, [wgt, x])
        return y.detach(), gx.detach(), d_w, d_x

    want = _fp32(lambda: run(lambda x, wgt: _ref(x, wgt, case)))
    before = dict(conv.launch_stats)
    got = run(lambda x, wgt: conv._lib_conv(x, wgt, stride, pad, transposed))
    assert conv.launch_stats["library"] == before["library"], "no library convolution may be involved"
    for name, a, b_ in zip(("y", "gx", "d penalty / d w", "d penalty / d x"), got, want):
        assert _err(a, b_) < 2 * TOL, name


def test_modulated_conv_gradients_flow_through_tc_kernels():
    """ModulatedConv2d / StyledConv as the generator runs them in training: gradients w.r.t. weight, style and input
    against the same module on fp32 library convolutions; every convolution launch on the tcgen05 path."""
    from oracle import synth
    from rick_b200 import conv
    from rick_b200 import stylegan2 as sg
    torch.manual_seed(0)
    for upsample in (False, True):
        m = sg.StyledConv(64, 128, 3, 512, upsample=upsample).cuda()
        m.noise.weight.data.fill_(0.3)
        x = _cl(_rand(2, 64, 16, 16, seed=40)).requires_grad_(True)
        style = _rand(2, 512, seed=41).requires_grad_(True)
        o = 2 * 16 if upsample else 16
        noise = _rand(2, 1, o, o, seed=42)

        def grads():
            y = m(x, style, noise=noise)
            return [y.detach()] + list(torch.autograd.grad(y.pow(2).mean(), [m.conv.weight, style, x, m.conv.modulation.weight]))

        want = _fp32(grads)
        before = dict(conv.launch_stats)
        got = grads()
        assert conv.launch_stats["library"] == before["library"] and conv.launch_stats["tc"] > before["tc"]
        for name, a, b_ in zip(("y", "d weight", "d style", "d x", "d modulation"), got, want):
            assert _err(a, b_) < 2 * TOL, (upsample, name)


def test_discriminator_convs_avoid_the_library():
    """One D forward / backward at 64 px: only the 3-channel from-RGB layer (broadcast form) is not a tcgen05 launch."""
    from oracle import synth
    from rick_b200 import conv
    from rick_b200 import stylegan2 as sg
    D = sg.Discriminator(64)
    D.load_state_dict(synth.d_state(64, 2))
    D = D.cuda()
    x = synth.shots(4, 64, 1).cuda()
    before = dict(conv.launch_stats)
    out, _ = D(x)
    F.softplus(out).mean().backward()
    assert conv.launch_stats["library"] == before["library"], "a D convolution fell back to the library"
    assert conv.launch_stats["tc"] - before["tc"] >= 3 * (3 * 4 + 1) - 1      # fprop / dgrad / wgrad of 13 convs (no dgrad into from-RGB input needed)


@pytest.mark.parametrize("b,h,cin,cout,k,stride,pad", [(2, 4, 512, 512, 3, 1, 1), (4, 8, 512, 512, 3, 1, 1),
                                                       (4, 16, 512, 512, 3, 1, 1), (4, 9, 512, 512, 3, 2, 0),
                                                       (2, 8, 128, 64, 3, 1, 1), (4, 4, 544, 512, 3, 1, 1)])
def test_split_k_equals_unsplit_with_fused_epilogue(b, h, cin, cout, k, stride, pad):
    """Small maps: the K loop is divided over otherwise idle SMs, partial sums folded in a fixed order with the epilogue
    (demod, noise, bias, leaky-ReLU, pre-modulated second output) applied by the fold kernel.  Must equal the one-kernel
    path up to fp32 summation order, be deterministic, and actually be in use for these shapes."""
    import ctypes
    from rick_b200 import _lib
    from rick_b200 import conv_tc as ct
    x = _rand(b, h, h, cin, seed=1)
    wgt = _cl(_rand(cout, cin, k, k, seed=2) / math.sqrt(cin * k * k))
    geom = ct.geom_conv(b, h, h, cin, cout, k, stride, pad)
    assert _lib.lib().rick_conv_tc_workspace(ctypes.byref(geom)) > 0, "expected a split-K plan for this shape"
    oh = geom.out_h
    kw = dict(demod=_rand(b, cout, seed=3).abs() + 0.5, noise=_rand(b, oh, oh, seed=4), noise_weight=_rand(1, seed=5),
              bias=_rand(cout, seed=6), act=True, s_next=_rand(b, cout, seed=7), want_out2=True)
    ct.SPLIT_K = False
    try:
        want, want2 = ct.conv_tc_nhwc(x, wgt, geom, **kw)
        plain_want = ct.conv_tc_nhwc(x, wgt, geom)
    finally:
        ct.SPLIT_K = True
    got, got2 = ct.conv_tc_nhwc(x, wgt, geom, **kw)
    # fp32 summation order differs (K ranges are added after the fact): a few ulps of the output scale
    assert _err(got, want) < 5e-5 and _err(got2, want2) < 5e-5
    assert _err(ct.conv_tc_nhwc(x, wgt, geom), plain_want) < 5e-5
    again, again2 = ct.conv_tc_nhwc(x, wgt, geom, **kw)
    assert torch.equal(again, got) and torch.equal(again2, got2)


def test_split_k_polyphase_and_transposed_weight():
    """split-K through the 4-phase transposed convolution and through an MN-major (data-gradient) weight operand"""
    from rick_b200 import conv_tc as ct
    b, h, cin, cout = 2, 8, 512, 512
    x = _rand(b, h, h, cin, seed=11)
    wgt = _cl(_rand(cout, cin, 3, 3, seed=12) / math.sqrt(cin * 9))
    for geom, tr, xin in ((ct.geom_conv_transpose_s2(b, h, h, cin, cout), False, x),
                          (ct.geom_conv_dgrad(b, h, h, cin, cout, 3, 1, 1), True, _rand(b, h, h, cout, seed=13)),
                          (ct.geom_conv_dgrad(b, 2 * h + 1, 2 * h + 1, cin, cout, 3, 2, 0), True, _rand(b, h, h, cout, seed=14))):
        ct.SPLIT_K = False
        try:
            want = ct.conv_tc_nhwc(xin, wgt, geom, transpose_weight=tr)
        finally:
            ct.SPLIT_K = True
        got = ct.conv_tc_nhwc(xin, wgt, geom, transpose_weight=tr)
        assert _err(got, want) < 5e-5

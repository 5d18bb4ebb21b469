"""Parity of the sm_100a upfirdn2d / bias-act kernels (called through the C ABI) against the CPU oracle and the
reference-generated golden vectors.  Tolerance (BASELINE.json north_star): fp32 within 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import ops_oracle as ops
from oracle import synth

pytestmark = pytest.mark.gpu
REL = 1e-5


def _close(got, want, rel=REL):
    want = want.to(torch.float64)
    got = got.detach().cpu().to(torch.float64)
    scale = max(want.abs().max().item(), 1e-30)
    err = (got - want).abs().max().item()
    assert err <= rel * scale, f"max-abs err {err:.3e} > {rel:.0e} * {scale:.3e}"


@pytest.fixture(scope="module")
def op():
    from rick_b200 import op as _op
    return _op


@pytest.mark.parametrize("case", synth.UPFIRDN_CASES, ids=[c[0] for c in synth.UPFIRDN_CASES])
def test_upfirdn2d_golden(case, golden, op):
    name, n, c, h, w, kh, kw, up, down, p0, p1, kind = case
    x, taps = synth.upfirdn_case_inputs(case)
    got = op.upfirdn2d(x.cuda(), taps.cuda(), up, down, (p0, p1))
    want = torch.from_numpy(golden("ops_golden.npz")[name])
    assert tuple(got.shape) == tuple(want.shape)
    _close(got, want)


MODEL_SHAPES = [
    # (N, C, H, W, up, down, pad, gain)  -- every call shape of G / D at 256 px, plus ragged / multi-tile ones
    (2, 512, 9, 9, 1, 1, (1, 1), 4), (2, 512, 17, 17, 1, 1, (1, 1), 4), (2, 512, 33, 33, 1, 1, (1, 1), 4),
    (2, 512, 65, 65, 1, 1, (1, 1), 4), (2, 256, 129, 129, 1, 1, (1, 1), 4), (2, 128, 257, 257, 1, 1, (1, 1), 4),
    (2, 3, 4, 4, 2, 1, (2, 1), 4), (2, 3, 32, 32, 2, 1, (2, 1), 4), (2, 3, 128, 128, 2, 1, (2, 1), 4),
    (2, 128, 256, 256, 1, 1, (2, 2), 1), (2, 128, 256, 256, 1, 1, (1, 1), 1), (2, 512, 8, 8, 1, 1, (2, 2), 1),
    (2, 512, 8, 8, 1, 1, (1, 1), 1), (3, 7, 64, 64, 1, 2, (1, 1), 1), (1, 5, 37, 53, 2, 1, (2, 1), 4),
    (1, 5, 37, 53, 1, 2, (2, 1), 1), (1, 4, 130, 259, 1, 1, (2, 2), 1), (1, 2, 300, 140, 2, 1, (1, 2), 4),
    (1, 2, 31, 31, 2, 1, (-1, 3), 4), (1, 2, 40, 40, 1, 1, (-1, -2), 1),
]


@pytest.mark.parametrize("shape", MODEL_SHAPES, ids=[f"{s[0]}x{s[1]}x{s[2]}x{s[3]}_u{s[4]}d{s[5]}p{s[6][0]}{s[6][1]}" for s in MODEL_SHAPES])
def test_upfirdn2d_model_shapes_fwd_bwd(shape, op):
    n, c, h, w, up, down, pad, gain = shape
    g = torch.Generator().manual_seed(h * 1000 + w + up * 7 + down)
    x = torch.randn(n, c, h, w, generator=g)
    taps = torch.tensor([1., 3., 3., 1.])
    taps = torch.outer(taps, taps)
    taps = taps / taps.sum() * gain
    xo = x.clone().requires_grad_(True)
    want = ops.upfirdn2d(xo, taps, up, down, pad)
    xg = x.cuda().requires_grad_(True)
    got = op.upfirdn2d(xg, taps.cuda(), up, down, pad)
    assert tuple(got.shape) == tuple(want.shape)
    _close(got, want.detach())
    go = torch.randn(want.shape, generator=g)
    (gw,) = torch.autograd.grad(want, xo, go)
    (gg,) = torch.autograd.grad(got, xg, go.cuda())
    _close(gg, gw)


ROWS_SHAPES = [
    # (N, C, H, W, down, pad) large up=1 planes -> the bulk-staged rows kernel: odd widths, left-over column counts of
    # 0 / 1 / 4 (edge columns) / 6 / 31 (partial group), ragged last row tile, crops, 1024-px-D widths, down = 2
    (2, 3, 129, 129, 1, (1, 1)), (2, 3, 128, 128, 1, (2, 2)), (1, 2, 71, 73, 1, (1, 1)), (1, 2, 140, 101, 1, (2, 2)),
    (1, 2, 90, 98, 1, (-1, -2)), (1, 1, 513, 513, 1, (1, 1)), (1, 1, 40, 1025, 1, (2, 2)), (2, 3, 128, 128, 2, (1, 1)),
    (1, 2, 257, 259, 2, (2, 2)), (1, 2, 150, 131, 2, (0, 3)), (3, 1, 67, 200, 1, (3, 0)), (1, 1, 512, 512, 2, (1, 1)),
    (1, 1, 100, 600, 1, (1, 1)), (1, 1, 90, 1000, 2, (2, 2)),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("shape", ROWS_SHAPES, ids=[f"{s[0]}x{s[1]}x{s[2]}x{s[3]}_d{s[4]}p{s[5][0]}{s[5][1]}" for s in ROWS_SHAPES])
def test_upfirdn2d_rows_kernel_asymmetric_taps(shape, dtype, op):
    """Random (asymmetric) 4x4 taps so a transposed / flipped window cannot pass; forward and backward (flipped taps)."""
    n, c, h, w, down, pad = shape
    g = torch.Generator().manual_seed(h * 977 + w * 13 + down)
    x = torch.randn(n, c, h, w, generator=g).to(dtype)
    taps = torch.randn(4, 4, generator=g)
    xo = x.float().clone().requires_grad_(True)
    want = ops.upfirdn2d(xo, taps, 1, down, pad)
    xg = x.cuda().requires_grad_(True)
    got = op.upfirdn2d(xg, taps.cuda(), 1, down, pad)
    assert got.dtype == dtype and tuple(got.shape) == tuple(want.shape)
    rel = REL if dtype == torch.float32 else 1e-2
    _close(got.float(), want.detach(), rel)
    go = torch.randn(want.shape, generator=g).to(dtype)
    (gw,) = torch.autograd.grad(want, xo, go.float())
    (gg,) = torch.autograd.grad(got, xg, go.cuda())
    _close(gg.float(), gw, rel)


LONG_TAPS = [
    # (N, C, H, W, kh, kw, up, down, pad, separable): the separable tiled kernel (5..16 taps per axis, up/down in {1,2})
    (2, 3, 70, 90, 12, 12, 2, 1, (6, 5), True), (2, 3, 141, 77, 12, 12, 1, 2, (5, 5), True),
    (1, 2, 64, 64, 12, 12, 1, 1, (3, 8), True), (1, 2, 33, 130, 7, 11, 2, 1, (3, 2), True),
    (1, 2, 50, 41, 16, 5, 1, 2, (7, 0), True), (1, 3, 40, 40, 12, 12, 2, 1, (-3, 9), True),
    (1, 2, 45, 67, 12, 12, 2, 1, (6, 5), False), (1, 2, 90, 70, 9, 9, 1, 2, (4, 4), False),
    (1, 1, 100, 100, 6, 6, 1, 1, (-2, -1), False), (2, 3, 282, 282, 12, 12, 2, 1, (0, 0), True),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("case", LONG_TAPS, ids=[f"{c[2]}x{c[3]}_k{c[4]}x{c[5]}_u{c[6]}d{c[7]}_{'sep' if c[9] else '2d'}" for c in LONG_TAPS])
def test_upfirdn2d_long_taps_separable_kernel(case, dtype, op):
    """non_leaking.py:321-359 (12 x 12 outer-product antialiasing taps, up = 2 / down = 2) and other long filters: rank-1
    kernels go through the in-kernel factorisation + two 1-D passes, anything else through the 2-D loop of the same
    kernel; forward and the gradient (the adjoint runs the flipped taps with up <-> down)."""
    n, c, h, w, kh, kw, up, down, pad, sep = case
    g = torch.Generator().manual_seed(kh * 100 + kw + h)
    x = torch.randn(n, c, h, w, generator=g)
    if sep:
        taps = torch.outer(torch.randn(kh, generator=g), torch.randn(kw, generator=g))
    else:
        taps = torch.randn(kh, kw, generator=g)
    rel = REL if dtype == torch.float32 else 2e-2
    xo = x.to(dtype).float().requires_grad_(True)
    want = ops.upfirdn2d(xo, taps, up, down, pad)
    xg = x.to(dtype).cuda().requires_grad_(True)
    got = op.upfirdn2d(xg, taps.cuda(), up, down, pad)
    assert tuple(got.shape) == tuple(want.shape) and got.dtype == dtype
    _close(got.float(), want.detach(), rel)
    if dtype == torch.float32:
        go = torch.randn(want.shape, generator=g)
        (gw,) = torch.autograd.grad(want, xo, go)
        (gg,) = torch.autograd.grad(got, xg, go.cuda())
        _close(gg, gw, rel)


def test_upfirdn2d_double_backward(op):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 12, 12, generator=g)
    taps = torch.randn(4, 4, generator=g)
    for up, down, pad in [(2, 1, (2, 1)), (1, 2, (1, 1)), (1, 1, (2, 2))]:
        def second_order(fn, xin, t):
            xin = xin.clone().requires_grad_(True)
            y = fn(xin, t, up, down, pad)
            (gx,) = torch.autograd.grad((y ** 2).sum(), xin, create_graph=True)
            (ggx,) = torch.autograd.grad((gx ** 3).sum(), xin)
            return gx.detach(), ggx
        gw, ggw = second_order(ops.upfirdn2d, x, taps)
        gg, ggg = second_order(op.upfirdn2d, x.cuda(), taps.cuda())
        _close(gg, gw, 1e-4)
        _close(ggg, ggw, 1e-4)


def test_upfirdn2d_channels_last_and_bf16(op):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 16, 20, 20, generator=g)
    taps = torch.randn(4, 4, generator=g)
    want = ops.upfirdn2d(x, taps, 2, 1, (2, 1))
    got = op.upfirdn2d(x.cuda().to(memory_format=torch.channels_last), taps.cuda(), 2, 1, (2, 1))
    assert got.is_contiguous(memory_format=torch.channels_last)
    _close(got, want)
    xb = x.to(torch.bfloat16)
    wantb = ops.upfirdn2d(xb.float(), taps, 1, 1, (1, 1))
    gotb = op.upfirdn2d(xb.cuda(), taps.cuda(), 1, 1, (1, 1))
    assert gotb.dtype == torch.bfloat16
    _close(gotb.float(), wantb, 1e-2)     # one bf16 rounding of the output


def test_upfirdn2d_large_properties(op):
    """BASELINE op-sweep size (32, 512, 128, 128) -> (32, 512, 256, 256): checked through size-independent properties
    -- impulse response, linearity, and a strided sample against the oracle."""
    taps = torch.tensor([1., 3., 3., 1.])
    taps = (torch.outer(taps, taps) / 64 * 4).cuda()
    x = torch.randn(32, 512, 128, 128, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    y = op.upfirdn2d(x, taps, 2, 1, (2, 1))
    assert y.shape == (32, 512, 256, 256)
    # up=2 with gain-4 [1,3,3,1] taps preserves the plane mean up to the border
    assert torch.allclose(y.mean(), x.mean(), atol=5e-4)
    # linearity
    y2 = op.upfirdn2d(x * 0.5 + 1.0, taps, 2, 1, (2, 1))
    ones = op.upfirdn2d(torch.ones_like(x[:1, :1]), taps, 2, 1, (2, 1))
    assert torch.allclose(y2, 0.5 * y + ones, atol=2e-5)
    # sampled planes against the oracle
    idx = [(0, 0), (7, 300), (31, 511)]
    sub = torch.stack([x[a, b] for a, b in idx])[None].cpu()
    want = ops.upfirdn2d(sub, taps.cpu(), 2, 1, (2, 1))[0]
    got = torch.stack([y[a, b] for a, b in idx])
    _close(got, want)


def test_upfirdn2d_errors(op):
    taps = torch.ones(4, 4).cuda()
    with pytest.raises(RuntimeError):
        op.upfirdn2d(torch.ones(1, 1, 4, 4), taps)                      # CPU input: no fallback
    with pytest.raises(RuntimeError):
        op.upfirdn2d(torch.ones(1, 1, 2, 2).cuda(), taps, 1, 1, (0, 0))  # empty output
    with pytest.raises(RuntimeError):
        op.upfirdn2d(torch.ones(1, 1, 4, 4, dtype=torch.float64).cuda(), taps)


BIAS_SHAPES = [(2, 512), (2, 512, 4, 4), (2, 512, 32, 32), (2, 128, 256, 256), (3, 5, 7, 9), (1, 3, 1, 1), (4, 6, 2, 2),
               (2, 4, 100, 100)]


@pytest.mark.parametrize("shape", BIAS_SHAPES, ids=["x".join(map(str, s)) for s in BIAS_SHAPES])
def test_fused_leaky_relu_fwd_bwd_gradgrad(shape, op):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    go = torch.randn(*shape, generator=g)

    def run(fn, dev):
        xi = x.to(dev).requires_grad_(True)
        bi = b.to(dev).requires_grad_(True)
        gi = go.to(dev).requires_grad_(True)
        y = fn(xi, bi)
        gx, gb = torch.autograd.grad(y, [xi, bi], gi, create_graph=True)
        # double backward: (gx, gb) are linear in the upstream gradient gi; differentiate w.r.t. it
        (ggo,) = torch.autograd.grad((gx * gx).sum() + (gb * gb).sum(), gi)
        return y.detach(), gx.detach(), gb.detach(), ggo

    wy, wgx, wgb, wgg = run(ops.fused_leaky_relu, "cpu")
    gy, ggx, ggb, ggg = run(op.fused_leaky_relu, "cuda")
    _close(gy, wy)
    _close(ggx, wgx)
    _close(ggb, wgb, 1e-4)          # a sum of up to N*H*W terms in a different (fixed) order
    _close(ggg, wgg, 1e-4)


def test_bias_act_native_modes(op):
    from rick_b200.op.fused_act import bias_act
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 6, 5, 8, generator=g)
    b = torch.randn(6, generator=g)
    ref = torch.randn(2, 6, 5, 8, generator=g)
    for act, grad in [(3, 0), (3, 1), (3, 2), (1, 0)]:
        want = ops.bias_act(x, b, ref, act, grad, 0.2, 2 ** 0.5)
        got = bias_act(x.cuda(), b.cuda(), ref.cuda(), act, grad, 0.2, 2 ** 0.5)
        _close(got, want)
    # absent bias / ref, as the reference's empty tensors
    _close(bias_act(x.cuda(), None, ref.cuda(), 3, 1, 0.2, 1.0), ops.bias_act(x, None, ref, 3, 1, 0.2, 1.0))
    _close(bias_act(x.cuda(), x.new_empty(0).cuda(), None, 3, 0, 0.2, 1.0), ops.bias_act(x, None, None, 3, 0, 0.2, 1.0))


def test_bias_act_bwd_is_deterministic(op):
    x = torch.randn(4, 64, 64, 64, device="cuda")
    b = torch.randn(64, device="cuda")
    go = torch.randn_like(x)
    outs = []
    for _ in range(3):
        xi, bi = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
        gx, gb = torch.autograd.grad(op.fused_leaky_relu(xi, bi), [xi, bi], go)
        outs.append((gx, gb))
    assert all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])


def test_fused_leaky_relu_bf16(op):
    x = torch.randn(2, 16, 8, 8).to(torch.bfloat16)
    b = torch.randn(16).to(torch.bfloat16)
    want = ops.fused_leaky_relu(x.float(), b.float())
    got = op.fused_leaky_relu(x.cuda(), b.cuda())
    assert got.dtype == torch.bfloat16
    _close(got.float(), want, 1e-2)


@pytest.mark.parametrize("pad", [(1, 1), (2, 2), (2, 1)])
def test_upfirdn2d_channels_last_blur_fwd_bwd_gradgrad(pad, op):
    """channels-last 4x4 blur goes through rick_blur_nhwc; all derivative orders against the oracle."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 16, 13, 9, generator=g)
    taps = torch.randn(4, 4, generator=g)            # asymmetric taps: catches a wrong flip in the adjoint

    def run(fn, xin, t):
        xin = xin.requires_grad_(True)
        y = fn(xin, t, 1, 1, pad)
        (gx,) = torch.autograd.grad((y ** 2).sum(), xin, create_graph=True)
        (ggx,) = torch.autograd.grad((gx ** 3).sum(), xin)
        return y.detach(), gx.detach(), ggx

    wy, wg, wgg = run(ops.upfirdn2d, x.clone(), taps)
    xc = x.cuda().to(memory_format=torch.channels_last)
    gy, gg, ggg = run(op.upfirdn2d, xc, taps.cuda())
    assert gy.is_contiguous(memory_format=torch.channels_last)
    _close(gy, wy)
    _close(gg, wg, 1e-4)
    _close(ggg, wgg, 1e-4)


@pytest.mark.parametrize("shape", [(2, 512, 4, 4), (2, 128, 64, 64), (3, 12, 7, 9), (2, 256, 33, 33)],
                         ids=lambda s: "x".join(map(str, s)))
def test_fused_leaky_relu_channels_last(shape, op):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    go = torch.randn(*shape, generator=g)

    def run(fn, dev, cl):
        conv = (lambda t: t.to(dev).to(memory_format=torch.channels_last)) if cl else (lambda t: t.to(dev))
        xi = conv(x).requires_grad_(True)
        bi = b.to(dev).requires_grad_(True)
        gi = conv(go).requires_grad_(True)
        y = fn(xi, bi)
        gx, gb = torch.autograd.grad(y, [xi, bi], gi, create_graph=True)
        (ggo,) = torch.autograd.grad((gx * gx).sum() + (gb * gb).sum(), gi)
        return y, gx.detach(), gb.detach(), ggo

    wy, wgx, wgb, wgg = run(ops.fused_leaky_relu, "cpu", False)
    gy, ggx, ggb, ggg = run(op.fused_leaky_relu, "cuda", True)
    assert gy.is_contiguous(memory_format=torch.channels_last) and ggx.is_contiguous(memory_format=torch.channels_last)
    _close(gy, wy.detach())
    _close(ggx, wgx)
    _close(ggb, wgb, 1e-4)
    _close(ggg, wgg, 1e-4)


@pytest.mark.parametrize("shape", [(2, 512, 4, 4), (2, 128, 37, 29), (3, 12, 9, 9), (2, 256, 64, 64)],
                         ids=lambda s: "x".join(map(str, s)))
def test_fused_modulate_and_styled_epilogue(shape):
    """rick_modulate_* / rick_styled_epilogue_* (fwd, fused bwd with all broadcast-gradient reductions, and the
    differentiable double-backward branch) against the plain torch composite on the CPU in float64."""
    from rick_b200.op import styled
    b, c, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    x, s = torch.randn(b, c, h, w, generator=g), torch.randn(b, c, generator=g)
    d = torch.rand(b, c, generator=g) + 0.5
    noise, nw, bias = torch.randn(b, 1, h, w, generator=g), torch.randn(1, generator=g), torch.randn(c, generator=g)
    go = torch.randn(b, c, h, w, generator=g)

    def composite(x, s, d, noise, nw, bias):
        a = x * s[:, :, None, None]
        return torch.nn.functional.leaky_relu(a * d[:, :, None, None] + nw * noise + bias[None, :, None, None], 0.2) * 2 ** 0.5

    def fused(x, s, d, noise, nw, bias):
        return styled.styled_epilogue(styled.modulate(x, s), d, noise, nw, bias)

    def run(fn, dev, dt):
        leaves = [t.to(dev, dt).requires_grad_(True) for t in (x, s, d, nw, bias)]
        xi, si, di, nwi, bi = leaves
        y = fn(xi, si, di, noise.to(dev, dt), nwi, bi)
        gi = go.to(dev, dt).requires_grad_(True)
        grads = torch.autograd.grad(y, leaves, gi, create_graph=True)
        # second order: differentiate a scalar of the first-order grads w.r.t. x, s and the upstream gradient
        scal = sum((gr * gr).sum() for gr in grads)
        gg = torch.autograd.grad(scal, [xi, si, gi], allow_unused=True)
        return y.detach(), [gr.detach() for gr in grads], gg

    wy, wg, wgg = run(composite, "cpu", torch.float64)
    gy, gg_, ggg = run(fused, "cuda", torch.float32)
    _close(gy, wy)
    for a_, b_ in zip(gg_, wg):
        _close(a_, b_, 2e-4)
    for a_, b_ in zip(ggg, wgg):
        assert (a_ is None) == (b_ is None)
        if a_ is not None:
            _close(a_, b_, 2e-3)
    # first-order fused kernels (no create_graph): same numbers as the composite branch
    leaves = [t.cuda().requires_grad_(True) for t in (x, s, d, nw, bias)]
    y = fused(leaves[0], leaves[1], leaves[2], noise.cuda(), leaves[3], leaves[4])
    grads = torch.autograd.grad(y, leaves, go.cuda())
    for a_, b_ in zip(grads, wg):
        _close(a_, b_, 2e-4)


def test_scale_all_matches_per_tensor_multiplies():
    """rick_scale_multi (all equalised-lr multipliers of a network in one launch): values, gradients, second-order."""
    from rick_b200.op.scale import scale_all
    g = torch.Generator().manual_seed(11)
    shapes = [(512, 512), (512,), (128, 64, 3, 3), (7, 5), (1, 513)]
    scales = [0.0442, 0.01, 1 / 24.0, 3.0, 1.0]
    ts = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    ts[2] = ts[2].detach().to(memory_format=torch.channels_last).requires_grad_(True)
    outs = scale_all(ts, scales)
    for o, t, s in zip(outs, ts, scales):
        assert o.stride() == t.stride()
        torch.testing.assert_close(o, t.detach() * s, rtol=0, atol=0)
    go = [torch.randn(s, generator=g).cuda() for s in shapes]
    grads = torch.autograd.grad(outs, ts, go)
    for gr, gg, s in zip(grads, go, scales):
        torch.testing.assert_close(gr, gg * s, rtol=0, atol=0)
    # second order through a nonlinearity on the outputs
    outs = scale_all(ts, scales)
    loss = sum((o ** 3).sum() for o in outs)
    g1 = torch.autograd.grad(loss, ts, create_graph=True)
    g2 = torch.autograd.grad(sum((a ** 2).sum() for a in g1), ts)
    for t, s, h in zip(ts, scales, g2):
        x = t.detach().double()
        want = 2 * (3 * s ** 3 * x ** 2) * (6 * s ** 3 * x)       # d/dx (3 s^3 x^2)^2
        torch.testing.assert_close(h.double(), want, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("shape,pad", [((2, 3, 64, 64), (2, 1)), ((1, 2, 70, 50), (2, 1)), ((1, 2, 40, 68), (1, 2)),
                                       ((1, 1, 33, 36), (3, 0))])
def test_upfirdn2d_up2_bf16_large(shape, pad):
    """bf16 storage on the large-map up = 2 kernel (8 outputs = one 16-byte store per thread and row), incl. widths that
    are not a multiple of 8 (scalar tail) and odd pad phases."""
    from rick_b200 import op
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(torch.bfloat16)
    taps = torch.randn(4, 4, generator=g)
    want = ops.upfirdn2d(x.float(), taps, 2, 1, pad)
    got = op.upfirdn2d(x.cuda(), taps.cuda(), 2, 1, pad)
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == tuple(want.shape)
    _close(got.float(), want, 1e-2)


# ------------------------------------------------------------------------------------------ fused glue ops
def test_weight_sqsum_any_layout_and_gradients():
    from rick_b200.op.glue import weight_sqsum
    g = torch.Generator().manual_seed(0)
    w = torch.randn(24, 20, 3, 3, generator=g).cuda()
    for t in (w, w.contiguous(memory_format=torch.channels_last), torch.randn(8, 12, 1, 1, generator=g).cuda()):
        t = t.clone().requires_grad_(True)
        got = weight_sqsum(t)
        want = t.pow(2).sum([2, 3])
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
        go = torch.randn_like(got)
        (ga,) = torch.autograd.grad(got, t, go, create_graph=True)
        (gb,) = torch.autograd.grad(want, t, go, create_graph=True)
        torch.testing.assert_close(ga, gb, rtol=1e-6, atol=1e-6)
        (ha,) = torch.autograd.grad(ga.pow(2).sum(), t)          # second order (path-length regularisation)
        (hb,) = torch.autograd.grad(gb.pow(2).sum(), t)
        torch.testing.assert_close(ha, hb, rtol=1e-5, atol=1e-6)


def test_linear_multi_matches_equal_linear_layers():
    """All modulation layers of a generator in one launch == the per-module EqualLinear results; weight / bias / latent
    gradients (first order through the fused wgrad kernel, and under create_graph through the composite branch)."""
    from rick_b200 import stylegan2 as sg
    torch.manual_seed(0)
    G = sg.Generator(32, 512, 8).cuda()
    for m, _ in G._modulated():
        m.modulation.bias.data.normal_()
    latent = torch.randn(2, G.n_latent, 512, device="cuda")
    plan = G._modulated()

    def fused(lat):
        return G.all_modulations(lat)

    def modular(lat):
        return [m.modulation(lat[:, i]) for m, i in plan]

    got = fused(latent)
    assert got is not None and len(got) == len(plan)
    for a, b in zip(got, modular(latent)):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
    params = [p for m, _ in plan for p in (m.modulation.weight, m.modulation.bias)]
    for create_graph in (False, True):
        res = []
        for fn in (fused, modular):
            lat = latent.clone().requires_grad_(True)
            outs = fn(lat)
            loss = sum((o * o).sum() * (k + 1) for k, o in enumerate(outs))
            gr = torch.autograd.grad(loss, [lat] + params, create_graph=create_graph)
            if create_graph:                                # differentiate once more, as the path-length penalty does
                gr = torch.autograd.grad(gr[0].pow(2).sum(), params, allow_unused=True)
            res.append(gr)
        for a, b in zip(*res):
            if a is None or b is None:
                assert a is None and b is None or (a if a is not None else b).abs().max() == 0
                continue
            torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-4)


def test_mapping_network_fused_forward():
    from rick_b200 import stylegan2 as sg
    torch.manual_seed(1)
    G = sg.Generator(32, 512, 8).cuda()
    for m in G.style:
        if hasattr(m, "bias"):
            m.bias.data.normal_()
    z = torch.randn(4, 512, device="cuda")
    with torch.no_grad():
        got = G.map_latent(z)
    want = G.style(z)                                           # autograd on: the module path
    torch.testing.assert_close(got, want.detach(), rtol=1e-4, atol=1e-5)


def test_from_rgb_fused_layer_matches_modules_to_second_order():
    """D's from-RGB ConvLayer as one fused pass each way: value, d/d image (kernel path), parameter gradients and the
    R1-style second derivative (composite path) against the EqualConv2d + FusedLeakyReLU modules it replaces."""
    from rick_b200 import stylegan2 as sg
    from rick_b200.op import glue
    torch.manual_seed(0)
    layer = sg.ConvLayer(3, 128, 1).cuda()
    layer[1].bias.data.normal_()
    conv, act = layer[0], layer[1]
    img = torch.randn(2, 3, 24, 20, device="cuda")
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        def fused(x):
            return glue.from_rgb(x, conv.weight, act.bias, conv.scale, act.negative_slope, act.scale)

        def modular(x):
            return layer(x)

        res = []
        for fn in (fused, modular):
            x = img.clone().requires_grad_(True)
            y = fn(x)
            go = torch.randn(2, 128, 24, 20, generator=torch.Generator().manual_seed(1)).cuda()
            (gx,) = torch.autograd.grad(y, x, go)                                           # first order, no graph
            x2 = img.clone().requires_grad_(True)
            y2 = fn(x2)
            gx2, gw, gb = torch.autograd.grad((y2 * go).sum(), [x2, conv.weight, act.bias], create_graph=True)
            (ggw,) = torch.autograd.grad(gx2.pow(2).sum(), conv.weight)                    # R1 pattern
            res.append((y.detach(), gx, gx2.detach(), gw.detach(), gb.detach(), ggw))
        assert res[0][0].is_contiguous(memory_format=torch.channels_last)
        for name, a, b in zip(("y", "gx", "gx (graph)", "gw", "gb", "d|gx|^2/dw"), *res):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5, msg=name)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("shape", [(2, 512, 512), (1, 128, 256), (4, 256, 128), (8, 64, 32)], ids=str)
def test_demod_fused_matches_module_code_to_second_order(shape):
    """rick_demod_fwd / rick_demod_bwd against the module code they replace (model_probe_tune.py:246-251 in algebraic form:
    pow / linear / mul / add / rsqrt), first derivatives from the fused backward, second derivatives through the
    differentiable composite branch (what the path-length regulariser differentiates)."""
    from rick_b200.op import glue
    b, cin, cout = shape
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    s0 = torch.randn(b, cin, device="cuda", generator=g) * 0.5 + 1.0
    w0 = torch.rand(cout, cin, device="cuda", generator=g) * 9.0
    scale2, eps, s_scale = 1.0 / (cin * 9), 1e-8, (cin * 9) ** -0.5
    gd = torch.randn(b, cout, device="cuda", generator=g)
    gs = torch.randn(b, cin, device="cuda", generator=g)

    def run(fn, second):
        s, w = s0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        d, so = fn(s, w, scale2, eps, s_scale)
        g_s, g_w = torch.autograd.grad((d, so), (s, w), (gd, gs), create_graph=second)
        if not second:
            return d, so, g_s, g_w
        gg_s, gg_w = torch.autograd.grad((g_s * gs).sum() + (g_w.square()).sum(), (s, w))
        return d, so, g_s, g_w, gg_s, gg_w

    for second in (False, True):
        got = run(glue.demod, second)
        want = run(glue._demod_composite, second)
        for a, e in zip(got, want):
            assert (a - e).abs().max().item() <= 2e-5 * max(e.abs().max().item(), 1e-6)


def test_add_scale_matches_residual_merge():
    """rick_add_scale = ResBlock's (out + skip) / sqrt(2) (model_probe_tune.py:655-660), NCHW and channels-last, with its
    gradient; odd element counts fall back to the torch expression."""
    from rick_b200.op import glue
    g = torch.Generator(device="cuda").manual_seed(3)
    for shape, cl in (((4, 256, 32, 32), True), ((2, 512, 8, 8), False), ((1, 3, 5, 5), False)):
        a = torch.randn(*shape, device="cuda", generator=g)
        b = torch.randn(*shape, device="cuda", generator=g)
        if cl:
            a, b = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
        a.requires_grad_(True), b.requires_grad_(True)
        out = glue.add_scale(a, b, 2 ** -0.5)
        want = (a.detach() + b.detach()) * 2 ** -0.5
        assert out.stride() == want.stride() and torch.equal(out, want)
        go = torch.randn_like(out)
        ga, gb = torch.autograd.grad(out, (a, b), go)
        assert torch.equal(ga, go * 2 ** -0.5) and torch.equal(gb, ga)

"""rick_b200's kernels against the REFERENCE'S OWN CUDA kernels (op/upfirdn2d_kernel.cu:209-369,
op/fused_bias_act_kernel.cu:52-98), compiled unmodified for sm_100a into oracle/_ref/ by oracle/build_ref_cuda.py.

Second oracle ("on the gpurun B200 box it is a second oracle and the kernel-to-beat"): the CPU
oracle is pinned to ``upfirdn2d_native``; this pins the same results to the native kernels the reference actually runs
on a GPU, on the call shapes of G / D at 256 px, the op-sweep shapes of BASELINE configs[3] and the 12x12-tap
augmentation shape.  The last test times both on the op-sweep shapes and writes gpurun_out/ref_ops_vs_ours.json.

Skipped when oracle/_ref/ has not been built (it is built in the authoring container, where /root/reference exists)."""
import json
import math
import os

import pytest
import torch

from oracle import build_ref_cuda

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    up, fu = build_ref_cuda.load("ref_upfirdn2d"), build_ref_cuda.load("ref_fused")
    if up is None or fu is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref_cuda.py needs /root/reference)")
    return up, fu


@pytest.fixture(scope="module")
def op():
    from rick_b200 import op as _op
    return _op


def ref_upfirdn2d(mod, x, taps, up, down, pad):
    """What UpFirDn2d.forward does around the native call (op/upfirdn2d.py:88-119): NCHW -> (N*C, H, W, 1) and back."""
    n, c, h, w = x.shape
    out = mod.upfirdn2d(x.reshape(-1, h, w, 1), taps, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    return out.view(n, c, out.shape[1], out.shape[2])


def _taps(gain=1.0, k=(1., 3., 3., 1.)):
    t = torch.tensor(k)
    t = torch.outer(t, t)
    return (t / t.sum() * gain).cuda()


def _rel(got, want):
    return ((got.double() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-30)).item()


UPFIRDN = [
    # (N, C, H, W, up, down, pad, gain): G blur after transposed conv, RGB upsample, D blurs, Downsample, ragged, crops
    (2, 512, 9, 9, 1, 1, (1, 1), 4), (2, 512, 65, 65, 1, 1, (1, 1), 4), (2, 256, 129, 129, 1, 1, (1, 1), 4),
    (2, 128, 257, 257, 1, 1, (1, 1), 4), (2, 3, 4, 4, 2, 1, (2, 1), 4), (2, 3, 128, 128, 2, 1, (2, 1), 4),
    (4, 128, 256, 256, 1, 1, (2, 2), 1), (4, 128, 256, 256, 1, 1, (1, 1), 1), (4, 512, 8, 8, 1, 1, (2, 2), 1),
    (3, 7, 64, 64, 1, 2, (1, 1), 1), (1, 5, 37, 53, 2, 1, (2, 1), 4), (1, 5, 37, 53, 1, 2, (2, 1), 1),
    (1, 2, 31, 31, 2, 1, (-1, 3), 4), (1, 2, 40, 40, 1, 1, (-1, -2), 1), (8, 64, 32, 32, 2, 1, (2, 1), 4),
]


@pytest.mark.parametrize("case", UPFIRDN, ids=[f"{c[0]}x{c[1]}x{c[2]}x{c[3]}_u{c[4]}d{c[5]}p{c[6][0]}{c[6][1]}" for c in UPFIRDN])
def test_upfirdn2d_vs_reference_cuda(case, ref, op):
    n, c, h, w, up, down, pad, gain = case
    g = torch.Generator(device="cuda").manual_seed(h * 131 + w + up)
    x = torch.randn(n, c, h, w, device="cuda", generator=g)
    taps = _taps(gain)
    want = ref_upfirdn2d(ref[0], x, taps, up, down, pad)
    got = op.upfirdn2d(x, taps, up, down, pad)
    assert got.shape == want.shape
    assert _rel(got, want) <= 1e-5


def test_upfirdn2d_12x12_taps_vs_reference_cuda(ref, op):
    """non_leaking.py:338,359: asymmetric 12-tap separable wavelet filters, up = 2 / down = 2."""
    g = torch.Generator(device="cuda").manual_seed(5)
    k1 = torch.randn(12, device="cuda", generator=g)
    taps = torch.outer(k1, k1.flip(0) * 0.7 + 0.1)
    x = torch.randn(8, 3, 70, 70, device="cuda", generator=g)
    for up, down, pad in ((2, 1, (6, 5)), (1, 2, (5, 5)), (1, 1, (3, 8))):
        want = ref_upfirdn2d(ref[0], x, taps, up, down, pad)
        got = op.upfirdn2d(x, taps, up, down, pad)
        assert got.shape == want.shape
        assert _rel(got, want) <= 1e-5, (up, down, pad)


def test_upfirdn2d_backward_vs_reference_cuda(ref, op):
    """The reference's backward is the same native op with up <-> down swapped, flipped taps and g_pad
    (op/upfirdn2d.py:19-58, 105-117); ours is the self-adjoint Function.  Compare the gradients they produce."""
    g = torch.Generator(device="cuda").manual_seed(9)
    for (n, c, h, w, up, down, pad) in ((2, 16, 33, 33, 1, 1, (1, 1)), (2, 3, 32, 32, 2, 1, (2, 1)), (2, 8, 64, 64, 1, 2, (1, 1))):
        x = torch.randn(n, c, h, w, device="cuda", generator=g, requires_grad=True)
        taps = _taps(up * up)
        out = op.upfirdn2d(x, taps, up, down, pad)
        go = torch.randn(out.shape, device="cuda", generator=g)
        (gx,) = torch.autograd.grad(out, x, go)
        kh = kw = 4
        out_h, out_w = out.shape[2:]
        gpx0, gpy0 = kw - pad[0] - 1, kh - pad[0] - 1
        gpx1 = w * up - out_w * down + pad[0] - up + 1
        gpy1 = h * up - out_h * down + pad[0] - up + 1
        want = ref[0].upfirdn2d(go.reshape(-1, out_h, out_w, 1), torch.flip(taps, [0, 1]), down, down, up, up,
                                gpx0, gpx1, gpy0, gpy1).view(n, c, h, w)
        assert _rel(gx, want) <= 1e-5, (n, c, h, w, up, down)


@pytest.mark.parametrize("shape", [(2, 512), (2, 512, 4, 4), (4, 128, 256, 256), (3, 7, 33, 65)], ids=str)
def test_fused_leaky_relu_vs_reference_cuda(shape, ref, op):
    """Forward (act=3, grad=0), the backward's grad_input (act=3, grad=1, ref=out) + grad_bias reduction
    (op/fused_act.py:19-70), and the double-backward mode (act=3, grad=1 with gradgrad_bias as bias, 44-46)."""
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn(*shape, device="cuda", generator=g, requires_grad=True)
    b = torch.randn(shape[1], device="cuda", generator=g, requires_grad=True)
    empty = x.new_empty(0)
    want = ref[1].fused_bias_act(x.detach(), b.detach(), empty, 3, 0, 0.2, math.sqrt(2))
    got = op.fused_leaky_relu(x, b)
    assert _rel(got, want) <= 1e-5
    go = torch.randn(*shape, device="cuda", generator=g, requires_grad=True)
    gx, gb = torch.autograd.grad(got, (x, b), go, create_graph=True)
    want_gx = ref[1].fused_bias_act(go.detach(), empty, want, 3, 1, 0.2, math.sqrt(2))
    dims = [0] + list(range(2, len(shape)))
    assert _rel(gx, want_gx) <= 1e-5
    assert _rel(gb, want_gx.sum(dims)) <= 1e-4            # different (deterministic) reduction order
    ggx = torch.randn(*shape, device="cuda", generator=g)
    ggb = torch.randn(shape[1], device="cuda", generator=g)
    (gg,) = torch.autograd.grad((gx * ggx).sum() + (gb * ggb).sum(), go)
    want_gg = ref[1].fused_bias_act(ggx, ggb, want, 3, 1, 0.2, math.sqrt(2))
    assert _rel(gg, want_gg) <= 1e-5


def _time(fn, flush, iters=5):
    """median device time (ms) of fn(): CUDA events on the launching stream, L2 flushed before every launch; the call is
    replayed from a CUDA graph so that neither side's Python dispatch is inside the events."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph):
            keep = fn()                                     # noqa: F841
        graph.replay()
    except Exception:
        graph = None
        torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        graph.replay() if graph is not None else fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def test_kernel_to_beat_op_sweep(ref, op):
    """BASELINE configs[3]: the reference's native kernels against this package's on the op-sweep shapes (batch 32,
    512 channels, fp32), CUDA events on the launching stream, L2 flushed before every launch.  Asserts only that ours
    is not slower on the large maps; the numbers go to gpurun_out/ref_ops_vs_ours.json (copied to profiles/)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    taps4, taps1 = _taps(4), _taps(1)
    for r in (4, 8, 16, 32, 64, 128):
        x = torch.randn(32, 512, r, r, device="cuda")
        for name, xin, tp, up, down, pad in (
                (f"upfirdn2d_up2_{r}->{2 * r}", x, taps4, 2, 1, (2, 1)),
                (f"blur_{2 * r + 1}->{2 * r}", torch.randn(32, 512, 2 * r + 1, 2 * r + 1, device="cuda") if r <= 64 else None,
                 taps4, 1, 1, (1, 1)),
                (f"d_blur_pad22_{r}->{r + 1}", x, taps1, 1, 1, (2, 2)),
                (f"d_blur_pad11_{r}->{r - 1}", x, taps1, 1, 1, (1, 1))):
            if xin is None:
                continue
            t_ref = _time(lambda: ref_upfirdn2d(ref[0], xin, tp, up, down, pad), flush)
            t_our = _time(lambda: op.upfirdn2d(xin, tp, up, down, pad), flush)
            out = op.upfirdn2d(xin, tp, up, down, pad)
            gb = 4 * (xin.numel() + out.numel()) / 1e9
            rows.append({"op": name, "algorithmic_GB": round(gb, 4), "reference_ms": t_ref, "ours_ms": t_our,
                         "reference_GBps": gb / t_ref * 1e3, "ours_GBps": gb / t_our * 1e3, "speedup": t_ref / t_our})
            del out
    empty = torch.empty(0, device="cuda")
    for R in (4, 8, 16, 32, 64, 128, 256):
        b = 32 if R < 256 else 16                              # 32 x 512 x 256^2 fp32 is 4.3 GB per tensor: keep 3 of them modest
        x = torch.randn(b, 512, R, R, device="cuda")
        bias = torch.randn(512, device="cuda")
        t_ref = _time(lambda: ref[1].fused_bias_act(x, bias, empty, 3, 0, 0.2, math.sqrt(2)), flush)
        t_our = _time(lambda: op.fused_leaky_relu(x, bias), flush)
        gb = 4 * 2 * x.numel() / 1e9
        rows.append({"op": f"bias_act_fwd_{R}_b{b}", "algorithmic_GB": round(gb, 4), "reference_ms": t_ref, "ours_ms": t_our,
                     "reference_GBps": gb / t_ref * 1e3, "ours_GBps": gb / t_our * 1e3, "speedup": t_ref / t_our})
        out = op.fused_leaky_relu(x, bias)
        go = torch.randn_like(x)

        def ref_bwd():                                         # FusedLeakyReLUFunctionBackward.forward, op/fused_act.py:21-39
            gi = ref[1].fused_bias_act(go, empty, out, 3, 1, 0.2, math.sqrt(2))
            return gi, gi.sum([0, 2, 3])

        from rick_b200.op import fused_act
        t_ref = _time(ref_bwd, flush)
        t_our = _time(lambda: fused_act.FusedLeakyReLUFunctionBackward.apply(go, out, 0.2, math.sqrt(2)), flush)
        gb = 4 * 3 * x.numel() / 1e9
        rows.append({"op": f"bias_act_bwd_{R}_b{b}", "algorithmic_GB": round(gb, 4), "reference_ms": t_ref, "ours_ms": t_our,
                     "reference_GBps": gb / t_ref * 1e3, "ours_GBps": gb / t_our * 1e3, "speedup": t_ref / t_our})
        del x, out, go
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_ops_vs_ours.json"), "w") as f:
        json.dump(rows, f, indent=1)
    for r in rows:
        print(f"{r['op']:28s} ref {r['reference_ms']:8.3f} ms {r['reference_GBps']:7.0f} GB/s | ours {r['ours_ms']:8.3f} ms "
              f"{r['ours_GBps']:7.0f} GB/s | x{r['speedup']:.2f}")
    big = [r for r in rows if r["algorithmic_GB"] >= 0.5]
    assert all(r["speedup"] >= 1.0 for r in big), [r for r in big if r["speedup"] < 1.0]

"""Host-side logic that needs no GPU: module wiring (with the two CUDA ops stubbed by the oracle -- a test of the
Python layer, not of the kernels), draw-stream determinism, layer tables, sharding, and the world_size-2 gloo path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import model_oracle as mo
from oracle import ops_oracle as ops
from oracle import synth


@pytest.fixture()
def cpu_stubbed(monkeypatch):
    """Route the two custom ops to the oracle and lift the CUDA guard, so the nn.Module wiring runs on CPU."""
    import rick_b200.conv as cv
    import rick_b200.stylegan2 as sg
    monkeypatch.setattr(sg, "upfirdn2d", ops.upfirdn2d)
    monkeypatch.setattr(sg, "fused_leaky_relu", ops.fused_leaky_relu)
    monkeypatch.setattr(sg.FusedLeakyReLU, "forward",
                        lambda self, x: ops.fused_leaky_relu(x, self.bias, self.negative_slope, self.scale))
    monkeypatch.setattr(cv, "_require_cuda", lambda *a: None)
    return sg


def test_module_wiring_matches_oracle_on_cpu(cpu_stubbed):
    sg = cpu_stubbed
    size = 32
    gp, dp = synth.g_state(size, 11), synth.d_state(size, 12)
    G, D = sg.Generator(size, 512, 8), sg.Discriminator(size)
    G.load_state_dict(gp)
    D.load_state_dict(dp)
    z, z2 = synth.latents(2, 21), synth.latents(2, 22)
    with torch.no_grad():
        a, _ = G([z], randomize_noise=False)
        b, _ = mo.g_forward(gp, [z], size, randomize_noise=False)
        assert (a - b).abs().max() < 5e-5 * b.abs().max()         # algebraic modulated conv: fp32 rounding only
        a2, lat = G([z, z2], inject_index=3, randomize_noise=False, return_latents=True)
        b2, lat2 = mo.g_forward(gp, [z, z2], size, randomize_noise=False, inject_index=3, return_latents=True)
        assert torch.equal(lat, lat2) and (a2 - b2).abs().max() < 5e-5 * b2.abs().max()
        la, feat = D(b)
        ld = mo.d_forward(dp, b, size)
        assert (la - ld).abs().max() < 1e-5 * max(1.0, ld.abs().max().item())   # from-RGB layer sums in another order
        assert len(feat) == 8
    names = [n for n, _ in G.named_parameters()]
    assert names == mo.g_param_names(size)
    assert [n for n, _ in D.named_parameters()] == mo.d_param_names(size)


def test_cuda_only_surface_raises_on_cpu():
    from rick_b200 import op
    with pytest.raises(RuntimeError, match="CUDA"):
        op.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        op.fused_leaky_relu(torch.zeros(1, 2, 4, 4), torch.zeros(2))
    from rick_b200 import conv
    with pytest.raises(RuntimeError, match="CUDA"):
        conv.modulated_conv2d(torch.zeros(1, 2, 4, 4), torch.zeros(2, 2, 3, 3), torch.ones(1, 2), None)


def test_draw_stream_is_reproducible_and_ordered():
    from rick_b200.adapt import DrawStream
    a, b = DrawStream(5, "cpu"), DrawStream(5, "cpu")
    za, zb = a.mixing_latents(2, 512, 0.9), b.mixing_latents(2, 512, 0.9)
    assert len(za) == len(zb) and all(torch.equal(x, y) for x, y in zip(za, zb))
    assert a.randint(1, 12) == b.randint(1, 12)
    na, nb = a.layer_noise(2, 256), b.layer_noise(2, 256)
    assert [t.shape[-1] for t in na] == [4, 8, 8, 16, 16, 32, 32, 64, 64, 128, 128, 256, 256]
    assert all(torch.equal(x, y) for x, y in zip(na, nb))
    assert len(DrawStream(1, "cpu").mixing_latents(2, 8, 0.0)) == 1


def test_layer_tables_cover_reference_keys():
    from rick_b200 import rick
    from oracle import rick_oracle as ro
    g = {k: torch.from_numpy(v) for k, v in synth.fisher_g(1).items()}
    d = {k: torch.from_numpy(v) for k, v in synth.fisher_d(2).items()}
    gl, dl = rick.generator_layers(g), rick.discriminator_layers(d)
    assert [l.weight for l in gl if l.group == "conv"] == ro.g_conv_keys()
    assert [(l.weight, l.bias) for l in gl if l.group == "fc"] == ro.g_fc_keys()
    assert [(l.weight, l.bias) for l in dl] == ro.d_layer_keys()
    assert sum(l.rows for l in gl if l.group == "conv") == 4864
    assert sum(l.rows for l in gl if l.group == "fc") == 5248
    assert sum(l.rows for l in dl) == 8064
    assert all(l.closed_low == (l.bias is None) for l in dl)


def test_sharding_helpers():
    from rick_b200 import dist as rd
    n_batches = (5000 + 63) // 64
    for world in (1, 2, 4, 8):
        seen = sorted(sum((rd.shard_batches(n_batches, r, world) for r in range(world)), []))
        assert seen == list(range(n_batches))
        imgs = sorted(sum((list(rd.shard_range(5, r, world)) for r in range(world)), []))
        assert imgs == list(range(5))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from rick_b200 import dist as rd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    rd.init_from_env("gloo")
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(40, 6, generator=g, dtype=torch.float64)
    # sample-generation statistics: each rank sees its shard of batches, one all-reduce at the end
    st = rd.FeatureStats(6, "cpu")
    for k in rd.shard_batches(10, rank, world):
        st.update(feats[4 * k:4 * k + 4])
    mu, cov = st.all_reduce().mean_cov()
    # adaptation: gradient averaging; Fisher: accumulator sum
    grads = [torch.full((3, 5), float(rank + 1)), torch.full((7,), float(10 * (rank + 1)))]
    rd.allreduce_mean_(grads, bucket_bytes=32)
    acc = [torch.full((4,), float(rank + 1))]
    rd.allreduce_sum_(acc)
    # adaptation, overlapped form: GradSync launches bucket all-reduces from gradient hooks during backward
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                              torch.nn.Linear(16, 4))
    params = list(net.parameters())
    sync = rd.GradSync(params, n_buckets=3, late=[params[1]])
    assert sum(len(b) for b in sync.buckets) == len(params) and sync.buckets[-1] == [params[1]]
    synced = []
    for step in range(2):                                   # second pass: counters re-armed, one gradient absent
        xin = torch.randn(5, 6, generator=torch.Generator().manual_seed(100 * step + rank))
        net.zero_grad(set_to_none=True)
        inputs = params if step == 0 else params[:-1]
        lg = torch.autograd.grad(net(xin).pow(2).sum(), inputs)       # this rank's own gradients (no hooks fire)
        local = [lg[k].clone() if k < len(inputs) else None for k in range(len(params))]
        sync.begin()
        torch.autograd.backward(net(xin).pow(2).sum(), inputs=inputs)
        sync.finish()
        synced.append(([None if p.grad is None else p.grad.clone() for p in params], local))
    torch.save({"mu": mu, "cov": cov, "grads": grads, "acc": acc, "synced": synced}, os.path.join(out_dir, f"r{rank}.pt"))
    rd.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_exchange_steps(tmp_path):
    world, port = 2, _free_port()
    mp.start_processes(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(40, 6, generator=g, dtype=torch.float64).numpy()
    for r in range(world):
        o = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        np.testing.assert_allclose(o["mu"].numpy(), feats.mean(0), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(o["cov"].numpy(), np.cov(feats, rowvar=False), rtol=1e-10, atol=1e-12)
        assert torch.equal(o["grads"][0], torch.full((3, 5), 1.5)) and torch.equal(o["grads"][1], torch.full((7,), 15.0))
        assert torch.equal(o["acc"][0], torch.full((4,), 3.0))
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for step in range(2):
        locals_ = [o["synced"][step][1] for o in outs]
        for r in range(world):
            got = outs[r]["synced"][step][0]
            for k, g in enumerate(got):
                if g is None:
                    assert step == 1 and all(l[k] is None for l in locals_)
                    continue
                want = sum(l[k] for l in locals_) / world
                torch.testing.assert_close(g, want, rtol=1e-6, atol=1e-7)


def _emulate_conv_geom(x_nhwc, wt, g):
    """Pure-torch reading of a rick_conv_geom (include/rick_b200.h): what rick_conv_tc must compute."""
    b, h, w, cin = x_nhwc.shape
    out = torch.zeros(b, g.out_h, g.out_w, g.cout, dtype=torch.float64)
    xp = x_nhwc.double()
    for pi in range(g.n_phases):
        ph = g.phase[pi]
        for m in range(ph.rows):
            for n in range(ph.cols):
                acc = torch.zeros(b, g.cout, dtype=torch.float64)
                for t in range(ph.n_taps):
                    iy, ix = m * g.in_stride + ph.dy[t], n * g.in_stride + ph.dx[t]
                    if 0 <= iy < h and 0 <= ix < w:
                        acc += xp[:, iy, ix, :] @ wt[ph.widx[t]].double().T
                out[:, m * g.out_stride + ph.out_y0, n * g.out_stride + ph.out_x0, :] = acc
    return out


def test_conv_geometry_tables_describe_the_right_convolutions():
    from torch.nn import functional as F
    from rick_b200 import conv_tc as ct
    g = torch.Generator().manual_seed(0)
    b, cin, cout = 2, 3, 4
    for (h, w, k, stride, pad) in [(5, 6, 3, 1, 1), (7, 7, 3, 2, 0), (6, 4, 1, 2, 0), (4, 4, 1, 1, 0)]:
        x = torch.randn(b, cin, h, w, generator=g)
        wgt = torch.randn(cout, cin, k, k, generator=g)
        geom = ct.geom_conv(b, h, w, cin, cout, k, stride, pad)
        got = _emulate_conv_geom(x.permute(0, 2, 3, 1), ct.pack_weight(wgt), geom).permute(0, 3, 1, 2)
        want = F.conv2d(x.double(), wgt.double(), stride=stride, padding=pad)
        assert got.shape == want.shape and torch.allclose(got, want, atol=1e-10), (h, w, k, stride, pad)
    for (h, w) in [(4, 4), (3, 5)]:
        x = torch.randn(b, cin, h, w, generator=g)
        wgt = torch.randn(cout, cin, 3, 3, generator=g)
        geom = ct.geom_conv_transpose_s2(b, h, w, cin, cout)
        got = _emulate_conv_geom(x.permute(0, 2, 3, 1), ct.pack_weight(wgt), geom).permute(0, 3, 1, 2)
        want = F.conv_transpose2d(x.double(), wgt.double().transpose(0, 1), stride=2, padding=0)
        assert got.shape == want.shape == (b, cout, 2 * h + 1, 2 * w + 1) and torch.allclose(got, want, atol=1e-10)


def _epilogue_cpu(acc, demod, noise, noise_weight, bias, act, alpha, scale, s_next, want_out2):
    """include/rick_b200.h rick_conv_epilogue semantics on NHWC tensors."""
    v = acc
    if demod is not None:
        v = v * demod[:, None, None, :]
    if noise is not None:
        v = v + noise_weight * noise[..., None]
    if bias is not None:
        v = v + bias
    if act:
        v = torch.where(v > 0, v, v * alpha) * scale
    vm = v if s_next is None else v * s_next[:, None, None, :]
    return (v, vm) if want_out2 else vm


def test_fused_generator_plan_on_cpu(cpu_stubbed, monkeypatch):
    """The execution plan of FusedGenerator (style indices, epilogue wiring, s_next chaining, skip upsampling) with its
    four kernels replaced by torch emulations of their documented semantics -- host logic only, no GPU."""
    from torch.nn import functional as F
    import rick_b200.fused as fused
    from rick_b200 import conv_tc as ct

    def conv_stub(xm, wt, geom, demod=None, noise=None, noise_weight=None, bias=None, act=False, alpha=0.2,
                  scale=2 ** 0.5, s_next=None, want_out2=False):
        k = int(round(geom.n_weight_taps ** 0.5))
        # the weight arrives either packed (taps, Cout, Cin) or as the (Cout, Cin, k, k) parameter read in place
        w = wt if wt.dim() == 4 else wt.view(k, k, geom.cout, geom.cin).permute(2, 3, 0, 1)
        x = xm.permute(0, 3, 1, 2)
        if geom.n_phases == 4:
            acc = F.conv_transpose2d(x, w.transpose(0, 1), stride=2)
        else:
            acc = F.conv2d(x, w, stride=geom.in_stride, padding=-geom.phase[0].dy[0])
        return _epilogue_cpu(acc.permute(0, 2, 3, 1), demod, noise, noise_weight, bias, act, alpha, scale, s_next, want_out2)

    def blur_stub(x, taps, pad, demod=None, noise=None, noise_weight=None, bias=None, act=False, alpha=0.2,
                  scale=2 ** 0.5, s_next=None, want_out2=False):
        acc = ops.upfirdn2d(x.permute(0, 3, 1, 2), taps, pad=pad).permute(0, 2, 3, 1)
        return _epilogue_cpu(acc, demod, noise, noise_weight, bias, act, alpha, scale, s_next, want_out2)

    def rgb_stub(y, wmod, bias, skip):
        out = torch.einsum("bhwc,boc->bohw", y, wmod) + bias[None, :, None, None]
        return out if skip is None else out + skip

    monkeypatch.setattr(ct, "conv_tc_nhwc", conv_stub)
    monkeypatch.setattr(ct, "blur_nhwc", blur_stub)
    monkeypatch.setattr(ct, "to_rgb_nhwc", rgb_stub)
    monkeypatch.setattr(fused, "upfirdn2d", ops.upfirdn2d)
    sg = cpu_stubbed
    size = 32
    gp = synth.g_state(size, 11)
    G = sg.Generator(size, 512, 8)
    G.load_state_dict(gp)
    FG = fused.FusedGenerator(G)
    z, z2 = synth.latents(2, 21), synth.latents(2, 22)
    want, _ = mo.g_forward(gp, [z], size, randomize_noise=False)
    got, none = FG([z], randomize_noise=False)
    assert none is None and (got - want).abs().max() < 1e-4 * want.abs().max()
    want2, lat2 = mo.g_forward(gp, [z, z2], size, randomize_noise=False, inject_index=3, return_latents=True)
    got2, lat = FG([z, z2], randomize_noise=False, inject_index=3, return_latents=True)
    assert torch.allclose(lat, lat2) and (got2 - want2).abs().max() < 1e-4 * want2.abs().max()
    explicit = [torch.randn(2, 1, 2 ** r, 2 ** r) for r in [2, 3, 3, 4, 4, 5, 5]]
    want3, _ = mo.g_forward(gp, [z], size, noise=explicit)
    got3, _ = FG([z], noise=explicit)
    assert (got3 - want3).abs().max() < 1e-4 * want3.abs().max()


# ------------------------------------------------------------------------------------------ checkpoints
def test_checkpoint_round_trip_in_reference_format(tmp_path):
    """rick_b200.checkpoint writes the reference's five keys (train:645-659); the file loads back, and -- when the
    reference tree is present -- straight into the reference's own Generator / Discriminator with strict=True."""
    import types
    from rick_b200 import checkpoint, stylegan2 as sg
    torch.manual_seed(0)
    size = 32

    def make():                                              # the attributes of RickAdapter that checkpoints touch
        g, d, ge, de = sg.Generator(size, 512, 8), sg.Discriminator(size), sg.Generator(size, 512, 8), sg.Discriminator(size)
        g_train = [p for n, p in g.named_parameters() if "convs" in n]
        d_train = [p for n, p in d.named_parameters() if ("convs" in n and "convs.0" not in n) or "final" in n]
        return types.SimpleNamespace(g=g, d=d, g_ema=ge, d_ema=de, g_train=g_train, d_train=d_train,
                                     g_optim=torch.optim.Adam(g_train, lr=0.0016, betas=(0.0, 0.99 ** 0.8)),
                                     d_optim=torch.optim.Adam(d_train, lr=0.0019, betas=(0.0, 0.99 ** (16 / 17))))
    A = make()
    nets = [A.g, A.d]
    for p in A.g_train[:3] + A.d_train[:3]:                  # give the optimisers some state
        p.grad = torch.randn_like(p)
    A.g_optim.step(), A.d_optim.step()
    path = tmp_path / "000100.pt"
    checkpoint.save(path, A)
    ckpt = torch.load(path, weights_only=False)
    assert sorted(ckpt) == ["d", "d_optim", "g", "g_ema", "g_optim"]
    assert all(v.is_contiguous() for v in ckpt["d"].values() if torch.is_tensor(v))
    assert set(ckpt["g"]) == set(nets[0].state_dict()) and set(ckpt["d"]) == set(nets[1].state_dict())

    B = make()
    fresh = [B.g, B.d]
    checkpoint.resume(path, B)
    for k, v in A.g.state_dict().items():
        assert torch.equal(B.g.state_dict()[k], v), k
    for k, v in A.d.state_dict().items():
        assert torch.equal(B.d_ema.state_dict()[k], v), k    # d_ema <- ckpt["d"] (train:879)
    sa, sb = A.g_optim.state_dict(), B.g_optim.state_dict()
    assert sa["param_groups"] == sb["param_groups"] and sorted(sa["state"]) == sorted(sb["state"])
    for i in sa["state"]:
        assert torch.equal(sa["state"][i]["exp_avg_sq"], sb["state"][i]["exp_avg_sq"])

    from oracle import ref_loader
    if ref_loader.available():                               # not on the GPU box
        ref = ref_loader.load().model
        rg, rd = ref.Generator(size, 512, 8), ref.Discriminator(size)
        rg.load_state_dict(ckpt["g_ema"], strict=True)
        rd.load_state_dict(ckpt["d"], strict=True)
        # and the other direction: a reference-written state dict loads into this package's modules
        fresh[0].load_state_dict(rg.state_dict(), strict=True)
        fresh[1].load_state_dict(rd.state_dict(), strict=True)


# ------------------------------------------------------------------------------------------ conv gradient structure
@pytest.mark.parametrize("transposed,stride,padding", [(False, 1, 1), (False, 2, 0), (True, 2, 0)])
def test_lib_conv_first_and_second_order_match_autograd(transposed, stride, padding):
    """rick_b200.conv._lib_conv spells out dgrad / wgrad and their derivatives (so R1 / path-length never hit autograd's
    feature-map-sized-filter double-backward); every order must equal what autograd derives for F.conv2d itself."""
    from torch.nn import functional as F
    from rick_b200.conv import _lib_conv
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 9, 9, dtype=torch.double, generator=g).requires_grad_(True)
    w = torch.randn(6, 4, 3, 3, dtype=torch.double, generator=g).requires_grad_(True)

    def ref(x, w):
        if transposed:
            return F.conv_transpose2d(x, w.transpose(0, 1), stride=stride, padding=padding)
        return F.conv2d(x, w, stride=stride, padding=padding)

    def new(x, w):
        return _lib_conv(x, w, stride, padding, transposed)     # w is (Cout, Cin, k, k) in both directions

    res = []
    for f in (ref, new):
        y = f(x, w)
        gx, gw = torch.autograd.grad((y ** 2).sum(), [x, w], create_graph=True)
        ramp = torch.arange(gw.numel(), dtype=torch.double).view_as(gw)
        ggx, ggw = torch.autograd.grad((gx ** 3).sum() + (gw ** 2 * ramp).sum(), [x, w])
        (gx_only,) = torch.autograd.grad((f(x, w) ** 2).sum(), [x], create_graph=True)     # the path-length pattern
        (gw_from_gx,) = torch.autograd.grad((gx_only ** 2).sum(), [w])
        res.append((y.detach(), gx.detach(), gw.detach(), ggx, ggw, gw_from_gx))
    for a, b in zip(*res):
        torch.testing.assert_close(a, b, rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------------------------------------ joint D pass
@pytest.mark.parametrize("batch,cap", [(2, 25), (6, 4), (8, 4), (50, 25), (3, 4)])
def test_d_pair_grouping_matches_two_calls(batch, cap):
    """adapt.d_pair stacks fake and real so that every minibatch-stddev group of the joint call is a group of ONE of the
    two batches -- checked on a stand-in discriminator that has nothing but that statistic (model_probe_tune.py:749-756)."""
    from rick_b200.adapt import d_pair

    class StddevOnly(torch.nn.Module):
        stddev_group = cap

        def forward(self, x, stddev_group=None):
            b = x.shape[0]
            group = min(b, self.stddev_group) if stddev_group is None else stddev_group
            sd = x.view(group, -1, *x.shape[1:])
            sd = torch.sqrt(sd.var(0, unbiased=False) + 1e-8).mean(dim=(1, 2, 3), keepdim=True)   # one value per group
            sd = sd.repeat(group, 1, 1, 1)
            return (x.mean(dim=(1, 2, 3)) + 10 * sd.view(b)).unsqueeze(1), None

    g = torch.Generator().manual_seed(batch)
    fake, real = torch.randn(batch, 3, 4, 4, generator=g), torch.randn(batch, 3, 4, 4, generator=g) * 2 + 1
    d = StddevOnly()
    if batch % min(batch, cap) != 0:                          # the reference's own view() would fail: falls back
        with pytest.raises(RuntimeError):
            d(fake)
        return
    want_f, want_r = d(fake)[0], d(real)[0]
    got_f, got_r = d_pair(d, fake, real)
    torch.testing.assert_close(got_f, want_f)
    torch.testing.assert_close(got_r, want_r)


# ---------------------------------------------------------------------------------------------------------------
# geometry builders of the tcgen05 convolution (rick_b200.conv_tc): interpreted on the CPU exactly as the kernels read
# them (include/rick_b200.h) and compared with torch's convolution / its gradients
# ---------------------------------------------------------------------------------------------------------------
def _interp_conv(x, w, geom, transpose_weight):
    """rick_conv_tc_w semantics: x (B,H,W,K) NHWC, w (Cout,Cin,k,k); returns (B,OH,OW,M)."""
    import torch
    B = geom.batch
    out = torch.zeros(B, geom.out_h, geom.out_w, geom.cout, dtype=torch.float64)
    k = w.shape[-1]
    for i in range(geom.n_phases):
        ph = geom.phase[i]
        for t in range(ph.n_taps):
            ky, kx = divmod(ph.widx[t], k)
            wm = w[:, :, ky, kx].double()                     # (Cout, Cin)
            op = wm.t() if transpose_weight else wm           # (M, K)
            for m in range(ph.rows):
                iy = m * geom.in_stride + ph.dy[t]
                if not 0 <= iy < geom.in_h:
                    continue
                for n in range(ph.cols):
                    ix = n * geom.in_stride + ph.dx[t]
                    if not 0 <= ix < geom.in_w:
                        continue
                    out[:, m * geom.out_stride + ph.out_y0, n * geom.out_stride + ph.out_x0] += x[:, iy, ix].double() @ op.t()
    return out


def _interp_wgrad(g, x, geom, k):
    import torch
    dw = torch.zeros(geom.cout, geom.cin, k, k, dtype=torch.float64)
    for t in range(geom.n_taps):
        ky, kx = divmod(t, k)
        for m in range(geom.rows):
            gy, xy = m * geom.g_stride + geom.gy[t], m * geom.x_stride + geom.xy[t]
            if not (0 <= gy < geom.g_h and 0 <= xy < geom.x_h):
                continue
            for n in range(geom.cols):
                gx, xx = n * geom.g_stride + geom.gx[t], n * geom.x_stride + geom.xx[t]
                if not (0 <= gx < geom.g_w and 0 <= xx < geom.x_w):
                    continue
                dw[:, :, ky, kx] += g[:, gy, gx].double().t() @ x[:, xy, xx].double()
    return dw


@pytest.mark.parametrize("k,stride,pad,h,w", [(3, 1, 1, 6, 5), (1, 1, 0, 4, 4), (3, 2, 0, 9, 9), (3, 2, 0, 7, 11),
                                              (1, 2, 0, 7, 7), (3, 2, 1, 8, 8), (3, 2, 0, 8, 10)])
def test_conv_geometries_match_torch_conv_and_gradients(k, stride, pad, h, w):
    import torch
    from torch.nn import functional as F
    from rick_b200 import conv_tc as ct
    g = torch.Generator().manual_seed(k * 100 + stride * 10 + h)
    B, cin, cout = 2, 3, 4
    x = torch.randn(B, cin, h, w, generator=g, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, wt, stride=stride, padding=pad)
    go = torch.randn(y.shape, generator=g, dtype=torch.float64)
    gx, gw = torch.autograd.grad(y, [x, wt], go)
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous()
    # forward
    got = _interp_conv(nhwc(x), wt.detach(), ct.geom_conv(B, h, w, cin, cout, k, stride, pad), False)
    torch.testing.assert_close(got, nhwc(y))
    # data gradient: the transposed problem over the same weight memory
    gd = ct.geom_conv_dgrad(B, h, w, cin, cout, k, stride, pad)
    assert (gd.in_h, gd.in_w, gd.out_h, gd.out_w, gd.cin, gd.cout) == (y.shape[2], y.shape[3], h, w, cout, cin)
    got = _interp_conv(nhwc(go), wt.detach(), gd, True)
    torch.testing.assert_close(got, nhwc(gx))
    if k == 1 and stride == 2:
        assert ct._covers_partially(gd)              # odd pixels receive no gradient and no phase writes them: pre-zeroed
    # weight gradient
    got = _interp_wgrad(nhwc(go), nhwc(x), ct.geom_wgrad(B, h, w, cin, cout, k, stride, pad), k)
    torch.testing.assert_close(got, gw)


def test_transposed_conv_geometries_match_torch():
    import torch
    from torch.nn import functional as F
    from rick_b200 import conv_tc as ct
    g = torch.Generator().manual_seed(3)
    B, cin, cout, h, w, k = 2, 3, 4, 5, 6, 3
    x = torch.randn(B, cin, h, w, generator=g, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64, requires_grad=True)     # ModulatedConv2d's (Cout, Cin)
    y = F.conv_transpose2d(x, wt.transpose(0, 1), stride=2)                                      # model_probe_tune.py:265
    go = torch.randn(y.shape, generator=g, dtype=torch.float64)
    gx, gw = torch.autograd.grad(y, [x, wt], go)
    nhwc = lambda t: t.detach().permute(0, 2, 3, 1).contiguous()
    got = _interp_conv(nhwc(x), wt.detach(), ct.geom_conv_transpose_s2(B, h, w, cin, cout), False)
    torch.testing.assert_close(got, nhwc(y))
    # its data gradient is an ordinary stride-2 convolution of g with the weight read transposed (M = Cin, K = Cout)
    gd = ct.geom_conv(B, 2 * h + 1, 2 * w + 1, cout, cin, k, 2, 0)
    got = _interp_conv(nhwc(go), wt.detach(), gd, True)
    torch.testing.assert_close(got, nhwc(gx))
    got = _interp_wgrad(nhwc(go), nhwc(x), ct.geom_wgrad(B, h, w, cin, cout, k, 2, 0, transposed=True), k)
    torch.testing.assert_close(got, gw)


def test_conv_tile_planner_host_logic():
    """rick_conv_tc_plan (host only, runs without a GPU): tiles hold <= 256 pixels and cover the map; single-phase launches
    are planned by waves over the 148 SMs (512 tiles of 256 pixels would leave the 4th round half empty: 592 tiles of 224);
    maps with fewer tiles than SMs get a split-K plan; several samples share a tile on tiny maps."""
    from rick_b200 import conv_tc as ct

    def plan(b, h, cin, cout, k=3, stride=1, pad=1):
        g = ct.geom_conv(b, h, h, cin, cout, k, stride, pad)
        p = ct.launch_plan(g)
        for (tw, th), ph in zip(p["tiles"], [g.phase[i] for i in range(g.n_phases)]):
            assert 1 <= tw <= ph.cols and 1 <= th <= ph.rows and tw * th * p["samples_per_tile"] <= 256
        return p

    p = plan(2, 256, 128, 128)                      # G's 256 px layer at batch 2
    assert p["ksplit"] == 1 and p["samples_per_tile"] == 1
    tw, th = p["tiles"][0]
    tiles = -(-256 // tw) * -(-256 // th) * 2
    assert tiles == p["total_tiles"] and tiles % 148 == 0 and tw * th < 256, p      # whole waves of smaller tiles
    p4 = plan(4, 256, 128, 128)                     # 1024 tiles of 256 pixels: 6.9 rounds, nothing to gain
    assert p4["tiles"][0][0] * p4["tiles"][0][1] == 256 and p4["total_tiles"] == 1024
    small = plan(2, 8, 512, 512)                    # 2 x 64 pixels: one pixel tile, 4 cout tiles -> split K
    assert small["samples_per_tile"] == 2 and small["total_tiles"] == 4 and small["ksplit"] > 1
    big = plan(64, 64, 512, 512)                    # sample generation: plenty of tiles, no split
    assert big["ksplit"] == 1 and big["total_tiles"] == 64 * 16 * 4
    up = ct.launch_plan(ct.geom_conv_transpose_s2(2, 64, 64, 512, 256, 3))
    assert len(up["tiles"]) == 4 and up["ksplit"] == 1

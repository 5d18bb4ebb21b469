"""Pin the CPU oracle to the golden vectors that oracle/make_golden.py produced by running the REAL reference
(no GPU needed; inputs are regenerated from the seeds in oracle/synth.py)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo
from oracle import ops_oracle as ops
from oracle import rick_oracle as ro
from oracle import synth
from conftest import ROOT


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("case", synth.UPFIRDN_CASES, ids=[c[0] for c in synth.UPFIRDN_CASES])
def test_upfirdn2d_oracle_matches_reference_golden(case, golden, oracle_c):
    name, n, c, h, w, kh, kw, up, down, p0, p1, kind = case
    x, taps = synth.upfirdn_case_inputs(case)
    want = golden("ops_golden.npz")[name]
    got = ops.upfirdn2d(x, taps, up, down, (p0, p1)).numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)                      # same torch ops in the same order: bit-identical
    # independent direct-form C statement (double accumulation): agrees to float32 rounding of a <=144-tap sum
    xc = np.ascontiguousarray(x.numpy().reshape(n * c, h, w))
    tc = np.ascontiguousarray(taps.numpy())
    out = np.empty((n * c,) + want.shape[2:], np.float32)
    st = oracle_c.oracle_upfirdn2d(_vp(xc), _vp(tc), _vp(out), ctypes.c_long(n * c), ctypes.c_long(h), ctypes.c_long(w),
                                   kh, kw, up, up, down, down, p0, p1, p0, p1)
    assert st == 0
    np.testing.assert_allclose(out.reshape(want.shape), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


def test_bias_act_oracle_modes(oracle_c):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 3, 4, generator=g)
    b = torch.randn(5, generator=g)
    ref = torch.randn(2, 5, 3, 4, generator=g)
    for act, grad in [(3, 0), (3, 1), (3, 2), (1, 0), (1, 1)]:
        want = ops.bias_act(x, b, ref, act, grad, 0.2, 2 ** 0.5).numpy()
        out = np.empty_like(want)
        oracle_c.oracle_bias_act(_vp(x.numpy()), _vp(b.numpy()), _vp(ref.numpy()), _vp(out), ctypes.c_long(x.numel()),
                                 ctypes.c_long(12), ctypes.c_long(5), act, grad, ctypes.c_float(0.2),
                                 ctypes.c_float(2 ** 0.5))
        assert np.array_equal(out, want), (act, grad)
    # the composite used by the model == kernel mode (act=3, grad=0); its autograd == kernel mode (3, 1)
    xr = x.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    y = ops.fused_leaky_relu(xr, br)
    assert torch.equal(y.detach(), ops.bias_act(x, b, None, 3, 0, 0.2, 2 ** 0.5))
    go = torch.randn(y.shape, generator=g)
    gx, gb = torch.autograd.grad(y, [xr, br], go)
    wx, wb = ops.fused_leaky_relu_backward(go, y.detach())
    assert torch.allclose(gx, wx) and torch.allclose(gb, wb, rtol=1e-5, atol=1e-6)


def test_model32_oracle_matches_reference_golden(golden):
    gold = golden("model32_golden.npz")
    size = 32
    gp, dp = synth.g_state(size, 11), synth.d_state(size, 12)
    z, z2, real = synth.latents(2, 21), synth.latents(2, 22), synth.shots(2, size, 5)
    with torch.no_grad():
        img, _ = mo.g_forward(gp, [z], size, randomize_noise=False)
        img_mix, _ = mo.g_forward(gp, [z, z2], size, randomize_noise=False, inject_index=3)
        assert np.array_equal(img.numpy(), gold["img"])
        assert np.array_equal(img_mix.numpy(), gold["img_mix"])
        assert np.array_equal(mo.d_forward(dp, img, size).numpy(), gold["logits_fake"])
        assert np.array_equal(mo.d_forward(dp, real, size).numpy(), gold["logits_real"])
    # losses + Fisher estimate on a batch of one (train:236-248)
    gnames, dnames = mo.g_param_names(size), mo.d_param_names(size)
    for n in gnames:
        gp[n].requires_grad_(True)
    for n in dnames:
        dp[n].requires_grad_(True)
    fake, _ = mo.g_forward(gp, [z[:1]], size, randomize_noise=False)
    fp = mo.d_forward(dp, fake, size)
    rp = mo.d_forward(dp, real[:1], size)
    g_loss, d_loss = mo.g_nonsaturating_loss(fp), mo.d_logistic_loss(rp, fp)
    assert np.array_equal(g_loss.detach().numpy(), gold["g_loss"])
    assert np.array_equal(d_loss.detach().numpy(), gold["d_loss"])
    fg = mo.estimate_fisher(g_loss, gp, gnames)
    fd = mo.estimate_fisher(d_loss, dp, dnames)
    for k in gold.files:
        if k.startswith("fg_head/"):
            np.testing.assert_allclose(fg[k[8:]].numpy().reshape(-1)[:64], gold[k], rtol=1e-4, atol=1e-30)
        if k.startswith("fd_head/"):
            np.testing.assert_allclose(fd[k[8:]].numpy().reshape(-1)[:64], gold[k], rtol=1e-4, atol=1e-30)
        if k.startswith("fg_sum/"):
            np.testing.assert_allclose(fg[k[7:]].numpy().sum(dtype=np.float64), gold[k], rtol=1e-5)
        if k.startswith("fd_sum/"):
            np.testing.assert_allclose(fd[k[7:]].numpy().sum(dtype=np.float64), gold[k], rtol=1e-5)
    # R1 and path-length scalars
    real_r = real.clone().requires_grad_(True)
    rp2 = mo.d_forward(dp, real_r, size)
    r1 = mo.d_r1_loss(rp2.view(2, -1).mean(dim=1).unsqueeze(1), real_r)
    np.testing.assert_allclose(r1.detach().numpy(), gold["r1"], rtol=1e-5)
    torch.manual_seed(77)
    fake2, lat = mo.g_forward(gp, [z], size, randomize_noise=False, return_latents=True)
    pl, pm, plen = mo.g_path_regularize(fake2, lat, 0)
    np.testing.assert_allclose(plen.detach().numpy(), gold["path_lengths"], rtol=1e-5)
    np.testing.assert_allclose(pl.detach().numpy(), gold["path_penalty"], rtol=1e-4)


def test_g256_oracle_matches_reference_golden(golden):
    gold = golden("g256_golden.npz")
    lat = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "fisher_latents.npy")))
    gp, dp = synth.g_state(256, 1), synth.d_state(256, 2)
    with torch.no_grad():
        img, _ = mo.g_forward(gp, [lat[:2]], 256, randomize_noise=False)
        assert np.array_equal(img[:, :, ::4, ::4].numpy(), gold["img_sub4"])
        assert np.array_equal(mo.d_forward(dp, img, 256).numpy(), gold["logits"])


def _unpack(bits, n):
    return np.unpackbits(bits)[:n].astype(bool)


def test_rick_mask_oracle_matches_reference_code_golden(golden, oracle_c):
    """rick_oracle.decide_* against index sets produced by EXECUTING train_dynamic_update_prune.py:277-393."""
    gold = golden("rick_masks_golden.npz")
    zero_g = zero_d = None
    for rnd, (sg, sd) in enumerate([(101, 102), (103, 104)]):
        fg, fd = synth.fisher_g(sg), synth.fisher_d(sd)
        fr_g, ft_g, pr_g, lines_g = ro.decide_g(fg, 40.0, 0.1)
        fr_d, ft_d, pr_d, lines_d = ro.decide_d(fd, 40.0, 0.1)
        zero_g = pr_g if rnd == 0 else ro.zero_idx_merge(zero_g, pr_g)
        zero_d = pr_d if rnd == 0 else ro.zero_idx_merge(zero_d, pr_d)
        want_lines = gold[f"r{rnd}/lines"]
        got_lines = np.array([lines_g["cut_conv"], lines_g["prune_conv"], lines_g["cut_fc"], lines_g["prune_fc"],
                              lines_d["cut"], lines_d["prune"]])
        assert np.array_equal(got_lines, want_lines)
        for tag, sets in (("freeze_g", fr_g), ("ft_g", ft_g), ("prune_g", pr_g), ("zero_g", zero_g),
                          ("freeze_d", fr_d), ("ft_d", ft_d), ("prune_d", pr_d), ("zero_d", zero_d)):
            keys = [k for k in gold.files if k.startswith(f"r{rnd}/{tag}/")]
            assert len(keys) == len(sets)
            for k in keys:
                name = k.split("/", 2)[2]
                f = fg if tag.endswith("_g") else fd
                n = f[name].shape[1] if f[name].ndim == 5 else f[name].shape[0]
                m = np.zeros(n, bool)
                m[sets[name]] = True
                assert np.array_equal(m, _unpack(gold[k], n)), k
        # the plain-C restatement of mean / percentile / decide reproduces the same numbers bit for bit
        if rnd == 0:
            fims = []
            for key in ro.g_conv_keys():
                a = np.ascontiguousarray(fg[key][0].reshape(fg[key].shape[1], -1))
                out = np.empty(a.shape[0], np.float32)
                oracle_c.oracle_row_mean_f32(_vp(a), None, _vp(out), ctypes.c_long(a.shape[0]), ctypes.c_long(a.shape[1]))
                assert np.array_equal(out, ro.fim_g_conv(fg, key))
                fims.append(out)
            pooled = np.concatenate(fims).astype(np.float64)
            for q, want in ((40.0, lines_g["cut_conv"]), (0.1, lines_g["prune_conv"])):
                got = oracle_c.oracle_percentile_linear(_vp(pooled), ctypes.c_long(pooled.size), ctypes.c_double(q))
                assert got == want
            st = np.empty(fims[0].size, np.uint8)
            oracle_c.oracle_decide(_vp(fims[0]), ctypes.c_long(fims[0].size), ctypes.c_double(lines_g["cut_conv"]),
                                   ctypes.c_double(lines_g["prune_conv"]), 0, _vp(st))
            key0 = ro.g_conv_keys()[0]
            assert np.array_equal((st & 1).nonzero()[0], fr_g[key0])
            assert np.array_equal((st & 2).nonzero()[0], pr_g[key0])
            assert np.array_equal((st & 4).nonzero()[0], ft_g[key0])

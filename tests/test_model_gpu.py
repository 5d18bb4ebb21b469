"""Parity of the module layer (rick_b200.stylegan2 on the sm_100a ops) against the reference-generated golden
vectors and the CPU oracle.  Tolerances (BASELINE.json north_star): generated images within 1e-2 max-abs on a
[-1, 1] image in TF32 mode.  Random-init generators are not confined to [-1, 1] (|img| reaches ~10), so the image is
normalised by the reference's max-abs before the 1e-2 bound is applied."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import model_oracle as mo
from oracle import synth

pytestmark = pytest.mark.gpu
IMG_TOL = 1e-2


def _img_err(got, want):
    want = torch.as_tensor(want, dtype=torch.float64)
    got = got.detach().cpu().to(torch.float64)
    return ((got - want).abs().max() / max(1.0, want.abs().max().item())).item()


def _build(size, gseed, dseed):
    from rick_b200 import stylegan2 as sg
    gp, dp = synth.g_state(size, gseed), synth.d_state(size, dseed)
    G = sg.Generator(size, 512, 8)
    D = sg.Discriminator(size)
    G.load_state_dict(gp)          # strict: every reference key present with the reference shape
    D.load_state_dict(dp)
    return G.cuda(), D.cuda(), gp, dp


def test_state_dict_keys_match_reference_names():
    from rick_b200 import stylegan2 as sg
    G, D = sg.Generator(256, 512, 8), sg.Discriminator(256)
    assert [n for n, _ in G.named_parameters()] == mo.g_param_names(256)
    assert [n for n, _ in D.named_parameters()] == mo.d_param_names(256)
    assert set(G.state_dict()) == set(synth.g_state(256, 0))
    assert set(D.state_dict()) == set(synth.d_state(256, 0))
    assert sum(p.numel() for p in G.parameters()) == 30034338      # SURVEY section 8c probe counts
    assert sum(p.numel() for p in D.parameters()) == 28864129


@pytest.mark.parametrize("channels_last", [True, False], ids=["channels_last", "nchw"])
def test_generator32_and_discriminator32_vs_golden(golden, channels_last):
    from rick_b200 import stylegan2 as sg
    sg.set_channels_last(channels_last)
    try:
        _g32_d32(golden)
    finally:
        sg.set_channels_last(True)


def _g32_d32(golden):
    gold = golden("model32_golden.npz")
    G, D, gp, dp = _build(32, 11, 12)
    z, z2, real = synth.latents(2, 21).cuda(), synth.latents(2, 22).cuda(), synth.shots(2, 32, 5).cuda()
    with torch.no_grad():
        img, none = G([z], randomize_noise=False)
        assert none is None
        assert _img_err(img, gold["img"]) < IMG_TOL
        img_mix, _ = G([z, z2], inject_index=3, randomize_noise=False)
        assert _img_err(img_mix, gold["img_mix"]) < IMG_TOL
        # D on the REFERENCE's images so that only D's own error is measured
        lf, feat = D(torch.from_numpy(gold["img"]).cuda())
        lr, _ = D(real)
        assert len(feat) == 1 + 2 * 3 + 1
        np.testing.assert_allclose(lf.cpu().numpy(), gold["logits_fake"], rtol=2e-2, atol=2e-2)
        np.testing.assert_allclose(lr.cpu().numpy(), gold["logits_real"], rtol=2e-2, atol=2e-2)


def test_generator256_vs_golden(golden):
    gold = golden("g256_golden.npz")
    G, D, gp, dp = _build(256, 1, 2)
    lat = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "fisher_latents.npy"))).cuda()
    with torch.no_grad():
        img, _ = G([lat[:2]], randomize_noise=False)
    assert img.shape == (2, 3, 256, 256)
    abs_err = (img[:, :, ::4, ::4].cpu().double() - torch.as_tensor(gold["img_sub4"]).double()).abs().max().item()
    mean, std, amax = gold["img_moments"]
    print(f"256 px image: max-abs error {abs_err:.3e} on a random-init image with max |pixel| {amax:.2f} "
          f"(= {abs_err / amax:.2e} of full scale; a trained generator's image spans [-1, 1])")
    assert _img_err(img[:, :, ::4, ::4], gold["img_sub4"]) < IMG_TOL
    assert abs(img.double().mean().item() - mean) < 1e-2 * amax and abs(img.double().std().item() - std) < 1e-2 * amax
    # D(256) on our image against the reference's logits for ITS image: G's and D's TF32 error compound
    with torch.no_grad():
        logits, _ = D(img)
    np.testing.assert_allclose(logits.cpu().numpy(), gold["logits"], rtol=3e-2, atol=3e-2)


def test_fused_generator256_vs_golden(golden):
    """The tcgen05 executor (rick_b200.fused.FusedGenerator: sample generation and the D step's fake batch) on the same
    golden image, with the absolute error printed."""
    from rick_b200.fused import FusedGenerator
    gold = golden("g256_golden.npz")
    G, D, gp, dp = _build(256, 1, 2)
    lat = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "fisher_latents.npy"))).cuda()
    img, _ = FusedGenerator(G)([lat[:2]], randomize_noise=False)
    abs_err = (img[:, :, ::4, ::4].cpu().double() - torch.as_tensor(gold["img_sub4"]).double()).abs().max().item()
    amax = gold["img_moments"][2]
    print(f"256 px image (tcgen05 executor): max-abs error {abs_err:.3e}, max |pixel| {amax:.2f} ({abs_err / amax:.2e} of full scale)")
    assert _img_err(img[:, :, ::4, ::4], gold["img_sub4"]) < IMG_TOL


def test_fp32_mode_matches_oracle_tightly():
    """With TF32 switched off in the library convs the whole stack must agree with the CPU oracle to fp32 rounding:
    this isolates the sm_100a upfirdn2d / bias-act kernels and the algebraic modulated conv from tensor-core rounding."""
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    os.environ["RICK_CONV_BACKEND"] = "cudnn"
    try:
        import importlib
        from rick_b200 import conv
        importlib.reload(conv)
        G, D, gp, dp = _build(32, 11, 12)
        z = synth.latents(2, 21)
        with torch.no_grad():
            want, _ = mo.g_forward(gp, [z], 32, randomize_noise=False)
            got, _ = G([z.cuda()], randomize_noise=False)
            assert _img_err(got, want) < 2e-5
            wl = mo.d_forward(dp, want, 32)
            gl, _ = D(want.cuda())
            np.testing.assert_allclose(gl.cpu().numpy(), wl.numpy(), rtol=1e-4, atol=1e-4)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
        os.environ.pop("RICK_CONV_BACKEND", None)
        importlib.reload(conv)


def test_estimate_fisher_vs_golden(golden):
    gold = golden("model32_golden.npz")
    G, D, gp, dp = _build(32, 11, 12)
    from rick_b200.adapt import d_logistic_loss, g_nonsaturating_loss
    z, real = synth.latents(2, 21).cuda(), synth.shots(2, 32, 5).cuda()
    fake, _ = G([z[:1]], randomize_noise=False)
    fp, _ = D(fake)
    rp, _ = D(real[:1])
    g_loss, d_loss = g_nonsaturating_loss(fp), d_logistic_loss(rp, fp)
    np.testing.assert_allclose(g_loss.item(), gold["g_loss"], rtol=2e-2, atol=2e-2)
    np.testing.assert_allclose(d_loss.item(), gold["d_loss"], rtol=2e-2, atol=2e-2)
    grads, fg = G.estimate_fisher(g_loss)
    _, fd = D.estimate_fisher(d_loss)
    assert len(grads) == len(list(G.parameters())) and set(fg) == {n for n, _ in G.named_parameters()}
    for k in gold.files:      # grad**2 summaries: TF32 gradients squared -> 10 % on the layer sums
        if k.startswith("fg_sum/"):
            np.testing.assert_allclose(fg[k[7:]].double().sum().item(), gold[k], rtol=0.1)
        if k.startswith("fd_sum/"):
            np.testing.assert_allclose(fd[k[7:]].double().sum().item(), gold[k], rtol=0.1)


def test_r1_and_path_length_second_order_vs_golden(golden):
    gold = golden("model32_golden.npz")
    G, D, gp, dp = _build(32, 11, 12)
    from rick_b200.adapt import d_r1_loss, g_path_regularize
    z, real = synth.latents(2, 21).cuda(), synth.shots(2, 32, 5).cuda()
    real_r = real.clone().requires_grad_(True)
    rp, _ = D(real_r)
    r1 = d_r1_loss(rp.view(2, -1).mean(dim=1).unsqueeze(1), real_r)
    np.testing.assert_allclose(r1.item(), gold["r1"], rtol=5e-2)
    r1.backward()                                            # double backward through D runs
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for n, p in D.named_parameters() if "convs" in n)
    torch.manual_seed(77)
    noise = torch.randn(2, 3, 32, 32)                        # the draw g_path_regularize made in make_golden.py
    fake, lat = G([z], return_latents=True, randomize_noise=False)
    pl, pm, plen = g_path_regularize(fake, lat, 0, noise.cuda())
    np.testing.assert_allclose(plen.detach().cpu().numpy(), gold["path_lengths"], rtol=5e-2)
    pl.backward()                                            # double backward through G runs
    assert all(torch.isfinite(p.grad).all() for n, p in G.named_parameters() if "convs" in n and p.grad is not None)


@pytest.mark.parametrize("batch", [2, 3, 50])
def test_d_pair_equals_two_discriminator_calls(batch):
    """One joint pass over D scores (fake, real) exactly as two separate calls do: the chunk-interleaved stacking keeps
    every minibatch-stddev group inside its own batch (group = min(B, 25): B = 50 -> 25 members x 2 groups)."""
    from rick_b200.adapt import d_pair
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        _, D, _, _ = _build(32, 11, 12)
        fake = synth.shots(batch, 32, 8).cuda() * 0.7
        real = synth.shots(batch, 32, 9).cuda()
        want_f, want_r = D(fake)[0], D(real)[0]
        got_f, got_r = d_pair(D, fake, real)
        assert got_f.shape == want_f.shape and got_r.shape == want_r.shape
        torch.testing.assert_close(got_f, want_f, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(got_r, want_r, rtol=1e-4, atol=1e-5)
        # a wrong grouping shows up immediately: scoring the plain concatenation mixes the two batches' statistics
        if batch == 2:
            mixed = D(torch.cat([fake, real]))[0]
            assert not torch.allclose(mixed[:batch], want_f, rtol=1e-4, atol=1e-5)
    finally:
        torch.backends.cudnn.allow_tf32 = prev

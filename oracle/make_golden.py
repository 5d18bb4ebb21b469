"""TEST INFRASTRUCTURE ONLY -- run the REAL reference on CPU and write tests/golden/*.

Usage (authoring container only; needs /root/reference):   python -m oracle.make_golden

The reference ships no golden vectors.  This script produces them from the
reference's own code so that the oracle (and through it the CUDA path) is pinned to the reference:

  ops_golden.npz        outputs of the reference's ``upfirdn2d_native`` (op/upfirdn2d.py:159-200) on the
                        seeded cases of oracle/synth.py:UPFIRDN_CASES
  model32_golden.npz    reference ``Generator(32)`` / ``Discriminator(32)`` (model_probe_tune.py) images, logits,
                        losses (train:82-118 executed from the reference file) and estimate_fisher summaries
  g256_golden.npz       reference ``Generator(256)`` image (every 4th pixel + moments) and ``Discriminator(256)``
                        logits for the reference's own Fisher latents ``_noise/0000-0004.pt``
  rick_masks_golden.npz freeze / fine-tune / prune / cumulative-zero index sets obtained by EXECUTING
                        train_dynamic_update_prune.py lines 277-393 on seeded synthetic Fisher dicts (two rounds)
  fisher_latents.npy    the reference's fixed Fisher latents (data fixture, 5 x 512 fp32)

Inputs are never stored: every test regenerates them from the same seeds (oracle/synth.py).
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch

from . import model_oracle as mo
from . import ref_loader, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def ops_golden(ref):
    out = {}
    for case in synth.UPFIRDN_CASES:
        name, n, c, h, w, kh, kw, up, down, p0, p1, kind = case
        x, taps = synth.upfirdn_case_inputs(case)
        out[name] = ref.upfirdn2d(x, taps, up, down, (p0, p1)).numpy()
    np.savez_compressed(os.path.join(OUT, "ops_golden.npz"), **out)
    print("ops_golden:", len(out), "cases")


def model32_golden(ref):
    size = 32
    gp, dp = synth.g_state(size, 11), synth.d_state(size, 12)
    G = ref.model.Generator(size, 512, 8)
    D = ref.model.Discriminator(size)
    G.load_state_dict(gp)
    D.load_state_dict(dp)
    fn = ref_loader.train_functions(["d_logistic_loss", "g_nonsaturating_loss", "d_r1_loss", "g_path_regularize"])
    z = synth.latents(2, 21)
    z2 = synth.latents(2, 22)
    real = synth.shots(2, size, 5)
    out = {}
    with torch.no_grad():
        img, _ = G([z], randomize_noise=False)
        img_mix, _ = G([z, z2], inject_index=3, randomize_noise=False)
        out["img"], out["img_mix"] = img.numpy(), img_mix.numpy()
        out["logits_fake"] = D(img)[0].numpy()
        out["logits_real"] = D(real)[0].numpy()
    # losses + Fisher (batch of one, as train:236-248)
    G.zero_grad(); D.zero_grad()
    z1 = z[:1]
    fake, _ = G([z1], randomize_noise=False)
    fp, _ = D(fake)
    rp, _ = D(real[:1])
    g_loss = fn.g_nonsaturating_loss(fp)
    d_loss = fn.d_logistic_loss(rp, fp)
    out["g_loss"], out["d_loss"] = g_loss.detach().numpy(), d_loss.detach().numpy()
    _, fg = G.estimate_fisher(g_loss)
    _, fd = D.estimate_fisher(d_loss)
    for k in ("convs.0.conv.weight", "convs.1.conv.modulation.weight", "convs.5.conv.modulation.bias", "to_rgb1.bias",
              "style.1.weight", "convs.2.noise.weight", "convs.3.activate.bias"):
        v = fg[k].numpy()
        out["fg_sum/" + k] = np.array(v.sum(dtype=np.float64))
        out["fg_head/" + k] = v.reshape(-1)[:64].copy()
    for k in ("convs.1.conv1.0.weight", "convs.2.conv2.1.weight", "convs.3.skip.1.weight", "convs.2.conv2.2.bias",
              "final_linear.0.weight", "final_conv.0.weight"):
        v = fd[k].numpy()
        out["fd_sum/" + k] = np.array(v.sum(dtype=np.float64))
        out["fd_head/" + k] = v.reshape(-1)[:64].copy()
    # R1 (train:464-475) and path-length (train:546-566) scalars, explicit noise
    real_r = real.clone().requires_grad_(True)
    rp2, _ = D(real_r)
    out["r1"] = fn.d_r1_loss(rp2.view(2, -1).mean(dim=1).unsqueeze(1), real_r).detach().numpy()
    torch.manual_seed(77)   # g_path_regularize draws randn_like(fake_img) internally
    fake2, lat = G([z], return_latents=True, randomize_noise=False)
    pl, pm, plen = fn.g_path_regularize(fake2, lat, 0)
    out["path_penalty"], out["path_mean"], out["path_lengths"] = pl.detach().numpy(), pm.numpy(), plen.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "model32_golden.npz"), **out)
    print("model32_golden:", {k: getattr(v, "shape", None) for k, v in list(out.items())[:6]})

    # cross-check the functional oracle against the real modules while we are here
    with torch.no_grad():
        o_img, _ = mo.g_forward(gp, [z], size, randomize_noise=False)
        assert torch.equal(o_img, img), "oracle g_forward != reference Generator"
        assert torch.equal(mo.d_forward(dp, real, size), torch.from_numpy(out["logits_real"]))


def g256_golden(ref):
    size = 256
    gp, dp = synth.g_state(size, 1), synth.d_state(size, 2)
    G = ref.model.Generator(size, 512, 8)
    D = ref.model.Discriminator(size)
    G.load_state_dict(gp)
    D.load_state_dict(dp)
    lat = torch.cat([torch.load(os.path.join(ref_loader.REF_ROOT, "_noise", f"{j:04d}.pt")) for j in range(5)], 0)
    np.save(os.path.join(OUT, "fisher_latents.npy"), lat.numpy())
    out = {}
    with torch.no_grad():
        img, _ = G([lat[:2]], randomize_noise=False)
        out["img_sub4"] = img[:, :, ::4, ::4].numpy()
        out["img_moments"] = np.array([img.double().mean().item(), img.double().std().item(),
                                       img.abs().max().item()])
        out["logits"] = D(img)[0].numpy()
        o_img, _ = mo.g_forward(gp, [lat[:2]], size, randomize_noise=False)
        assert torch.equal(o_img, img), "oracle g_forward(256) != reference Generator(256)"
        assert torch.equal(mo.d_forward(dp, img, size), torch.from_numpy(out["logits"]))
    np.savez_compressed(os.path.join(OUT, "g256_golden.npz"), **out)
    print("g256_golden: moments", out["img_moments"], "logits", out["logits"].ravel())


def _pack(sets, lengths):
    """index sets -> one packed bit array per key."""
    out = {}
    for k, idx in sets.items():
        m = np.zeros(lengths[k], dtype=bool)
        m[np.asarray(idx, dtype=np.int64)] = True
        out[k] = np.packbits(m)
    return out


def rick_masks_golden(ref):
    """Execute the reference's own decision block (train:277-393) twice (init round, merge round)."""
    block = ref_loader.train_source_lines(277, 393)
    code = compile(block, "train_dynamic_update_prune.py[277:393]", "exec")
    helper = ref_loader.train_functions(["zero_idx_merge"])
    args = types.SimpleNamespace(fisher_quantile=40.0, prune_quantile=0.1, warmup_iter=250)
    ns = {"np": np, "args": args, "zero_idx_merge": helper.zero_idx_merge}
    out = {}
    for rnd, (i, sg, sd) in enumerate([(250, 101, 102), (300, 103, 104)]):
        fg, fd = synth.fisher_g(sg), synth.fisher_d(sd)
        ns.update(i=i, filter_fisher_g=fg, filter_fisher_d=fd)
        exec(code, ns)
        len_g = {k: (fg[k].shape[1] if fg[k].ndim == 5 else fg[k].shape[0]) for k in ns["idx_freeze_g"]}
        len_d = {k: fd[k].shape[0] for k in ns["idx_freeze_d"]}
        for tag, sets, ln in (("freeze_g", ns["idx_freeze_g"], len_g), ("ft_g", ns["idx_ft_g"], len_g),
                              ("prune_g", ns["idx_prune_g"], len_g), ("zero_g", ns["zero_filter_idx_g"], len_g),
                              ("freeze_d", ns["idx_freeze_d"], len_d), ("ft_d", ns["idx_ft_d"], len_d),
                              ("prune_d", ns["idx_prune_d"], len_d), ("zero_d", ns["zero_filter_idx_d"], len_d)):
            for k, v in _pack(sets, ln).items():
                out[f"r{rnd}/{tag}/{k}"] = v
        out[f"r{rnd}/lines"] = np.array([ns["cutline_g_conv"], ns["pruneline_g_conv"], ns["cutline_g_fc"],
                                         ns["pruneline_g_fc"], ns["cutline_d_conv"], ns["pruneline_d_conv"]],
                                        dtype=np.float64)
        npr = sum(len(v) for v in ns["idx_prune_g"].values()), sum(len(v) for v in ns["idx_prune_d"].values())
        print(f"rick round {rnd}: lines {out[f'r{rnd}/lines']}, pruned (g,d) = {npr}")
    np.savez_compressed(os.path.join(OUT, "rick_masks_golden.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_loader.load()
    ops_golden(ref)
    model32_golden(ref)
    rick_masks_golden(ref)
    g256_golden(ref)


if __name__ == "__main__":
    main()

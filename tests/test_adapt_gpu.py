"""A short RICK adaptation run (warm-up, two Fisher rounds, R1, path-length, masks, EMA) on the GPU path must track
the CPU oracle consuming the identical random draws.

Stated tolerance.  With Adam beta1 = 0 the first update of every weight is lr * sign(grad), so rounding-level gradient
differences flip individual updates and GAN losses drift apart over iterations even between two fp32 runs.  Measured
on B200 (round 1): TF32 convolutions (PyTorch's GPU default, what the reference runs) drift up to 9 % in the g loss by
iteration 7.  Bounds: iteration 0 within 1e-2; fp32-conv mode within 5e-2 abs + 5 % rel over 12 iterations;
TF32 mode within 5e-2 abs + 20 % rel."""
import numpy as np
import pytest
import torch

from oracle import adapt_oracle as ao
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tf32,rtol", [(False, 5e-2), (True, 0.2)], ids=["fp32_convs", "tf32_convs"])
def test_adaptation_loss_curve_tracks_oracle(tf32, rtol):
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        _run_curve(rtol)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def _run_curve(rtol):
    from rick_b200 import stylegan2 as sg
    from rick_b200.adapt import AdaptConfig, DrawStream, RickAdapter
    size, iters = 32, 12
    cfg = AdaptConfig(size=size, batch=2, warmup_iter=2, fisher_freq=4, num_fisher_img=2, d_reg_every=4, g_reg_every=2,
                      prune_quantile=1.0)
    gp, dp = synth.g_state(size, 1), synth.d_state(size, 2)
    shots = synth.shots(10, size, 0)
    fisher_lat = synth.latents(cfg.num_fisher_img, 9)

    def nets(cls_g, cls_d):
        G, Ge = cls_g(size, 512, 8), cls_g(size, 512, 8)
        D, De = cls_d(size), cls_d(size)
        G.load_state_dict(gp), Ge.load_state_dict(gp), D.load_state_dict(dp), De.load_state_dict(dp)
        return G.cuda(), D.cuda(), Ge.cuda(), De.cuda()

    G, D, Ge, De = nets(sg.Generator, sg.Discriminator)
    gpu = RickAdapter(cfg, G, D, Ge, De)
    cpu = ao.OracleAdapter(cfg, dict(gp), dict(dp), {k: v.clone() for k, v in gp.items()},
                           {k: v.clone() for k, v in dp.items()})
    draws_gpu, draws_cpu = DrawStream(5, "cuda"), DrawStream(5, "cpu")

    curve_gpu, curve_cpu = [], []
    for i in range(iters):
        if (i - cfg.warmup_iter) % cfg.fisher_freq == 0 and i >= cfg.warmup_iter:
            reals = shots[:cfg.num_fisher_img]
            n_cpu = [draws_cpu.layer_noise(1, size) for _ in range(cfg.num_fisher_img)]
            n_gpu = [draws_gpu.layer_noise(1, size) for _ in range(cfg.num_fisher_img)]
            cpu.fisher_round(fisher_lat, reals, n_cpu)
            gpu.fisher_round(fisher_lat.cuda(), reals.cuda(), n_gpu)
            # same Fisher tensor -> same masks is tested bit-exactly elsewhere; here the Fisher tensors differ by
            # TF32 rounding, so compare set sizes and overlap
            fr, ft, pr, zero = gpu.masks_g.index_sets()
            n_freeze_gpu = sum(len(v) for v in fr.values())
            n_freeze_cpu = sum(len(v) for v in cpu.freeze_g.values())
            assert abs(n_freeze_gpu - n_freeze_cpu) <= 2
            inter = sum(len(np.intersect1d(fr[k], cpu.freeze_g[k])) for k in fr)
            assert inter >= 0.97 * n_freeze_cpu
        real = shots[2 * (i % 5):2 * (i % 5) + 2]
        o_cpu = cpu.step(i, real, draws_cpu, explicit_layer_noise=True)
        o_gpu = gpu.step(i, real.cuda(), draws_gpu, explicit_layer_noise=True)
        assert set(o_cpu) == {k for k in o_gpu if k in o_cpu} and set(o_cpu) <= set(o_gpu)
        curve_cpu.append([float(o_cpu["d"]), float(o_cpu["g"])])
        curve_gpu.append([float(o_gpu["d"]), float(o_gpu["g"])])
    curve_cpu, curve_gpu = np.array(curve_cpu), np.array(curve_gpu)
    print("oracle curve\n", curve_cpu, "\ngpu curve\n", curve_gpu)
    assert np.all(np.isfinite(curve_gpu))
    np.testing.assert_allclose(curve_gpu[0], curve_cpu[0], atol=1e-2, rtol=1e-2)
    np.testing.assert_allclose(curve_gpu, curve_cpu, atol=5e-2, rtol=rtol)
    # pruned filters are exactly zero in the adapted generator and stay zero (Adam beta1 = 0 + zeroed grads)
    fr, ft, pr, zero = gpu.masks_g.index_sets()
    named = dict(G.named_parameters())
    for k, idx in zero.items():
        if len(idx):
            p = named[k]
            rows = p[0] if p.dim() == 5 else p
            assert torch.count_nonzero(rows[torch.as_tensor(idx, device="cuda")]) == 0, k

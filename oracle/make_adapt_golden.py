"""TEST INFRASTRUCTURE ONLY -- the 50-iteration 256 px adaptation loss curve of BASELINE.json config 2, from the oracle.

Usage (authoring container; ~30 min on 8 cores):   python -m oracle.make_adapt_golden [--iters 50] [--size 256]

Runs ``oracle.adapt_oracle.OracleAdapter`` -- the CPU restatement of train_dynamic_update_prune.py:193-699, itself pinned
against the reference by tests/test_oracle_golden.py -- on the seeded inputs below and stores, per iteration, every
loss the reference logs (train:596-603): d, g, r1, path, path_length; plus the sizes of the freeze / prune sets of the
Fisher round at iteration 0.  ``tests/test_adapt_gpu.py::test_graphed_256px_curve_tracks_oracle_golden`` replays the
SAME draws (``DrawStream(seed, cpu_seeded=True)`` is bit-identical on every machine) through the executor bench.py
times (CUDA graphs + tcgen05 generator) and compares.

Protocol (must stay in sync with the test):
    cfg      AdaptConfig(size, batch 2, warmup_iter 0, fisher_freq 50, num_fisher_img 5, fisher_quantile 40,
             prune_quantile 0.1, d_reg_every 16, g_reg_every 4, mixing 0.9, lr 0.002)
    weights  synth.g_state(size, 1), synth.d_state(size, 2) for (G, G_ema), (D, D_ema)
    shots    synth.shots(10, size, 0); iteration i trains on shots[2*(i%5) : 2*(i%5)+2]
    Fisher   iteration 0: latents synth.latents(5, 9), reals shots[:5], per-image layer noise drawn from the stream
    draws    DrawStream(5, "cpu"), explicit per-layer noise
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch

from . import adapt_oracle as ao
from . import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
KEYS = ("d", "g", "r1", "path", "path_length")


def protocol_cfg(size: int):
    from rick_b200.adapt import AdaptConfig
    return AdaptConfig(size=size, batch=2, warmup_iter=0, fisher_freq=50, num_fisher_img=5, fisher_quantile=40.0,
                       prune_quantile=0.1, d_reg_every=16, g_reg_every=4, mixing=0.9, lr=0.002)


def run(size: int, iters: int, log=print):
    from rick_b200.adapt import DrawStream
    cfg = protocol_cfg(size)
    gp, dp = synth.g_state(size, 1), synth.d_state(size, 2)
    shots = synth.shots(10, size, 0)
    lat = synth.latents(cfg.num_fisher_img, 9)
    A = ao.OracleAdapter(cfg, dict(gp), dict(dp), {k: v.clone() for k, v in gp.items()},
                         {k: v.clone() for k, v in dp.items()})
    draws = DrawStream(5, "cpu")
    curve = np.full((iters, len(KEYS)), np.nan, dtype=np.float64)
    extra = {}
    for i in range(iters):
        t = time.perf_counter()
        if i % cfg.fisher_freq == 0:
            noise = [draws.layer_noise(1, size) for _ in range(cfg.num_fisher_img)]
            A.fisher_round(lat, shots[:cfg.num_fisher_img], noise)
            if i == 0:
                extra["n_freeze_g"] = sum(len(v) for v in A.freeze_g.values())
                extra["n_freeze_d"] = sum(len(v) for v in A.freeze_d.values())
                extra["n_zero_g"] = sum(len(v) for v in A.zero_g.values())
                extra["n_zero_d"] = sum(len(v) for v in A.zero_d.values())
        j = 2 * (i % 5)
        out = A.step(i, shots[j:j + 2], draws, explicit_layer_noise=True)
        for c, k in enumerate(KEYS):
            if k in out:
                curve[i, c] = float(out[k])
        log(f"iter {i:3d}  " + "  ".join(f"{k} {curve[i, c]:.5f}" for c, k in enumerate(KEYS) if k in out) +
            f"   [{time.perf_counter() - t:.1f} s]")
    return curve, extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    if args.threads:
        torch.set_num_threads(args.threads)
    curve, extra = run(args.size, args.iters, log=lambda s: print(s, flush=True))
    path = os.path.join(OUT, f"adapt{args.size}_curve.npz")
    np.savez_compressed(path, curve=curve, keys=np.array(KEYS), **{k: np.array(v) for k, v in extra.items()})
    print("wrote", path)


if __name__ == "__main__":
    main()

"""FID-style sample generation, batch-sharded over ranks (BASELINE config 3; gan_training/eval.py:31-46).

    python -m rick_b200.sample --n 5000 --batch 64                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 -m rick_b200.sample ...       # rank r takes batches r, r+W, ...

Every rank runs the tcgen05 generator (rick_b200.fused.FusedGenerator) on its batches with no communication; the only
exchange is one all-reduce of float64 feature statistics (sum, outer-product sum, count) at the end -- here of a cheap
stand-in feature (8x8 average-pooled pixels), since the Inception weights the reference downloads are not available
offline.  Prints one JSON line with the whole-job samples/s.
"""
from __future__ import annotations

import argparse
import json

import torch
import torch.nn.functional as F

from . import dist as rdist
from . import stylegan2 as sg
from .adapt import generate_samples


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=5000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--ckpt", type=str, default=None, help="reference checkpoint with a 'g_ema' state_dict")
    ap.add_argument("--trust-pickle", action="store_true",
                    help="allow full unpickling of --ckpt (legacy files; executes code embedded in the file)")
    args = ap.parse_args()
    rank, world, local_rank = rdist.init_from_env()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    torch.manual_seed(1)
    G = sg.Generator(args.size, 512, 8)
    if args.ckpt:
        from .checkpoint import read
        G.load_state_dict(read(args.ckpt, trust_pickle=args.trust_pickle)["g_ema"], strict=False)
    G = G.to(device).eval()
    stats = rdist.FeatureStats(3 * 8 * 8, device)
    it = generate_samples(G, args.n, args.batch, rank=rank, world=world, seed=args.seed, fused=True)
    _, img = next(it)                              # warm-up batch (weights packed, kernels loaded), not timed
    stats.update(F.adaptive_avg_pool2d(img, 8))
    rdist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    n = 0
    for _, img in it:
        n += img.shape[0]
        stats.update(F.adaptive_avg_pool2d(img, 8))
    stats.all_reduce()                             # the one exchange step
    end.record()
    torch.cuda.synchronize()
    ms = torch.tensor([start.elapsed_time(end)], device=device)
    cnt = torch.tensor([float(n)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
    mu, cov = stats.mean_cov()
    if rank == 0:
        print(json.dumps({"metric": "g_samples_per_s", "value": cnt.item() / (ms.item() / 1e3), "unit": "samples/s",
                          "n_gpus": world, "samples_timed": int(cnt.item()), "batch": args.batch, "size": args.size,
                          "feature_mean_norm": float(mu.norm()), "feature_cov_trace": float(cov.trace())}), flush=True)


if __name__ == "__main__":
    main()

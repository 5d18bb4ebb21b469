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


def run(G, n: int = 5000, batch: int = 64, rank: int = 0, world: int = 1, seed: int = 1000, to_host: bool = True,
        stats: bool = True):
    """The sample-generation loop of Evaluator.compute_inception_score (gan_training/eval.py:31-46), batch-sharded:
    rank r generates batches r, r+W, ... with the tcgen05 generator, copies every image to (pinned) host memory as the
    reference does (``s.cpu().numpy()``) and folds a feature of it into float64 sufficient statistics; the ONE exchange
    step is the all-reduce of those statistics.  Returns a dict with the whole-job samples/s (device-timed, max over
    ranks) -- the first batch of every rank is an untimed warm-up."""
    device = next(G.parameters()).device
    feat = rdist.FeatureStats(3 * 8 * 8, device)
    it = generate_samples(G, n, batch, rank=rank, world=world, seed=seed, fused=True)
    size = G.size
    host = [torch.empty(batch, 3, size, size).pin_memory() for _ in range(2)] if to_host else None
    copied = [torch.cuda.Event() for _ in range(2)]
    first = next(it, None)
    if first is not None and stats:
        feat.update(F.adaptive_avg_pool2d(first[1], 8))
    rdist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    count, k = 0, 0
    for _, img in it:
        count += img.shape[0]
        if to_host:                                    # double-buffered D2H of every image (3.9 GB for 5000 samples)
            copied[k % 2].synchronize()
            host[k % 2][:img.shape[0]].copy_(img, non_blocking=True)
            copied[k % 2].record()
        if stats:
            feat.update(F.adaptive_avg_pool2d(img, 8))
        k += 1
    if stats:
        feat.all_reduce()                              # the one exchange step
    end.record()
    torch.cuda.synchronize()
    ms = torch.tensor([start.elapsed_time(end)], device=device)
    cnt = torch.tensor([float(count)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
    mu, cov = feat.mean_cov() if stats else (torch.zeros(1), torch.zeros(1, 1))
    return {"samples_per_s": cnt.item() / max(ms.item() / 1e3, 1e-9), "samples_timed": int(cnt.item()), "batch": batch,
            "n_requested": n, "incl_d2h": bool(to_host), "incl_stats_allreduce": bool(stats), "ms": ms.item(),
            "feature_mean_norm": float(mu.norm()), "feature_cov_trace": float(cov.trace())}


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
    res = run(G, args.n, args.batch, rank, world, args.seed)
    if rank == 0:
        print(json.dumps({"metric": "g_samples_per_s", "value": res["samples_per_s"], "unit": "samples/s", "n_gpus": world,
                          "size": args.size, **res}), flush=True)


if __name__ == "__main__":
    main()

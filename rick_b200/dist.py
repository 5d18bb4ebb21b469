"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch on the B200 box, gloo in
CPU tests).  Only the exchange steps the path really has:

  * sample generation   batch-sharded, NO communication while generating; one all-reduce of the feature sufficient
                        statistics (sum x, sum x x^T, count; float64) when an Inception-style statistic is requested
                        -- the reference gathers nothing because it is single-process (gan_training/eval.py:31-46,
                        metrics/fid_score.py:132-142 computes mean / cov on one host).
  * adaptation          DDP-style gradient averaging: flat fp32 buckets all-reduced (SUM) and scaled by 1/world, issued
                        as soon as backward finishes (the reference's nn.DataParallel re-broadcasts ~235 MB of
                        parameters on every forward instead, train:941-944).
  * Fisher round        images sharded over ranks, one all-reduce (SUM) of the grad^2 accumulators; every rank then
                        derives identical masks from identical Fisher tensors (no further traffic).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; single-process when WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank) if backend == "nccl" else None)
    return rank, world, local_rank


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_batches(n_batches: int, rank: int, world: int) -> List[int]:
    """rank r takes batches r, r+W, r+2W, ...."""
    return list(range(rank, n_batches, world))


def shard_range(n: int, rank: int, world: int) -> range:
    """contiguous split of n items, first ranks take the remainder (Fisher images over ranks)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class FeatureStats:
    """Streaming sufficient statistics of a feature stream (what FID needs: mean and covariance), float64."""

    def __init__(self, dim: int, device):
        self.n = torch.zeros((), dtype=torch.float64, device=device)
        self.sum = torch.zeros(dim, dtype=torch.float64, device=device)
        self.outer = torch.zeros(dim, dim, dtype=torch.float64, device=device)

    def update(self, feats: torch.Tensor):
        f = feats.reshape(feats.shape[0], -1).to(torch.float64)
        self.n += f.shape[0]
        self.sum += f.sum(0)
        self.outer += f.T @ f

    def all_reduce(self):
        if world_size() > 1:
            for t in (self.n, self.sum, self.outer):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return self

    def mean_cov(self):
        """np.mean(act, axis=0), np.cov(act, rowvar=False) (metrics/fid_score.py:140-141) from the statistics."""
        mu = self.sum / self.n
        cov = (self.outer - self.n * torch.outer(mu, mu)) / (self.n - 1)
        return mu, cov


def _allreduce_buckets_(tensors: Sequence[torch.Tensor], scale: float, bucket_bytes: int):
    """In-place all-reduce (SUM) of ``tensors`` in flat buckets, multiplied by ``scale`` afterwards when it is not 1
    (NVSwitch makes cost per-launch, not per-link bound: few large buckets)."""
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        if len(bucket) == 1 and bucket[0].is_contiguous():       # a flat buffer already: reduce it where it lies
            dist.all_reduce(bucket[0], op=dist.ReduceOp.SUM)
            if scale != 1.0:
                bucket[0].mul_(scale)
            bucket, size = [], 0
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if scale != 1.0:
            flat.mul_(scale)
        off = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        bucket, size = [], 0

    for t in tensors:
        if t is None:
            continue
        bucket.append(t)
        size += t.numel() * t.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


def allreduce_mean_(tensors: Sequence[torch.Tensor], bucket_bytes: int = 64 << 20):
    """In-place average of ``tensors`` over ranks."""
    w = world_size()
    if w == 1:
        return
    _allreduce_buckets_(tensors, 1.0 / w, bucket_bytes)


def allreduce_sum_(tensors: Iterable[torch.Tensor], bucket_bytes: int = 128 << 20):
    """In-place SUM over ranks (Fisher accumulators): the exact sum, no scaling pass (a mean followed by ``* world``
    would round twice for world sizes that are not powers of two and could move filters across a percentile)."""
    if world_size() == 1:
        return
    _allreduce_buckets_(list(tensors), 1.0, bucket_bytes)


def barrier():
    if world_size() > 1:
        dist.barrier()

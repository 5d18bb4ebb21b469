"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch on the B200 box, gloo in
CPU tests).  Only the exchange steps the path really has:

  * sample generation   batch-sharded, NO communication while generating; one all-reduce of the feature sufficient
                        statistics (sum x, sum x x^T, count; float64) when an Inception-style statistic is requested
                        -- the reference gathers nothing because it is single-process (gan_training/eval.py:31-46,
                        metrics/fid_score.py:132-142 computes mean / cov on one host).
  * adaptation          DDP-style gradient averaging, overlapped with the backward pass (``GradSync``): the trainable
                        parameters are cut into a few buckets in the order backward produces them; the moment a
                        bucket's last gradient lands, ONE grouped NCCL all-reduce (AVG, in place on the gradient
                        tensors -- no flat copy, no scaling pass) is issued asynchronously on the process group's own
                        stream while backward continues on the compute stream; the optimiser step waits for the
                        handles.  Recorded inside the CUDA graphs as a fork / join.  (The reference's
                        nn.DataParallel re-broadcasts ~235 MB of parameters on every forward instead, train:941-944.)
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


class GradSync:
    """The DDP exchange step of one network, overlapped with its backward pass.

    ``params`` are the trainable parameters in registration (= forward) order; backward produces their gradients
    roughly in reverse, so bucket 0 holds the LAST layers; buckets are exchanged in index order.  Usage per pass::

        sync.begin()                       # arm the hooks
        autograd.backward(loss, inputs=params)
        sync.finish()                      # launch whatever has not fired, wait for every handle

    A bucket is launched from the post-accumulate-grad hook of the parameter that completes it: one coalesced
    all-reduce over the bucket's gradient tensors where they lie (``ncclGroupStart/End`` around one all-reduce per
    tensor: a single NCCL kernel), ``async_op=True`` so that it runs on the process group's communication stream behind
    an event dependency on the compute stream.  NCCL averages natively (``ReduceOp.AVG``); gloo (CPU tests) sums and the
    1 / world multiply happens in ``finish``.  Every rank runs the same program, so hooks fire in the same order
    everywhere; gradients a pass does not produce (restricted backward) are simply absent from their bucket on all
    ranks alike."""

    def __init__(self, params: Sequence[torch.Tensor], n_buckets: int = 4, late: Sequence[torch.Tensor] = ()):
        """``late``: parameters whose gradients only land when the whole backward pass is over (they hang off one
        multi-output autograd node, e.g. all modulation layers of the generator): kept out of the early buckets so
        that those can complete -- and start their exchange -- while backward is still running."""
        self.params = list(params)
        self.world = world_size()
        late_ids = {id(p) for p in late}
        rev = [p for p in reversed(self.params) if id(p) not in late_ids]
        tail = [p for p in self.params if id(p) in late_ids]
        total = sum(p.numel() for p in rev)
        self.buckets: List[List[torch.Tensor]] = [[]]
        acc = 0
        for p in rev:
            if self.buckets[-1] and len(self.buckets) < n_buckets and acc >= total * len(self.buckets) / n_buckets:
                self.buckets.append([])
            self.buckets[-1].append(p)
            acc += p.numel()
        if tail:
            self.buckets.append(tail)
        self.buckets = [b for b in self.buckets if b]
        self._bucket_of = {id(p): k for k, b in enumerate(self.buckets) for p in b}
        self._pending = [0] * len(self.buckets)
        self._launched = [True] * len(self.buckets)
        self._next = 0
        self._works: list = []
        self._armed = False
        self._avg = None
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params] if self.world > 1 else []

    def begin(self):
        if self.world == 1:
            return
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._next = 0
        self._works = []
        self._armed = True

    def _hook(self, p):
        if not self._armed:
            return
        self._pending[self._bucket_of[id(p)]] -= 1
        # buckets are launched strictly in index order, so the sequence of collectives is the same on every rank
        # whatever order the gradients arrive in
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def _launch(self, k: int):
        self._launched[k] = True
        grads = [p.grad for p in self.buckets[k] if p.grad is not None]
        if not grads:
            return
        if self._avg is None:
            self._avg = dist.get_backend() == "nccl"
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        from torch.distributed.distributed_c10d import _coalescing_manager
        with _coalescing_manager(async_ops=True) as cm:
            for g in grads:
                dist.all_reduce(g, op=op)
        self._works.append((cm, grads))

    def finish(self):
        if self.world == 1:
            return
        self._armed = False
        for k in range(len(self.buckets)):
            if not self._launched[k]:
                self._launch(k)
        for cm, grads in self._works:
            cm.wait()
            if not self._avg:
                torch._foreach_mul_(grads, 1.0 / self.world)
        self._works = []

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


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

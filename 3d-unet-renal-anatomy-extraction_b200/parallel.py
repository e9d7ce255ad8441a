"""Multi-GPU plumbing over torch.distributed (one process per GPU; NCCL on the box, gloo in CPU tests).

The reference has no multi-GPU code.  The path shards naturally:
  * training is data parallel over patches -- InstanceNorm and dropout are per sample, so the only
    exchange is the gradient all-reduce (average) of the parameters that received a gradient;
    the 22 never-used skip_conv tensors keep grad=None and are skipped, as in the reference;
  * inference deals the windows of one volume to the ranks and sums the partial blend buffers.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist

BUCKET_BYTES = 64 << 20


def rank_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(items: list, rank: int, world: int) -> list:
    """Round-robin share of a work list (windows of a volume, cases of a dataset)."""
    return items[rank::world]


def all_reduce_sum(tensors: Iterable[torch.Tensor]):
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def gradient_buckets(params: List[torch.nn.Parameter], bucket_bytes: int = BUCKET_BYTES) -> List[List[torch.nn.Parameter]]:
    """Parameters that have a gradient, in reverse registration order (roughly the order backward
    produces them), cut into buckets of about `bucket_bytes`."""
    live = [p for p in reversed(params) if p.grad is not None]
    buckets, cur, size = [], [], 0
    for p in live:
        cur.append(p)
        size += p.grad.numel() * p.grad.element_size()
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    return buckets


def all_reduce_gradients(model: torch.nn.Module, average: bool = True):
    """Bucketed gradient all-reduce (mean over ranks).  No-op in a single-process run."""
    rank, world = rank_world()
    if world == 1:
        return
    works = []
    for bucket in gradient_buckets(list(model.parameters())):
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bucket))
    for work, flat, bucket in works:
        work.wait()
        if average:
            flat.div_(world)
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n

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


def shard_contiguous(items: list, rank: int, world: int) -> list:
    """Contiguous share of an ordered work list, sizes differing by at most one.  The windows of a volume are listed
    z-fastest / x-slowest (trainer.py:54-65), so a contiguous share is an x-slab: a rank uploads only its slab."""
    n = len(items)
    lo = rank * n // world
    hi = (rank + 1) * n // world
    return items[lo:hi]


def all_reduce_sum(tensors: Iterable[torch.Tensor]):
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def broadcast_object(obj, src: int = 0):
    """`obj` of rank `src` on every rank (a picklable host object: index lists, small dicts).  Identity in a
    single-process run."""
    if rank_world()[1] == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def broadcast_parameters(model: torch.nn.Module, src: int = 0):
    """Parameters and buffers of rank `src` on every rank (start of a data-parallel fit: the ranks may have been
    initialised with different seeds).  In place; no-op in a single-process run."""
    if rank_world()[1] == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src)
    for eng in _engines(model):
        eng.invalidate_weights()


def all_reduce_mean_results(sums: dict, count: int, device) -> dict:
    """Epoch means over ALL ranks' steps: `sums` = per-key sums of this rank's step results, `count` = how many steps
    they cover.  Every rank gets the same dict back, so schedulers / best-checkpoint logic stay in lock step."""
    keys = sorted(sums)
    t = torch.tensor([float(sums[k]) for k in keys] + [float(count)], dtype=torch.float64, device=device)
    if rank_world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    n = max(float(t[-1].item()), 1.0)
    return {k: float(t[i].item()) / n for i, k in enumerate(keys)}


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


def _engines(model: torch.nn.Module):
    for m in model.modules():
        eng = getattr(m, "_engine", None)
        if eng is not None and hasattr(eng, "flat_weight_gradients"):
            yield eng


def prescale_gradients(model: torch.nn.Module, enable: bool = True):
    """Fold the data-parallel 1 / world into the engine's weight-gradient unpack (99.9 % of the gradient bytes): the
    all-reduce then only has to SUM, and the separate divide pass over the 336 MB flat buffer disappears.
    all_reduce_gradients(average=True) recognises prescaled engines and divides only what is left (biases, stem, head).
    Engines are created lazily at the first forward: call after it (GraphedTrainStep / Trainer do)."""
    world = rank_world()[1]
    n = 0
    for eng in _engines(model):
        eng.set_grad_prescale((1.0 / world) if (enable and world > 1) else None)
        n += 1
    return n


def overlap_gradient_all_reduce(model: torch.nn.Module, enable: bool = True):
    """Ask the engine(s) inside `model` to hand the first ~2/3 of the weight gradients (head, decoder, bottom level) to an
    asynchronous NCCL all-reduce as soon as they are complete, so that the collective overlaps the encoder half of the
    backward pass; all_reduce_gradients() then waits for it and sends the rest.  Call once after the model is built
    (forward must have run at least once, or call it again later: engines are created lazily)."""
    def hook(t):
        if rank_world()[1] == 1:
            return None
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
    n = 0
    for eng in _engines(model):
        eng.grad_chunk_hook = hook if enable else None
        n += 1
    return n


def _flat_gradient_buffers(model: torch.nn.Module):
    """(flat buffer, pending handle or None) pieces that already hold (as views) the conv weight gradients."""
    out = []
    for eng in _engines(model):
        out.extend(eng.flat_weight_gradients() or [])
    return out


def all_reduce_gradients(model: torch.nn.Module, average: bool = True):
    """Gradient all-reduce (mean over ranks).  No-op in a single-process run.  The engine's flat weight-gradient
    buffer (335 MB of the 336 MB at the default net) is reduced in place with one collective (two when the first
    chunk was already sent during the backward pass, see overlap_gradient_all_reduce); the remaining small tensors
    (biases, stem, head) go through flattened buckets."""
    rank, world = rank_world()
    if world == 1:
        return
    works = []
    engines = list(_engines(model))
    prescaled = bool(engines) and all(e.conv_weight_gradients_prescaled() for e in engines)
    if prescaled and not average:
        raise RuntimeError("gradients were prescaled by 1 / world (parallel.prescale_gradients) but a SUM was requested")
    flats = _flat_gradient_buffers(model)
    spans = [(f.data_ptr(), f.data_ptr() + f.numel() * f.element_size()) for f, _ in flats]
    conv_w = set()
    if prescaled:
        # tensors the engine produced through its weight-gradient path (flat views, or the first backward's own tensors)
        for e in engines:
            conv_w.update(id(w) for _, w in e._pack_bind.values() if not isinstance(w, (list, tuple)))
            conv_w.update(id(t) for _, w in e._pack_bind.values() if isinstance(w, (list, tuple)) for t in w)
    for f, handle in flats:
        if f.numel() == 0:
            continue
        works.append((handle if handle is not None else dist.all_reduce(f, op=dist.ReduceOp.SUM, async_op=True), f, None,
                      average and not prescaled))
    rest = [p for p in model.parameters()
            if p.grad is not None and not any(a <= p.grad.data_ptr() < b for a, b in spans)]
    pre = [p for p in rest if id(p) in conv_w]              # first backward of a shape: prescaled, but not in a flat buffer
    rest = [p for p in rest if id(p) not in conv_w]
    for group, div in ((pre, False), (rest, average)):
        for bucket in gradient_buckets(group):
            flat = torch.cat([p.grad.reshape(-1) for p in bucket])
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bucket, div))
    for work, flat, bucket, div in works:
        work.wait()
        if div:
            flat.div_(world)
        if bucket is None:
            continue
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n


def enable_sync_batchnorm(model: torch.nn.Module, enabled: bool = True):
    """BatchNorm3d variant (ResAttrBNUnet3D, network.py:38-69) under data parallelism: normalise with the statistics of
    the GLOBAL batch (what torch.nn.SyncBatchNorm does around the reference) instead of every rank's own shard.  Each
    norm application then adds one all-reduce of 2 x C float64 sums in the forward pass and one in the backward pass;
    with ``DiceLoss(global_batch=True)`` and summed gradients the result is exactly the single-process gradient on
    the concatenated batch (tests/test_multi_gpu.py).  Every rank must hold the same number of samples.  The
    all-reduces are issued from Python between kernel launches, so this mode does not combine with GraphedTrainStep."""
    net = model.net if hasattr(model, "net") else model
    net.sync_bn = bool(enabled)            # read by the engine at every BatchNorm application
    return model

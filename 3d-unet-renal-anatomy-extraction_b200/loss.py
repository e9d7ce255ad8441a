"""Drop-in replacements for the reference's ``loss.py`` modules, on the fused CUDA loss kernels.

  DiceLoss(weight_c=None, weight_v=None, alpha=0.5, beta=0.5, smooth=1e-7)      loss.py:123-166
  FocalLoss(gamma=2, weight_c=None, weight_v=None)                              loss.py:169-193
  HybirdLoss(gamma=2, weight_c=None, weight_v=None, alpha=0.5, beta=0.5, ...)   loss.py:196-254
  Dice(weight_v=None, alpha=0.5, beta=0.5, smooth=1e-7)   (metric)              loss.py:85-120

forward(input (N, C, ...) fp32 logits, target (N, ...) int64) -> 0-dim tensor (on the logits' device;
the reference returns it on the CPU -- ``.backward()``, ``.item()`` and ``torch.isnan`` behave the same).

One kernel pass computes softmax and the per-class sums {TP, sum p, sum g, focal}; the one-hot
target (24 B/voxel of int64 in the reference, loss.py:27) is never materialised.  As in the reference,
``weight_c`` is accepted and ignored (it is overwritten before use, loss.py:154-155) and the sums
run over the whole batch (batch-Dice).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


_W_CACHE = {}
# The kernels skip labels outside [0, C) silently where the reference's F.one_hot raises (loss.py:27).  Checking costs a
# device synchronisation per loss call, so it is opt-in: set unet3d_b200.loss.CHECK_TARGET_RANGE = True while debugging
# a data pipeline (or export U3D_CHECK_TARGETS=1).
import os as _os
CHECK_TARGET_RANGE = bool(_os.environ.get("U3D_CHECK_TARGETS"))


def _class_weights(c: int, weight_v, device) -> torch.Tensor:
    """L1-normalised class weights (F.normalize(p=1), loss.py:155), cached on the device (no host-to-device copy
    inside the step, so the loss can be captured in a CUDA graph)."""
    key = (c, None if weight_v is None else tuple(float(v) for v in weight_v), str(device))
    if key not in _W_CACHE:
        w = torch.ones(c, dtype=torch.float64) if weight_v is None else torch.as_tensor(weight_v, dtype=torch.float64)
        _W_CACHE[key] = (w / w.abs().sum().clamp_min(1e-12)).to(device)
    return _W_CACHE[key]


class _SegLossFn(torch.autograd.Function):
    """mode bits: 1 = dice term (1 - d), 2 = focal term, 4 = dice metric (+d instead of 1 - d)."""

    @staticmethod
    def forward(ctx, logits, target, w, alpha, beta, smooth, gamma, mode, global_batch=False):
        if not logits.is_cuda:
            raise RuntimeError("unet3d_b200 losses run on CUDA tensors only")
        n, k = logits.shape[:2]
        if k < 2:
            raise RuntimeError("single-channel (sigmoid) losses are not supported: the reference itself "
                               "fails in one_hot for C == 1 (SURVEY.md 3.4)")
        lg = logits.detach().contiguous().float()
        tg = target.detach().contiguous()
        if tg.dtype != torch.int64:
            tg = tg.long()
        v = lg[0, 0].numel()
        if tg.device != lg.device:
            raise RuntimeError(f"logits on {lg.device}, target on {tg.device}")
        if CHECK_TARGET_RANGE and (int(tg.min()) < 0 or int(tg.max()) >= k):
            raise RuntimeError(f"target labels must lie in [0, {k}) (F.one_hot raises here in the reference, loss.py:27)")
        sums = torch.zeros(k, 4, dtype=torch.float64, device=lg.device)
        with torch.cuda.device(lg.device):           # launches go to the current device's stream
            ops.loss_fwd(lg, tg, sums, gamma)
        if global_batch:
            # exact large-batch equivalence under data parallelism: the per-class sums of ALL ranks
            # enter the (nonlinear) Dice ratio, so every rank's logits gradient is the gradient of the one global loss
            # and the parameter gradients must then be SUMMED over ranks (parallel.all_reduce_gradients(average=False))
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(sums, op=dist.ReduceOp.SUM)
                n = n * dist.get_world_size()
        tp, sp, sg, fo = sums[:, 0], sums[:, 1], sums[:, 2], sums[:, 3]
        den = (1 - alpha - beta) * tp + alpha * sg + beta * sp + smooth
        dice = (tp + smooth) / den
        total = torch.zeros((), dtype=torch.float64, device=lg.device)
        a = torch.zeros(k, dtype=torch.float64, device=lg.device)
        b = torch.zeros_like(a)
        f = torch.zeros_like(a)
        ddice_dtp = (den - (tp + smooth) * (1 - alpha - beta)) / (den * den)
        ddice_dsp = -(tp + smooth) * beta / (den * den)
        if mode & 1:
            total = total + (w * (1 - dice)).sum()
            a, b = a - w * ddice_dtp, b - w * ddice_dsp
        if mode & 4:
            total = total + (w * dice).sum()
            a, b = a + w * ddice_dtp, b + w * ddice_dsp
        if mode & 2:
            scale = float(k) / float(n * v)                              # C * mean over rows, loss.py:80
            total = total + (w * fo * scale).sum()
            f = w * scale
        coef = torch.stack([a, b, f, torch.zeros_like(a)], dim=1).float().contiguous()
        ctx.save_for_backward(lg, tg, coef)
        ctx.gamma, ctx.use_focal = gamma, bool(mode & 2)
        return total.float()

    @staticmethod
    def backward(ctx, gout):
        lg, tg, coef = ctx.saved_tensors
        dl = torch.empty_like(lg)
        with torch.cuda.device(lg.device):
            ops.loss_bwd(lg, tg, coef, gout.detach().float().contiguous().view(1), dl, ctx.gamma, ctx.use_focal)
        return dl, None, None, None, None, None, None, None, None


class DiceLoss(nn.Module):
    def __init__(self, weight_c=None, weight_v=None, alpha=0.5, beta=0.5, smooth=1e-7, *, global_batch=False):
        super().__init__()
        self.weight_c, self.weight_v, self.alpha, self.beta, self.smooth = weight_c, weight_v, alpha, beta, smooth
        self.global_batch = global_batch        # not in the reference: Dice over the batch of ALL data-parallel ranks

    def forward(self, input, target):
        w = _class_weights(input.size(1), self.weight_v, input.device)
        return _SegLossFn.apply(input, target, w, self.alpha, self.beta, self.smooth, 0.0, 1, self.global_batch)


class Dice(nn.Module):
    def __init__(self, weight_v=None, alpha=0.5, beta=0.5, smooth=1e-7):
        super().__init__()
        self.weight_v, self.alpha, self.beta, self.smooth = weight_v, alpha, beta, smooth

    def forward(self, input, target):
        w = _class_weights(input.size(1), self.weight_v, input.device)
        return _SegLossFn.apply(input, target, w, self.alpha, self.beta, self.smooth, 0.0, 4)


class FocalLoss(nn.Module):
    def __init__(self, gamma=2, weight_c=None, weight_v=None, *, global_batch=False):
        super().__init__()
        self.gamma, self.weight_c, self.weight_v = gamma, weight_c, weight_v
        self.global_batch = global_batch

    def forward(self, input, target):
        w = _class_weights(input.size(1), self.weight_v, input.device)
        return _SegLossFn.apply(input, target, w, 0.5, 0.5, 1e-7, float(self.gamma), 2, self.global_batch)


class HybirdLoss(nn.Module):
    """(sic) the reference's spelling, loss.py:196."""

    def __init__(self, gamma=2, weight_c=None, weight_v=None, alpha=0.5, beta=0.5, smooth=1e-7, *, global_batch=False):
        super().__init__()
        self.gamma, self.weight_c, self.weight_v = gamma, weight_c, weight_v
        self.alpha, self.beta, self.smooth = alpha, beta, smooth
        self.global_batch = global_batch

    def forward(self, input, target):
        w = _class_weights(input.size(1), self.weight_v, input.device)
        return _SegLossFn.apply(input, target, w, self.alpha, self.beta, self.smooth, float(self.gamma), 3, self.global_batch)


def dice(input, target, alpha=0.5, beta=0.5, smooth=1e-7):
    """Free function of loss.py:32-48 (probabilities / one-hot in, scalar out), used by trainer.evaluate_case
    on CPU arrays: tiny bookkeeping, plain tensor arithmetic on whatever device the inputs live on."""
    p = input.reshape(-1).float()
    g = target.reshape(-1).float()
    tp = (p * g).sum()
    fn = ((1 - p) * g).sum()
    fp = (p * (1 - g)).sum()
    return (tp + smooth) / (tp + alpha * fn + beta * fp + smooth)

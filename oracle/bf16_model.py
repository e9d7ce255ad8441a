"""bf16-storage model of the oracle.  TEST INFRASTRUCTURE ONLY (see unet3d_oracle.py).

Same algorithm as ``unet3d_oracle.resunet3d_forward`` (reference network.py:104-132,549-565), but
every tensor the CUDA path keeps in 16-bit storage is rounded to that type at the same point:
weights fed to the tensor cores, conv outputs before the norm, activations after it, skip-conv
outputs.  All arithmetic stays fp32 (the kernels accumulate in fp32).  It separates the two
questions a parity number mixes up:

  * do the kernels compute the reference algorithm?   -> CUDA vs THIS model, tight tolerance
  * what does 16-bit storage cost on this network?     -> this model vs the fp32 oracle

Measured here (default net, 1x32^3, seed 0; tests/golden/make_golden.py setup): bf16 rounding of the
WEIGHTS ALONE moves the logits by rel-L2 1.2e-2, rounding the conv INPUTS alone by 1.4e-2, the full
bf16 storage model by 2.2e-2 -- so the north-star's 1e-2 logits bar is not reachable by any
bf16-operand implementation on this randomly initialised, 42-norm-deep network; fp16 storage gives 2.8e-3.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import unet3d_oracle as O

Tensor = torch.Tensor


def _q(t: Tensor, dtype) -> Tensor:
    return t if dtype is None else t.to(dtype).float()


def _res_block(sd, pre, x, cin, cout, stride, dt, masks):
    w = lambda k: _q(sd[pre + k], dt)
    if cin != cout or stride != 1:
        skip = _q(F.conv3d(x, w("skip_conv.weight"), sd[pre + "skip_conv.bias"], stride=stride), dt)
    else:
        skip = x
    y = _q(F.conv3d(x, w("conv1.weight"), None, stride=stride, padding=1), dt)     # bias cancels in the norm (S1)
    m = masks.next(y.shape[0], y.shape[1])
    if m is not None:
        y = y * m
    a = _q(O._lrelu(O._inorm(y)), dt)
    y2 = _q(F.conv3d(a, w("conv2.weight"), None, padding=1), dt)
    return _q(O._lrelu(O._inorm(y2) + skip), dt)


def resunet3d_forward(sd: Dict[str, Tensor], x: Tensor, num_pool: int = 4, num_features: int = 30,
                      dtype=torch.bfloat16, masks: Optional[O.DropoutMasks] = None) -> Tensor:
    masks = masks or O.DropoutMasks(train=False)
    pf = O.paired_features(num_pool, num_features)
    n = len(pf)
    x = _q(F.conv3d(x, sd["net.conv.weight"], sd["net.conv.bias"], padding=1), dtype)     # stem: fp32 weights
    skips = []
    for i in range(num_pool):
        for j in range(max(i, 1)):
            x = _res_block(sd, f"net.encode_blocks.{i}.res_blocks.{j}.", x, pf[i][0] if j == 0 else pf[i][1], pf[i][1], 1,
                           dtype, masks)
        skips.append(x)
        x = _res_block(sd, f"net.pool_blocks.{i}.", x, pf[i][1], pf[i + 1][0], 2, dtype, masks)
    for j in range(max(num_pool, 1)):
        x = _res_block(sd, f"net.encode_blocks.{num_pool}.res_blocks.{j}.", x, pf[num_pool][0], pf[num_pool][1], 1, dtype,
                       masks)
    for i in range(num_pool - 1, -1, -1):
        pre = f"net.up_blocks.{i}.conv_trans."
        y = F.conv_transpose3d(x, _q(sd[pre + "up.0.weight"], dtype), sd[pre + "up.0.bias"], stride=2, padding=1)
        y = _q(F.pad(y, (0, 1, 0, 1, 0, 1)), dtype)
        u = _q(O._lrelu(O._inorm(y)), dtype)
        x = _res_block(sd, f"net.decode_blocks.{i}.", torch.cat((u, skips[i]), 1), pf[n - i - 1][0] + pf[i][1],
                       pf[n - i - 1][1], 1, dtype, masks)
    return F.conv3d(x, sd["net.fc.weight"], sd["net.fc.bias"])        # head: fp32 weights, fp32 logits

"""Forward / backward executor of the 3D U-Net family on the library's sm_100a kernels.

Walks the module tree of :class:`network.Unet` in the order of the reference's ``Unet.forward``
(network.py:549-565) and dispatches on the block type

  ResBlock / ResBlockStack   (network.py:374-449)   conv -> dropout -> IN -> LReLU -> conv -> IN -> (+skip) -> LReLU
  ConvBlock / ConvBlockStack (network.py:153-214)   conv -> dropout -> IN -> LReLU
  MaxPoolBlock               (network.py:452-463)   max-pool k2 s2
  UpConcat                   (network.py:298-350)   convT k3 s2 + zero pad -> IN -> LReLU, concat by addressing

enqueueing  conv (tcgen05 shifted GEMM, IN statistics in the epilogue) -> in_finalize -> in_apply (+residual)
per layer; the backward pass is the hand-derived reverse (no autograd graph over activations).
The whole network is ONE ``torch.autograd.Function`` whose outputs are the logits and whose
backward returns the parameter gradients, so ``loss.backward(); optimizer.step()`` in the
reference's step loop (trainer.py:490-496) works unchanged.

Layout: activations 16-bit NDHWC (channels padded to 16), statistics fp64, master weights fp32.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from . import plan as P

IN_EPS = 1e-5      # nn.InstanceNorm3d default (network.py:163,388: no eps passed)
import os as _os
# overlapped all-reduce: the gradient bytes are cut into chunks in the order the backward pass completes them --
#   chunk 1 = head, decoder and bottom level (72 % of the bytes, complete ~40 % into the backward pass),
#   chunk 2 = pooling block 3, encoder level 3, pooling block 2 (up to 97 %), chunk 3 = the rest (~10 MB, sent behind the pass)
# -- and every chunk but the last is all-reduced on NCCL's stream while the backward pass goes on.  Measured, ms per step:
#   2 B200s (18.93 on one GPU): no overlap 19.84; one cut at 0.8 / 0.9 / 0.97: 19.57 / 19.29-19.59 / 19.50; cuts 0.72,0.97: 19.54
#   8 B200s: no overlap 20.11; one cut at 0.9: 19.92; cuts 0.72,0.97: 19.82 (32 CTAs) / 19.79 (16 CTAs)
# What stays exposed (~0.85 ms at 8 GPUs) is SM contention: NCCL's CTAs cannot share an SM with the 227 KB tensor CTAs.
CHUNK_SPLIT_FRACTIONS = tuple(float(v) for v in _os.environ.get("U3D_CHUNK_SPLIT", "0.72,0.97").split(","))
PACK_SYNC_LAYERS = 5   # weight packs made on the main stream at the start of a pass; the rest overlaps the first layers
# InstanceNorm backward of a tensor whose per-sample slice is at most this many bytes runs as ONE cluster kernel instead of
# the reduce + apply pair (levels 3-4 of the default net at 128^3: 2 MB / 0.5 MB per sample); 0 = always the pair.
# Measured (profiles/r02_notes.md): level 4 10.7-12.2 us vs 19.2 for the pair, level 3 15.4-16.8 vs 25.5, level 2 (8 MB)
# 38-55 vs 35 -- hence the 2 MB default; 20 launches fewer per step, ~0.1 ms of kernel time, within the noise of the
# graph-replayed step (these launches are latency chains either way).  U3D_IN_BWD_SMALL overrides.
# Library launches after each overlapped all-reduce chunk whose grids leave NCCL_MAX_CTAS SMs to NCCL (ops.set_sm_limit):
# the persistent GEMM kernels assign tiles to their 148 CTAs statically, so with 16 SMs taken by NCCL the 16 CTAs that
# wait for an SM start when the first ones exit and the launch takes up to twice as long; sized for 132 SMs it takes
# 12 % longer.  Needs NCCL_MAX_CTAS in the environment (bench.py sets 16); "0" = off.  Measured, one step of cfg-2
# (profiles/r02_sm_limit_window.json): 2 B200s 19.26 (off) / 19.20 (20,4) / 19.10 (40,8) / 19.19 (80,16) ms;
# 8 B200s 19.44 (off) / 19.25 (48,10) ms against 18.63 ms on one.
NCCL_WINDOW = tuple(int(v) for v in _os.environ.get("U3D_NCCL_WINDOW", "48,10").split(","))
IN_BWD_SMALL_BYTES = int(_os.environ.get("U3D_IN_BWD_SMALL", str(2 << 20)))


class _ConvOp:
    """Forward, data-gradient and weight-gradient plans of one conv / transposed-conv layer call."""

    def __init__(self, eng, kind: str, ks: int, stride: int, in_C: Sequence[int], out_C: int, grid, skip_k1: bool = False,
                 fwd_only: bool = False):
        self.kind, self.ks, self.stride = kind, ks, stride
        self.in_C, self.out_C = list(in_C), out_C
        self.grid = grid            # (N, D, H, W) tile grid (coarse grid for strided / transposed layers)
        dev = eng.device
        depth = grid[1]
        self.dgrad_in = None
        if fwd_only:
            self.fwd = ops.DeviceConvPlan(P.make_conv_plan("conv_fwd", ks, stride, in_C, [out_C], depth, grid), dev)
            return
        if kind == "conv":
            self.fwd = ops.DeviceConvPlan(P.make_conv_plan("conv_fwd", ks, stride, in_C, [out_C], depth, grid), dev)
            self.dgrad = ops.DeviceConvPlan(P.make_conv_plan("conv_dgrad", ks, stride, [out_C], in_C, depth, grid), dev)
            if skip_k1:     # d(block input) = dgrad_conv1(dy1) + dgrad_skip(g2) in one launch (two A sources)
                self.dgrad_in = ops.DeviceConvPlan(
                    P.make_conv_plan("conv_dgrad", 3, stride, [out_C] * 2, in_C, depth, grid, skip_k1=True), dev)
        else:
            self.fwd = ops.DeviceConvPlan(P.make_conv_plan("convT_fwd", 3, 2, in_C, [out_C], depth, grid), dev)
            self.dgrad = ops.DeviceConvPlan(P.make_conv_plan("convT_dgrad", 3, 2, [out_C], in_C, depth, grid), dev)
        self.wgrad = ops.DeviceWgradPlan(P.make_wgrad_plan(kind, ks, stride, in_C, out_C, grid, eng.num_sms), dev)


class UNetEngine:
    """Bound to one :class:`network.Unet` (the module that owns conv / blocks / fc)."""

    def __init__(self, net, owner=None):
        self.net = net
        self.owner = owner if owner is not None else net      # carries .precision / .training / .last_dropout_masks
        self.device = None
        self.num_sms = 148
        self._ops: Dict[Tuple, _ConvOp] = {}
        self._inv_scale = None
        self._last_was_train = True      # the first forward always packs
        self.grad_chunk_hook = None      # callable(flat prefix tensor) -> handle; set by parallel.overlap_gradient_all_reduce
        # data parallel: 1 / world folded into the weight-gradient unpack (set by parallel.prescale_gradients), so the
        # flat buffer that the all-reduce SUMS already holds this rank's share of the mean -- no 336 MB divide pass
        self.grad_prescale = None        # float or None
        self._prescale_dev = None
        # tests only (teacher-forced block parity, tests/test_block_parity_gpu.py): when a list, every block of the
        # backward pass appends the tensors it consumed and produced; `capture_scale` is the fp16 gradient scale
        self.capture = None
        self.capture_scale = None
        self._cap_drop = {}
        self._reset_caches()

    def _reset_caches(self):
        self._ops = {}
        self._zb_arena, self._zb_used, self._zb_demand, self._zb_size = None, 0, 0, 0
        self._pack_bind, self._pack_table, self._pack_table_key, self._pack_ptrs = {}, None, None, None
        self._pack_table_late, self._pack_late_ids, self._pack_join, self._pack_stream = None, frozenset(), None, None
        self._dw_slots, self._dw_order, self._dw_table, self._dw_table_n = {}, [], None, 0
        self._dw_arena, self._gflat, self._g_total, self._dw_ready = None, None, 0, False
        self._z_arena, self._z_used, self._z_demand, self._z_size = None, 0, 0, 0
        self._last_gflat = None
        self._dw_done = 0
        self._chunks = []              # overlapped all-reduce: [(wgrad calls done when the chunk is complete, flat begin, flat end, UnpackTable)]
        self._chunk_next, self._chunk_handles = 0, []
        self._infer_graphs = {}          # predict_per_patch's captured window forwards, by (batch, channels, patch, precision)
        self._infer_sig = None           # what the packed weights / graphs of the last no-grad forward were built from

    # ------------------------------------------------------------------ public entry
    def run(self, x: torch.Tensor) -> torch.Tensor:
        net = self.net
        if not x.is_cuda:
            raise RuntimeError("unet3d_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != net.conv.in_channels:
            raise RuntimeError(f"expected input (N, {net.conv.in_channels}, D, H, W), got {tuple(x.shape)}")
        if net.conv.in_channels * 27 * P.pad_channels(net.conv.out_channels) * 4 > 44 * 1024:
            raise RuntimeError("the stem kernel keeps all Cin x 27 x C weights in shared memory: in_channels too large")
        np_ = net.num_pool
        D, H, W = x.shape[2:]
        if D % (1 << np_) or H % (1 << np_) or W % (1 << np_):
            raise RuntimeError(f"spatial size {(D, H, W)} must be divisible by 2^num_pool = {1 << np_} "
                               f"(the reference fails in torch.cat, network.py:350)")
        params = list(net.parameters())
        if any(p.device != x.device for p in params):
            raise RuntimeError(f"input on {x.device}, parameters on {params[0].device}: move the model with .to(device)")
        # every launch below goes to the CURRENT device's stream: make the tensors' device current for the duration
        # (a model on cuda:1 must work without a prior torch.cuda.set_device(1), like the reference's PyTorch modules)
        with torch.cuda.device(x.device):
            if self.device != x.device:
                from . import _lib
                self.device = x.device
                self.num_sms = _lib.lib().unet3d_num_sms()
                self._reset_caches()
            need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
            if need_grad or self._last_was_train:
                ops.PACK_EPOCH += 1          # an optimizer step may lie behind us: re-pack the weights (see ops.PACK_EPOCH)
            self._last_was_train = need_grad
            if need_grad:
                return _UNetFn.apply(x, self, *params)
            with torch.no_grad():
                logits, _ = self.forward_impl(x.detach(), save=False)
            if not torch.cuda.is_current_stream_capturing():
                self._infer_sig = self._param_signature()
            return logits

    def invalidate_weights(self):
        """Call after changing parameters IN PLACE behind autograd's back between two no-grad forwards (``p.data.copy_``,
        EMA / SWA averaging, a fused or foreach optimizer step without a training forward in between): tensor version
        counters do not see those writes, so the cached 16-bit weight packs and the captured inference graphs would be
        stale.  ``load_state_dict`` and ``parallel.broadcast_parameters`` call it themselves."""
        ops.PACK_EPOCH += 1
        self._infer_sig = None
        self._infer_graphs = {}

    def _param_signature(self):
        return (ops.PACK_EPOCH, self.act_dtype, self.owner.training,
                tuple((p.data_ptr(), p._version) for p in self.net.parameters()),
                tuple((b.data_ptr(), b._version) for b in self.net.buffers()))

    def replay_is_current(self) -> bool:
        """True when nothing a captured inference graph depends on has changed since the last no-grad forward of this
        engine: no training forward (fused optimizers do not bump tensor versions, so that alone disqualifies), same
        parameter / buffer tensors at the same versions, same precision and mode.  predict_per_patch then replays its
        cached window graph from the first window on instead of running one eager forward per volume."""
        return (not self._last_was_train) and self._infer_sig is not None and self._infer_sig == self._param_signature()

    # ------------------------------------------------------------------ helpers
    @property
    def act_dtype(self):
        """16-bit storage type of activations, packed weights and gradients: bf16 (default) or fp16
        (``model.precision = "fp16"``; what apex O1 gave the reference)."""
        prec = getattr(self.owner, "precision", "bf16")
        if prec not in ("bf16", "fp16"):
            raise RuntimeError(f"precision must be 'bf16' or 'fp16', got {prec!r}")
        return torch.float16 if prec == "fp16" else torch.bfloat16

    def _op(self, key, kind, ks, stride, in_C, out_C, grid, skip_k1=False, fwd_only=False) -> _ConvOp:
        k = (key, tuple(grid))
        if k not in self._ops:
            self._ops[k] = _ConvOp(self, kind, ks, stride, in_C, out_C, grid, skip_k1, fwd_only)
        return self._ops[k]

    def _new_act(self, n, dims, c):
        return torch.empty((n, *dims, P.pad_channels(c)), dtype=self.act_dtype, device=self.device)

    @staticmethod
    def _grad_like(t):
        """Gradient tensors use the storage format of the activations (one tcgen05.mma cannot mix fp16 and bf16
        operands -- it faults with an illegal instruction -- and the weight gradient multiplies the two)."""
        return torch.empty_like(t)

    def _apply(self, y, skip, table, shift=None):
        out = torch.empty_like(y)
        ops.in_apply(y, skip, out, table, shift)
        return out

    def _drop_scale(self, n, c, p):
        """Dropout3d channel mask drawn exactly as F.dropout3d does, padded to Cp."""
        m = torch.empty(n, c, 1, 1, 1, device=self.device, dtype=torch.float32).bernoulli_(1 - p).div_(1 - p)
        self.owner.last_dropout_masks.append(m)
        out = torch.zeros(n, P.pad_channels(c), device=self.device, dtype=torch.float32)
        out[:, :c] = m.view(n, c)
        return out

    def _conv_in(self, op: _ConvOp, inputs, weight, out_dims, drop=None, bias=None, zero_last=False, norm=None):
        """conv (+ statistics) -> norm tables: returns (y, table, bn) with bn = None for InstanceNorm or the
        BatchNorm record (shift table + what the backward pass needs)."""
        n = inputs[0].shape[0]
        y = self._new_act(n, out_dims, op.out_C)
        stats = self._z64(n, y.shape[-1], 2)
        ops.conv_gemm(op.fwd, inputs, self._pw(op.fwd, weight), [y], op.grid,
                      bias=op.fwd.packed_bias(bias), stats=stats, zero_last=zero_last)
        count = out_dims[0] * out_dims[1] * out_dims[2]
        if isinstance(norm, torch.nn.BatchNorm3d):
            table, bn = self._bn_tables(norm, stats, drop, count)
            return y, table, bn
        table = torch.empty(n, y.shape[-1], 2, dtype=torch.float32, device=self.device)
        # (folding this into in_apply's prologue was measured: -42 launches, but the fp64 divide / sqrt per thread made
        # in_apply 0.36 ms per step slower -- 1.07 -> 1.43 ms -- so the table stays a separate 2 us launch)
        ops.in_finalize(stats, drop, table, count, IN_EPS)
        return y, table, None

    # ------------------------------------------------------------------ BatchNorm3d (network.py:38-69 variant)
    def _bn_tables(self, norm, stats, drop, count):
        """BatchNorm3d(affine, running statistics) on top of the InstanceNorm kernels: the conv epilogue's per-(n, c)
        sums are combined over the batch with O(C) tensor arithmetic, and the normalisation is handed to in_apply as
        out = y * scale[n,c] + shift[n,c] (scale carries the Dropout3d mask m, gamma and 1/sigma; shift carries beta and
        the mean -- a dropped channel becomes the constant beta - mean * gamma / sigma, exactly what BatchNorm makes
        of an all-zero channel)."""
        n, cp = stats.shape[0], stats.shape[1]
        c = norm.num_features
        dev = self.device
        m = drop.double() if drop is not None else torch.ones(n, cp, dtype=torch.float64, device=dev)
        gamma = torch.zeros(cp, dtype=torch.float64, device=dev)
        beta = torch.zeros(cp, dtype=torch.float64, device=dev)
        gamma[:c] = norm.weight.detach().double()
        beta[:c] = norm.bias.detach().double()
        train = self.owner.training or not norm.track_running_stats
        cnt = float(n * count)
        if train:
            s0, s1 = (m * stats[..., 0]).sum(0), (m * m * stats[..., 1]).sum(0)
            if getattr(self.net, "sync_bn", False):      # parallel.enable_sync_batchnorm
                # SyncBN: the 2 Cp per-channel sums of every rank are added (one small all-reduce per norm application);
                # every rank contributes the same number of samples, so the count is just multiplied
                from . import parallel
                world = parallel.rank_world()[1]
                if world > 1:
                    both = torch.stack([s0, s1])
                    parallel.all_reduce_sum([both])
                    s0, s1 = both[0], both[1]
                    cnt *= world
            mean = s0 / cnt
            var = (s1 / cnt - mean * mean).clamp_min(0.0)
            if norm.track_running_stats and self.owner.training:
                with torch.no_grad():
                    norm.num_batches_tracked += 1
                    mom = norm.momentum if norm.momentum is not None else 1.0 / float(norm.num_batches_tracked)
                    norm.running_mean.mul_(1 - mom).add_(mean[:c].to(norm.running_mean.dtype), alpha=mom)
                    norm.running_var.mul_(1 - mom).add_((var[:c] * (cnt / max(cnt - 1.0, 1.0))).to(norm.running_var.dtype),
                                                        alpha=mom)
        else:
            mean = torch.zeros(cp, dtype=torch.float64, device=dev)
            var = torch.ones(cp, dtype=torch.float64, device=dev)
            mean[:c] = norm.running_mean.double()
            var[:c] = norm.running_var.double()
        r = torch.rsqrt(var + norm.eps)
        table = torch.zeros(n, cp, 2, dtype=torch.float32, device=dev)
        table[..., 1] = (m * (r * gamma)).float()
        shift = (beta - mean * r * gamma).float().expand(n, cp).contiguous()
        return table, dict(norm=norm, shift=shift, m=m, r=r, mean=mean, gamma=gamma, train=train, cnt=cnt, c=c)

    def _bn_bwd(self, bn, dout, dout2, out, y, grads, zero_last=False):
        """Backward of the above: x_hat = y * (m r) - mean r;  d gamma = sum g x_hat, d beta = sum g;
        dz = gamma r (g - mean(g) - x_hat mean(g x_hat)) (training) or gamma r g (running statistics); dy = m dz is
        given to in_bwd_apply as per-(n, c) coefficients dy = g A + y B + C."""
        n, cp = y.shape[0], y.shape[-1]
        m, r, mean, gamma, c = bn["m"], bn["r"], bn["mean"], bn["gamma"], bn["c"]
        hat = torch.zeros(n, cp, 2, dtype=torch.float32, device=self.device)
        hat[..., 1] = (m * r).float()
        hat_shift = (-mean * r).float().expand(n, cp).contiguous()
        g = self._grad_like(y)
        sums = self._z64(n, cp, 2)
        ops.in_bwd_reduce(dout, dout2, out, y, g, hat, sums, shift=hat_shift)
        s1, s2 = sums[..., 0].sum(0), sums[..., 1].sum(0)
        a = m * (gamma * r)
        if bn["train"]:
            t1, t2 = s1, s2
            if getattr(self.net, "sync_bn", False):
                # the batch statistics couple the ranks: mean(g), mean(g x_hat) run over all ranks' samples (bn["cnt"] is
                # already the global count); d gamma / d beta below stay local -- the gradient all-reduce adds them
                from . import parallel
                if parallel.rank_world()[1] > 1:
                    both = torch.stack([s1, s2])
                    parallel.all_reduce_sum([both])
                    t1, t2 = both[0], both[1]
            g1, g2 = t1 / bn["cnt"], t2 / bn["cnt"]
            b = -a * g2 * (m * r)
            cc = a * (-g1 + mean * r * g2)
        else:
            b = torch.zeros_like(a)
            cc = torch.zeros_like(a)
        coef = torch.stack([a, b, cc], dim=-1).float().contiguous()
        dy = self._grad_like(y)
        dsum = self._z64(cp)
        ops.in_bwd_apply(g, y, dy, hat, sums, dsum, zero_last, coef=coef)
        norm = bn["norm"]
        dg, db = self._unscale(s2[:c].float()), self._unscale(s1[:c].float())
        grads[norm.weight] = dg if norm.weight not in grads else grads[norm.weight] + dg
        grads[norm.bias] = db if norm.bias not in grads else grads[norm.bias] + db
        return g, dy, sums, dsum

    def _norm_bwd(self, bn, dout, dout2, act, residual, y, table, grads, zero_last=False, want_dsum=False):
        """(g, dy, sums, dsum) of one norm + LeakyReLU application; `act` is its output, `residual` says whether a skip
        tensor was added before the activation (only then InstanceNorm has to read `act` for the sign)."""
        if bn is None:
            return self._in_bwd(dout, dout2, act if residual else None, y, table, zero_last, want_dsum)
        return self._bn_bwd(bn, dout, dout2, act, y, grads, zero_last)

    def _in_bwd(self, dout, dout2, out, y, table, zero_last=False, want_dsum=False):
        n, cp = y.shape[0], y.shape[-1]
        sums = self._z64(n, cp, 2)
        dy = self._grad_like(y)
        dsum = self._z64(cp) if want_dsum else None
        if IN_BWD_SMALL_BYTES and not zero_last and not want_dsum and y[0].numel() * 2 <= IN_BWD_SMALL_BYTES:
            # levels 3-4: one launch for both passes (each of the two kernels is launch- / latency-bound there)
            g = None if (out is None and dout2 is None and self.capture is None) else self._grad_like(y)
            ops.in_bwd_small(dout, dout2, out, y, g, dy, table, sums)
            return g, dy, sums, dsum
        if out is None and dout2 is None and self.capture is None:
            # norm without residual input: g = dout * lrelu'(y_hat) is needed by nobody else, so pass 1 only reduces and
            # pass 2 recomputes it from dout (2 bytes per element less to write, 16-bit round trip of g avoided)
            ops.in_bwd_reduce(dout, None, None, y, None, table, sums)
            ops.in_bwd_apply(dout, y, dy, table, sums, dsum, zero_last, g_is_dout=True)
            return None, dy, sums, dsum
        g = self._grad_like(y)
        ops.in_bwd_reduce(dout, dout2, out, y, g, table, sums)
        ops.in_bwd_apply(g, y, dy, table, sums, dsum, zero_last)
        return g, dy, sums, dsum

    def _wgrad(self, op: _ConvOp, xs, dy, param, more=()):
        """Weight gradient of `param`; `more` = further (xs, dy) pairs of the same layer shape whose products are
        accumulated into the same buffer (a weight applied several times: the attention gate's shared conv).
        Steady state: the fp32 accumulator is a slice of one arena zeroed once per backward, and the gradient
        tensor is a view of one flat buffer that ONE batched gather fills at the end of the backward pass
        (_finish_wgrads); the first backward of a shape runs layer by layer and records the layout."""
        pl = op.wgrad.plan
        slot = self._dw_slots.get(id(op.wgrad)) if self._dw_ready else None
        if slot is not None:
            dw = self._dw_arena[slot[0]:slot[0] + pl.dw_numel + 1]
        else:
            dw = torch.zeros(pl.dw_numel + 1, dtype=torch.float32, device=self.device)
        ops.wgrad_gemm(op.wgrad, xs, dy, dw, op.grid)
        for xs2, dy2 in more:
            ops.wgrad_gemm(op.wgrad, xs2, dy2, dw, op.grid)
        if slot is not None:
            self._wgrad_done()
            return self._gflat[slot[1]:slot[1] + param.numel()].view_as(param)
        if id(op.wgrad) not in self._dw_slots:
            self._dw_slots[id(op.wgrad)] = None
            self._dw_order.append((op.wgrad, param.numel()))
        g = dw.index_select(0, op.wgrad.gidx).view_as(param)
        sc = self._unpack_scale()
        return g if sc is None else g.mul_(sc)

    def _begin_wgrads(self):
        """Called at the start of a backward pass: zero the accumulator arena, allocate this step's flat gradient buffer."""
        self._dw_ready = False
        self._dw_done = 0
        self._chunk_next, self._chunk_handles = 0, []
        if self._dw_table is not None and self._dw_table_n == len(self._dw_order):
            self._dw_arena.zero_()
            self._gflat = torch.empty(self._g_total, dtype=torch.float32, device=self.device)
            self._dw_ready = True

    def set_grad_prescale(self, value):
        """value = 1 / world (data parallel averaging folded into the unpack) or None."""
        self.grad_prescale = None if value is None else float(value)
        self._prescale_dev = None
        if value is not None and self.device is not None:
            self._prescale_dev = torch.full((1,), float(value), dtype=torch.float32, device=self.device)

    def _unpack_scale(self):
        """Device scalar the weight-gradient unpack multiplies in: 1 / fp16-gradient-scale and / or the data-parallel 1 / world."""
        if self.grad_prescale is None:
            return self._inv_scale
        if self._prescale_dev is None:
            self._prescale_dev = torch.full((1,), self.grad_prescale, dtype=torch.float32, device=self.device)
        return self._prescale_dev if self._inv_scale is None else self._inv_scale * self._prescale_dev

    def _wgrad_done(self):
        """Bookkeeping after every weight-gradient launch in steady state: when the layers of the FIRST chunk (backward
        order: head, decoder, bottom level = ~2/3 of the gradient bytes) are complete and somebody asked for it
        (parallel.overlap_gradient_all_reduce), unpack that chunk now and hand its slice of the flat buffer to the hook
        -- the NCCL all-reduce of the prefix then overlaps the encoder half of the backward pass."""
        self._dw_done += 1
        if self._dw_ready and self.grad_chunk_hook is not None:
            while self._chunk_next < len(self._chunks) - 1 and self._dw_done >= self._chunks[self._chunk_next][0]:
                _, g0, g1, table = self._chunks[self._chunk_next]
                table.launch(scale=self._unpack_scale(), out_base=self._gflat)
                self._chunk_handles.append((self._gflat[g0:g1], self.grad_chunk_hook(self._gflat[g0:g1])))
                window = NCCL_WINDOW[min(self._chunk_next, len(NCCL_WINDOW) - 1)]
                nccl_ctas = int(_os.environ.get("NCCL_MAX_CTAS", "0"))
                if window > 0 and nccl_ctas > 0 and self._chunk_handles[-1][1] is not None:
                    ops.set_sm_limit(ops.num_sms() - nccl_ctas, window)
                self._chunk_next += 1

    def _finish_wgrads(self):
        self._last_gflat = None
        ops.set_sm_limit(0)
        if self._dw_ready:
            if self.grad_chunk_hook is not None and len(self._chunks) > 1:
                # chunks whose hook has not fired yet (normally only the last one) are unpacked now, un-sent
                rest = []
                for i in range(self._chunk_next, len(self._chunks)):
                    _, g0, g1, table = self._chunks[i]
                    table.launch(scale=self._unpack_scale(), out_base=self._gflat)
                    rest.append((self._gflat[g0:g1], None))
                self._last_gflat = self._chunk_handles + rest
            else:
                self._dw_table.launch(scale=self._unpack_scale(), out_base=self._gflat)
                self._last_gflat = [(self._gflat, None)]
            self._dw_ready = False
        elif self._dw_order and (self._dw_table is None or self._dw_table_n != len(self._dw_order)):
            off = goff = 0
            jobs = []
            for wp, n_param in self._dw_order:                 # 16-byte aligned slices
                self._dw_slots[id(wp)] = (off, goff)
                off += (wp.plan.dw_numel + 1 + 3) // 4 * 4
                goff += (n_param + 3) // 4 * 4
            self._dw_arena = torch.empty(off, dtype=torch.float32, device=self.device)
            self._g_total = goff
            for wp, n_param in self._dw_order:
                o, g = self._dw_slots[id(wp)]
                jobs.append(dict(dw=self._dw_arena[o + wp.plan.unpack["origin"]:o + wp.plan.dw_numel], rowmap=wp.rowmap,
                                 out=4 * g, **{k: v for k, v in wp.plan.unpack.items() if k not in ("rowmap", "origin")}))
            self._dw_table = ops.UnpackTable(jobs, self.device)
            self._dw_table_n = len(self._dw_order)
            # chunked variant for the overlapped all-reduce: cut where the cumulative gradient bytes pass each fraction
            # of CHUNK_SPLIT_FRACTIONS (never inside a layer that accumulates several launches: the cut counts wgrad CALLS)
            total = sum(n for _, n in self._dw_order)
            cuts, acc, fi = [], 0, 0
            for i, (wp, n_param) in enumerate(self._dw_order):
                acc += n_param
                if fi < len(CHUNK_SPLIT_FRACTIONS) and acc >= CHUNK_SPLIT_FRACTIONS[fi] * total and i + 1 < len(jobs):
                    cuts.append(i + 1)
                    while fi < len(CHUNK_SPLIT_FRACTIONS) and acc >= CHUNK_SPLIT_FRACTIONS[fi] * total:
                        fi += 1
            bounds = [0] + cuts + [len(jobs)]
            self._chunks = []
            for a, b in zip(bounds[:-1], bounds[1:]):
                g0 = self._dw_slots[id(self._dw_order[a][0])][1]
                g1 = self._dw_slots[id(self._dw_order[b][0])][1] if b < len(jobs) else goff
                self._chunks.append((b, g0, g1, ops.UnpackTable(jobs[a:b], self.device)))

    def flat_weight_gradients(self):
        """[(flat fp32 tensor, pending all-reduce handle or None), ...]: the buffer(s) the conv weight gradients of the
        most recent backward are views of (None while the first, layer-by-layer backward of a shape has not recorded
        the layout).  One or two all-reduces cover 99.9 % of the gradient bytes without flattening copies; a piece
        with a handle was already sent by grad_chunk_hook during the backward pass."""
        return self._last_gflat

    def conv_weight_gradients_prescaled(self) -> bool:
        """True when the conv weight gradients (flat buffers AND the tensors of the first, layer-by-layer backward)
        were already multiplied by grad_prescale (1 / world): the all-reduce must then only sum them."""
        return self.grad_prescale is not None

    # ------------------------------------------------------------------ packed weights: one batched launch per step
    def _pw(self, dp, w):
        """16-bit tile stream of `w` for plan `dp` (cached per ops.PACK_EPOCH); remembers the pairing so that the next
        steps can pack every layer in one launch (_prepack)."""
        if id(dp) not in self._pack_bind:
            self._pack_bind[id(dp)] = (dp, w)
        if self._pack_join is not None and id(dp) in self._pack_late_ids:
            self._join_pack()               # first layer whose pack was left to the side stream: wait for it here
        return dp.packed_weight(w, self.act_dtype)

    def _join_pack(self):
        if self._pack_join is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._pack_join)
            self._pack_join = None

    def _prepack(self):
        """Pack all known (plan, parameter) pairs whose cached tile stream is stale with one gather_multi launch."""
        if not self._pack_bind:
            return
        capturing = torch.cuda.is_current_stream_capturing()
        dt = self.act_dtype
        binds = list(self._pack_bind.values())
        if capturing:
            # inside a training graph the pack must be part of every replay; only an already built table can be
            # launched there (building one copies job tables to the device)
            if self._pack_table is None or self._pack_table_key != (len(binds), dt) or not torch.is_grad_enabled():
                return
        elif all(dp._w_version == dp.weight_key(w, dt) for dp, w in binds):
            return
        if self._pack_table is None or self._pack_table_key != (len(binds), dt):
            jobs = []
            for dp, w in binds:
                if isinstance(w, (list, tuple)):
                    if len(w) != 2:
                        return
                    src0, src1, n0 = w[0].detach(), w[1].detach(), w[0].numel()
                else:
                    src0, src1, n0 = w.detach(), None, 0
                if src0.dtype != torch.float32 or not src0.is_contiguous() or (src1 is not None and not src1.is_contiguous()):
                    return
                jobs.append(dict(src0=src0.view(-1), src1=None if src1 is None else src1.view(-1), idx=dp.widx,
                                 out=dp.pack_buffer(dt, self.device), n0=n0, mode=int(dt == torch.float16)))
            # The packs of the first few layers (first-use order = forward order: the level-0 encoder block and the
            # first pooling block, tiny weights) are made on the main stream; everything else -- 99 % of the bytes,
            # including all data-gradient streams -- is packed on a side stream WHILE those first layers run (the pack
            # kernel is bound by scattered 4-byte L2 reads and needs no shared memory, so it shares the SMs with the
            # persistent tensor kernels), and the main stream joins at the first layer that needs one of them.
            k = min(PACK_SYNC_LAYERS, len(jobs))
            self._pack_table = ops.GatherTable(jobs[:k], self.device)
            self._pack_table_late = ops.GatherTable(jobs[k:], self.device) if len(jobs) > k else None
            self._pack_late_ids = frozenset(id(dp) for dp, _ in binds[k:])
            self._pack_table_key = (len(binds), dt)
            self._pack_ptrs = [tuple(t.data_ptr() for t in (w if isinstance(w, (list, tuple)) else [w])) for _, w in binds]
        if self._pack_ptrs != [tuple(t.data_ptr() for t in (w if isinstance(w, (list, tuple)) else [w])) for _, w in binds]:
            self._pack_table = None              # a parameter was re-allocated (e.g. .to()): rebuild next time
            return
        self._pack_table.launch()
        if self._pack_table_late is not None:
            if self._pack_stream is None:
                self._pack_stream = torch.cuda.Stream(device=self.device)
            self._pack_stream.wait_stream(torch.cuda.current_stream(self.device))      # fork (also inside a capture)
            with torch.cuda.stream(self._pack_stream):
                self._pack_table_late.launch()
            self._pack_join = self._pack_stream
        for dp, w in binds:
            dp._w_version = dp.weight_key(w, dt)
            dp._packed_in_capture = capturing

    # ------------------------------------------------------------------ zero-initialised fp64 scratch (statistics)
    def _z64(self, *shape):
        """fp64 zeros for statistic accumulators, carved out of one arena zeroed once per pass instead of one fill
        kernel per layer; the arena is sized from the previous pass's demand."""
        n = 1
        for v in shape:
            n *= v
        self._z_demand += (n + 1) // 2 * 2
        if self._z_arena is not None and self._z_used + n <= self._z_arena.numel():
            out = self._z_arena[self._z_used:self._z_used + n].view(*shape)
            self._z_used += (n + 1) // 2 * 2
            return out
        return torch.zeros(*shape, dtype=torch.float64, device=self.device)

    def _z_begin(self):
        size = max(self._z_demand, self._z_size)
        self._z_size = size
        self._z_demand = 0
        self._z_used = 0
        self._z_arena = torch.zeros(size, dtype=torch.float64, device=self.device) if size else None

    def _zero_grad_of(self, p):
        """Gradient of a conv bias that InstanceNorm cancels exactly: exact zeros, carved out of ONE fp32
        arena zeroed once per backward pass (38 fill launches per step at the default net otherwise).  A fresh view per
        step, so autograd's AccumulateGrad can adopt it without a copy (a cached tensor would be cloned)."""
        n = p.numel()
        self._zb_demand += (n + 3) // 4 * 4
        if self._zb_arena is not None and self._zb_used + n <= self._zb_arena.numel():
            out = self._zb_arena[self._zb_used:self._zb_used + n].view_as(p)
            self._zb_used += (n + 3) // 4 * 4
            return out
        return torch.zeros_like(p)

    def _zb_begin(self):
        size = max(self._zb_demand, self._zb_size)
        self._zb_size, self._zb_demand, self._zb_used = size, 0, 0
        self._zb_arena = torch.zeros(size, dtype=torch.float32, device=self.device) if size else None

    def _unscale(self, g):
        return g if self._inv_scale is None else g * self._inv_scale

    # ------------------------------------------------------------------ blocks: forward
    def _block_fwd(self, blk, key, inputs, out_dims, train, save):
        """Returns (output, record).  `inputs` is a list (a channel concat is a list of two tensors)."""
        from . import network as NW
        if isinstance(blk, NW.ResBlock):
            return self._res_block_fwd(blk, key, inputs, out_dims, train, save)
        if isinstance(blk, NW.ResBlockStack):
            recs, cur = [], inputs
            for j, b in enumerate(blk.res_blocks):
                out, r = self._res_block_fwd(b, key + (j,), cur, out_dims, train, save)
                recs.append(r)
                cur = [out]
            return cur[0], recs
        if isinstance(blk, NW.ConvBlock):
            return self._conv_block_fwd(blk, key, inputs, out_dims, train, save)
        if isinstance(blk, NW.ConvBlockStack):
            recs, cur = [], inputs
            for j, b in enumerate(blk.conv_blocks):
                out, r = self._conv_block_fwd(b, key + (j,), cur, out_dims, train, save)
                recs.append(r)
                cur = [out]
            return cur[0], recs
        if isinstance(blk, NW.MaxPoolBlock):
            x = inputs[0]
            n = x.shape[0]
            out = torch.empty((n, *out_dims, x.shape[-1]), dtype=x.dtype, device=x.device)
            idx = torch.empty((n, *out_dims, x.shape[-1]), dtype=torch.uint8, device=x.device)
            ops.maxpool_fwd(x, out, idx)
            return out, ((x.shape, idx) if save else None)
        raise RuntimeError(f"unsupported block type {type(blk).__name__}")

    def _res_block_fwd(self, blk, key, inputs, out_dims, train, save):
        n = inputs[0].shape[0]
        grid = (n, *out_dims)
        in_C = self._split_channels(blk.in_channels, inputs)
        c1 = self._op(key + ("conv1",), "conv", 3, blk.stride, in_C, blk.out_channels, grid, skip_k1=blk.uses_skip_conv)
        c2 = self._op(key + ("conv2",), "conv", 3, 1, [blk.out_channels], blk.out_channels, grid)
        drop = self._drop_scale(n, blk.out_channels, blk.dropout_p) if (train and blk.dropout_p > 0) else None
        if self.capture is not None:
            self._cap_drop[key] = drop
        # conv biases directly followed by InstanceNorm(affine=False) cancel exactly: not applied.
        # In front of BatchNorm they shift the running mean (and count in eval mode): applied.
        bn = isinstance(blk.norm, torch.nn.BatchNorm3d)
        y1, t1, b1 = self._conv_in(c1, inputs, blk.conv1.weight, out_dims, drop=drop,
                                   bias=blk.conv1.bias if bn else None, norm=blk.norm)
        a1 = self._apply(y1, None, t1, None if b1 is None else b1["shift"])
        y2, t2, b2 = self._conv_in(c2, [a1], blk.conv2.weight, out_dims, bias=blk.conv2.bias if bn else None,
                                   norm=blk.norm)      # the block's ONE norm module is applied twice (network.py:401-416)
        if blk.uses_skip_conv:
            sk = self._op(key + ("skip",), "conv", 1, blk.stride, in_C, blk.out_channels, grid)
            s = self._new_act(n, out_dims, blk.out_channels)
            ops.conv_gemm(sk.fwd, inputs, self._pw(sk.fwd, blk.skip_conv.weight), [s], sk.grid,
                          bias=sk.fwd.packed_bias(blk.skip_conv.bias))
        else:
            s = inputs[0]
        out = self._apply(y2, s, t2, None if b2 is None else b2["shift"])
        return out, ((inputs, y1, t1, a1, y2, t2, out, b1, b2) if save else None)

    def _conv_block_fwd(self, blk, key, inputs, out_dims, train, save):
        n = inputs[0].shape[0]
        in_C = self._split_channels(blk.in_channels, inputs)
        op = self._op(key + ("conv",), "conv", 3, 1, in_C, blk.out_channels, (n, *out_dims))
        drop = self._drop_scale(n, blk.out_channels, blk.dropout_p) if (train and blk.dropout_p > 0) else None
        bn = isinstance(blk.norm, torch.nn.BatchNorm3d)
        y, t, b = self._conv_in(op, inputs, blk.conv.weight, out_dims, drop=drop, bias=blk.conv.bias if bn else None,
                                norm=blk.norm)
        a = self._apply(y, None, t, None if b is None else b["shift"])
        return a, ((inputs, y, t, a, b) if save else None)

    # ------------------------------------------------------------------ attention gate (network.py:353-371)
    def _att_ops(self, gate, level, grid):
        c = gate.conv.in_channels
        if gate.conv.out_channels != c:
            raise RuntimeError("AttBlock conv must keep the width")
        k1 = self._op(("att", level), "conv", 1, 1, [c], c, grid)
        k2 = self._op(("att2", level), "conv", 1, 1, [c, c], c, grid, fwd_only=True)
        return c, k1, k2

    def _att_fwd(self, gate, level, skip, au, dims):
        """xs = conv(skip); f = lrelu(conv(skip) + conv(gate)) as ONE two-source GEMM with the weight repeated
        along K and twice the bias; z = conv(f); result = xs * sigmoid(z).  All three use gate.conv."""
        n = skip.shape[0]
        c, k1, k2 = self._att_ops(gate, level, (n, *dims))
        w, b = gate.conv.weight, gate.conv.bias
        wp = self._pw(k1.fwd, w)
        bp = k1.fwd.packed_bias(b)
        xs, f, z, out = (self._new_act(n, dims, c) for _ in range(4))
        ops.conv_gemm(k1.fwd, [skip], wp, [xs], k1.grid, bias=bp)
        ops.conv_gemm(k2.fwd, [skip, au], k2.fwd.packed_weight(torch.cat([w.detach(), w.detach()], 1), self.act_dtype,
                                                           key=(w.data_ptr(), w._version, self.act_dtype, "x2")), [f],
                      k2.grid, bias=bp * 2.0, act=1)
        ops.conv_gemm(k1.fwd, [f], wp, [z], k1.grid, bias=bp)
        ops.att_gate_fwd(xs, z, out)
        return out, (skip, au, xs, f, z)

    def _att_bwd(self, gate, level, rec, d_up, d_out, grads):
        """Returns (d(upsampled), d(skip)) with the gate's contributions folded in."""
        skip, au, xs, f, z = rec
        n, dims = skip.shape[0], tuple(skip.shape[1:4])
        c, k1, _ = self._att_ops(gate, level, (n, *dims))
        w = gate.conv.weight
        cp = xs.shape[-1]
        dxs, dz, df, dpre, t, dskip, dup = (self._grad_like(xs) for _ in range(7))
        sums = self._z64(cp, 2)
        psum = self._z64(cp)
        ops.att_gate_bwd(d_out, xs, z, dxs, dz, sums)
        wp = self._pw(k1.dgrad, w)
        ops.conv_gemm(k1.dgrad, [dz], wp, [df], k1.grid)
        ops.att_mid_bwd(df, f, dxs, dpre, t, psum)
        ops.conv_gemm(k1.dgrad, [t], wp, [dskip], k1.grid)                       # W^T (dxs + dpre)
        ops.conv_gemm(k1.dgrad, [dpre], wp, [dup], k1.grid, addends=[d_up])      # d_up + W^T dpre
        grads[w] = self._wgrad(k1, [f], dz, w, more=[([skip], t), ([au], dpre)])
        grads[gate.conv.bias] = self._unscale((sums[:c, 0] + sums[:c, 1] + 2.0 * psum[:c]).float())
        return dup, dskip

    @staticmethod
    def _split_channels(total: int, inputs) -> List[int]:
        """Real channel counts of the tensors of a concat: [total] or [c_up, total - c_up]; the real count of the
        upsampled tensor is recorded on it by its producer (attribute _c)."""
        if len(inputs) == 1:
            return [total]
        c0 = getattr(inputs[0], "_c")
        return [c0, total - c0]

    # ------------------------------------------------------------------ forward
    def forward_impl(self, x: torch.Tensor, save: bool):
        net, owner = self.net, self.owner
        train = owner.training
        owner.last_dropout_masks = []
        self._prepack()
        self._z_begin()
        N, _, D, H, W = x.shape
        np_ = net.num_pool
        dims = [(D >> i, H >> i, W >> i) for i in range(np_ + 1)]
        x32 = x.contiguous().float()
        tape = {}
        c0 = net.conv.out_channels
        cp0 = P.pad_channels(c0)
        cin = net.conv.in_channels
        w0 = torch.zeros(cin, 27, cp0, device=self.device)             # stem weights as [Cin][27][Cp] fp32 (+ bias [Cp])
        w0[:, :, :c0] = net.conv.weight.detach().reshape(c0, cin, 27).permute(1, 2, 0)
        b0 = torch.zeros(cp0, device=self.device)
        b0[:c0] = net.conv.bias.detach()
        cur = self._new_act(N, dims[0], c0)
        ops.stem_fwd(x32, w0, b0, cur)
        tape["x"] = x32
        skips = []
        for i in range(np_):
            cur, tape[("enc", i)] = self._block_fwd(net.encode_blocks[i], ("enc", i), [cur], dims[i], train, save)
            skips.append(cur)
            cur, tape[("pool", i)] = self._block_fwd(net.pool_blocks[i], ("pool", i), [cur], dims[i + 1], train, save)
        cur, tape[("enc", np_)] = self._block_fwd(net.encode_blocks[np_], ("enc", np_), [cur], dims[np_], train, save)
        for i in range(np_ - 1, -1, -1):
            up = net.up_blocks[i]
            ct = up.conv_trans.up[0]
            uop = self._op(("up", i), "convT", 3, 2, [ct.in_channels], ct.out_channels, (N, *dims[i + 1]))
            yu, tu, bu = self._conv_in(uop, [cur], ct.weight, dims[i], bias=ct.bias, zero_last=True,
                                       norm=up.conv_trans.up[2])
            au = self._apply(yu, None, tu, None if bu is None else bu["shift"])
            au._c = ct.out_channels
            if save:
                tape[("up", i)] = (cur, yu, tu, au, bu)
            skip = skips[i]
            if getattr(up, "attention", False):
                skip, rec = self._att_fwd(up.att_gate, i, skip, au, dims[i])
                if save:
                    tape[("att", i)] = rec
            cur, tape[("dec", i)] = self._block_fwd(net.decode_blocks[i], ("dec", i), [au, skip], dims[i], train, save)
        K = net.fc.out_channels
        cl = net.fc.in_channels
        wf = torch.zeros(K, P.pad_channels(cl), device=self.device)
        wf[:, :cl] = net.fc.weight.detach().reshape(K, cl)
        logits = torch.empty(N, K, D, H, W, device=self.device, dtype=torch.float32)
        ops.head_fwd(cur, wf, net.fc.bias.detach().float().contiguous(), logits)
        self._join_pack()           # normally joined long ago (first deep layer); a fork must never outlive the pass
        if save:
            tape["head"] = (cur, wf)
        return logits, (tape if save else None)

    # ------------------------------------------------------------------ blocks: backward
    def _block_bwd(self, blk, key, rec, dout, dout2, grads):
        """Returns the list of gradients w.r.t. the block's inputs."""
        from . import network as NW
        if isinstance(blk, NW.ResBlock):
            return self._res_block_bwd(blk, key, rec, dout, dout2, grads)
        if isinstance(blk, (NW.ResBlockStack, NW.ConvBlockStack)):
            blocks = blk.res_blocks if isinstance(blk, NW.ResBlockStack) else blk.conv_blocks
            fn = self._res_block_bwd if isinstance(blk, NW.ResBlockStack) else self._conv_block_bwd
            d, d2 = dout, dout2
            for j in range(len(blocks) - 1, -1, -1):
                dins = fn(blocks[j], key + (j,), rec[j], d, d2, grads)
                d, d2 = dins[0], None
            return dins
        if isinstance(blk, NW.ConvBlock):
            return self._conv_block_bwd(blk, key, rec, dout, dout2, grads)
        if isinstance(blk, NW.MaxPoolBlock):
            shape, idx = rec
            if dout2 is not None:
                dout = dout + dout2
            dx = torch.empty(shape, dtype=dout.dtype, device=dout.device)
            ops.maxpool_bwd(dout, idx, dx)
            return [dx]
        raise RuntimeError(f"unsupported block type {type(blk).__name__}")

    def _res_block_bwd(self, blk, key, rec, dout, dout2, grads):
        inputs, y1, t1, a1, y2, t2, out, b1, b2 = rec
        grid = (inputs[0].shape[0], *y1.shape[1:4])
        in_C = self._split_channels(blk.in_channels, inputs)
        c1 = self._op(key + ("conv1",), "conv", 3, blk.stride, in_C, blk.out_channels, grid, skip_k1=blk.uses_skip_conv)
        c2 = self._op(key + ("conv2",), "conv", 3, 1, [blk.out_channels], blk.out_channels, grid)
        co = blk.out_channels
        g2, dy2, sums2, ds2 = self._norm_bwd(b2, dout, dout2, out, True, y2, t2, grads)
        grads[blk.conv2.weight] = self._wgrad(c2, [a1], dy2, blk.conv2.weight)
        # InstanceNorm cancels the bias exactly (S1); under BatchNorm its gradient is sum(dy) (zero up to rounding in
        # training mode, real in eval mode)
        grads[blk.conv2.bias] = self._zero_grad_of(blk.conv2.bias) if b2 is None else self._unscale(ds2[:co].float())
        da1 = self._grad_like(a1)
        ops.conv_gemm(c2.dgrad, [dy2], self._pw(c2.dgrad, blk.conv2.weight), [da1], c2.grid)
        g1, dy1, _, ds1 = self._norm_bwd(b1, da1, None, a1, False, y1, t1, grads)
        grads[blk.conv1.weight] = self._wgrad(c1, inputs, dy1, blk.conv1.weight)
        grads[blk.conv1.bias] = self._zero_grad_of(blk.conv1.bias) if b1 is None else self._unscale(ds1[:co].float())
        dins = [self._grad_like(t) for t in inputs]
        if blk.uses_skip_conv:
            sk = self._op(key + ("skip",), "conv", 1, blk.stride, in_C, blk.out_channels, grid)
            grads[blk.skip_conv.weight] = self._wgrad(sk, inputs, g2, blk.skip_conv.weight)
            grads[blk.skip_conv.bias] = self._unscale(sums2[:, :blk.out_channels, 0].sum(0).float())
            dp = c1.dgrad_in
            ops.conv_gemm(dp, [dy1, g2], self._pw(dp, [blk.conv1.weight, blk.skip_conv.weight]), dins,
                          c1.grid)
        else:
            ops.conv_gemm(c1.dgrad, [dy1], self._pw(c1.dgrad, blk.conv1.weight), dins, c1.grid,
                          addends=[g2])
        if self.capture is not None:
            self.capture.append(dict(kind="res", key=key, blk=blk, inputs=inputs, in_C=in_C, dout=dout, dout2=dout2,
                                     dins=dins, drop=self._cap_drop.get(key),
                                     fwd=dict(y1=y1, a1=a1, y2=y2, out=out), tab=dict(t1=t1, t2=t2),
                                     bwd=dict(g2=g2, dy2=dy2, da1=da1, g1=g1, dy1=dy1)))
        return dins

    def _conv_block_bwd(self, blk, key, rec, dout, dout2, grads):
        inputs, y, t, a, b = rec
        in_C = self._split_channels(blk.in_channels, inputs)
        op = self._op(key + ("conv",), "conv", 3, 1, in_C, blk.out_channels, (inputs[0].shape[0], *y.shape[1:4]))
        _, dy, _, ds = self._norm_bwd(b, dout, dout2, a, False, y, t, grads)
        grads[blk.conv.weight] = self._wgrad(op, inputs, dy, blk.conv.weight)
        grads[blk.conv.bias] = (self._zero_grad_of(blk.conv.bias) if b is None          # cancelled by InstanceNorm (S1)
                                else self._unscale(ds[:blk.out_channels].float()))
        dins = [self._grad_like(x) for x in inputs]
        ops.conv_gemm(op.dgrad, [dy], self._pw(op.dgrad, blk.conv.weight), dins, op.grid)
        return dins

    # ------------------------------------------------------------------ backward
    def backward_impl(self, tape, x_shape, dlogits: torch.Tensor) -> Dict[torch.nn.Parameter, torch.Tensor]:
        net = self.net
        np_ = net.num_pool
        N, _, D, H, W = x_shape
        dims = [(D >> i, H >> i, W >> i) for i in range(np_ + 1)]
        grads: Dict[torch.nn.Parameter, torch.Tensor] = {}
        self._z_begin()
        self._zb_begin()
        self._begin_wgrads()
        # fp16 gradients need a scale to stay inside fp16's range (Dice gradients are ~1e-7 per voxel).  It is
        # internal and dynamic: a power of two that puts max|dlogits| at 64, taken from this step's dlogits on the
        # device (no host sync), multiplied in by head_bwd and divided out of every parameter gradient.
        self._inv_scale = None
        gscale = None
        if self.act_dtype == torch.float16:
            amax = dlogits.detach().abs().amax().clamp_min(1e-30)
            gscale = torch.exp2(torch.floor(torch.log2(64.0 / amax))).clamp(1.0, 2.0 ** 40).float().reshape(1)
            self._inv_scale = (1.0 / gscale)
        a_last, wf = tape["head"]
        K, cl = net.fc.out_channels, net.fc.in_channels
        d_cur = self._grad_like(a_last)
        dwf = torch.zeros(K * wf.shape[1] + K, device=self.device)
        ops.head_bwd(dlogits.contiguous(), a_last, wf, d_cur, dwf, gscale)
        if self.capture is not None:
            self.capture_scale = gscale
            self.capture.append(dict(kind="head", key=("head",), a=a_last, dlogits=dlogits, din=d_cur))
        grads[net.fc.weight] = dwf[:K * wf.shape[1]].view(K, wf.shape[1])[:, :cl].reshape(net.fc.weight.shape)
        grads[net.fc.bias] = dwf[K * wf.shape[1]:].clone()
        pending = {}
        for i in range(np_):
            d_up, d_skip = self._block_bwd(net.decode_blocks[i], ("dec", i), tape[("dec", i)], d_cur, None, grads)
            if ("att", i) in tape:
                d_up0, d_gated = d_up, d_skip
                d_up, d_skip = self._att_bwd(net.up_blocks[i].att_gate, i, tape[("att", i)], d_up, d_skip, grads)
                if self.capture is not None:
                    self.capture.append(dict(kind="att", key=("att", i), gate=net.up_blocks[i].att_gate,
                                             skip=tape[("att", i)][0], au=tape[("att", i)][1], d_up_in=d_up0,
                                             dout=d_gated, dup=d_up, dskip=d_skip,
                                             fwd=dict(zip(("xs", "f", "z"), tape[("att", i)][2:5]))))
            pending[i] = d_skip
            xin, yu, tu, au, bu = tape[("up", i)]
            ct = net.up_blocks[i].conv_trans.up[0]
            uop = self._op(("up", i), "convT", 3, 2, [ct.in_channels], ct.out_channels, (N, *dims[i + 1]))
            _, dyu, _, dsum = self._norm_bwd(bu, d_up, None, au, False, yu, tu, grads, zero_last=True, want_dsum=True)
            grads[ct.weight] = self._wgrad(uop, [xin], dyu, ct.weight)
            grads[ct.bias] = self._unscale(dsum[:ct.out_channels].float())
            d_cur = self._grad_like(xin)
            ops.conv_gemm(uop.dgrad, [dyu], self._pw(uop.dgrad, ct.weight), [d_cur], uop.grid)
            if self.capture is not None:
                self.capture.append(dict(kind="up", key=("up", i), ct=ct, xin=xin, dout=d_up, din=d_cur,
                                         fwd=dict(y=yu, a=au), tab=dict(t=tu)))
        d_cur = self._block_bwd(net.encode_blocks[np_], ("enc", np_), tape[("enc", np_)], d_cur, None, grads)[0]
        for i in range(np_ - 1, -1, -1):
            d_cur = self._block_bwd(net.pool_blocks[i], ("pool", i), tape[("pool", i)], d_cur, None, grads)[0]
            # the level's encoder output fed both the pooling block and the skip connection
            d_cur = self._block_bwd(net.encode_blocks[i], ("enc", i), tape[("enc", i)], d_cur, pending[i], grads)[0]
        c0 = net.conv.out_channels
        cp0 = d_cur.shape[-1]
        cin = net.conv.in_channels
        dw0 = torch.zeros(cin, 28, cp0, device=self.device)
        ops.stem_wgrad(tape["x"], d_cur, dw0)
        if self.capture is not None:
            self.capture.append(dict(kind="stem", key=("stem",), x=tape["x"], dout=d_cur))
        grads[net.conv.weight] = self._unscale(dw0[:, :27, :c0].permute(2, 0, 1).reshape(net.conv.weight.shape))
        grads[net.conv.bias] = self._unscale(dw0[0, 27, :c0].clone())
        self._finish_wgrads()
        self._inv_scale = None
        return grads


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng: UNetEngine, *params):
        logits, tape = eng.forward_impl(x.detach(), save=True)
        ctx.eng, ctx.tape, ctx.x_shape, ctx.params = eng, tape, tuple(x.shape), params
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        with torch.cuda.device(dlogits.device):
            grads = ctx.eng.backward_impl(ctx.tape, ctx.x_shape, dlogits)
        ctx.tape = None
        out = [grads.get(p) if p.requires_grad else None for p in ctx.params]
        return (None, None, *out)

"""Torch-tensor level wrappers over the C ABI: device-resident plans and one function per kernel.

Activations are ``torch.bfloat16`` tensors of shape (N, D, H, W, Cp), contiguous, Cp = pad_channels(C).
Nothing here computes on the host or through PyTorch ops: every function enqueues one of the
library's kernels on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import plan as P

_err_words: Dict[int, torch.Tensor] = {}

# bench.py instrumentation: when PROFILE is a list, every launch of the library appends
# (kernel name, algorithmic FLOPs, start event, end event, algorithmic HBM bytes); LAUNCHES counts every kernel enqueued.
PROFILE = None
PROFILE_DETAIL = None   # tools/step_table.py: when a list, one descriptor string per PROFILE entry (layer shape, plan)
LAUNCHES = 0

# Packed-weight caches are keyed by (address, tensor version, dtype, PACK_EPOCH).  The version counter alone is NOT
# enough: fused optimizers (torch.optim.Adam(fused=True)) update parameters without bumping it.  The engine bumps
# PACK_EPOCH at every training-mode forward and at the first no-grad forward after one, so a pack is reused only
# across forwards between which no optimizer step can have happened (e.g. the windows of one inference volume).
PACK_EPOCH = 0
BIAS_EPOCH = 0          # bumped whenever a plan's packed bias vector is re-created (BatchNorm variant: conv biases are applied)
DBG_OUT = None          # tuning experiments: int64[8] device tensor receiving the conv MMA warp's cycle counters
# bench.py instrumentation: padded -> real channel count of the benchmarked net ({32: 30, 64: 60, 128: 120} for the
# default ResUnet3D).  SURVEY.md 8d makes UNPADDED channels the algorithmic figure, so the HBM-kernel byte counts that
# go with PROFILE are scaled by real / padded channels when the map knows the width (padding is overhead, not credit).
REAL_CHANNELS: Dict[int, int] = {}


def _alg_numel(t: torch.Tensor) -> float:
    """Element count of an NDHWC activation with its REAL channel count (see REAL_CHANNELS)."""
    cp = t.shape[-1]
    return t.numel() * (REAL_CHANNELS.get(cp, cp) / float(cp))


SM_LIMIT = 0            # > 0: grids are being sized for this many SMs (set_sm_limit), until LAUNCHES reaches _SM_LIMIT_UNTIL
_SM_LIMIT_UNTIL = 0


_NUM_SMS = {}


def num_sms() -> int:
    """SM count of the current device (not the set_sm_limit value)."""
    dev = torch.cuda.current_device()
    if dev not in _NUM_SMS:
        if SM_LIMIT:
            set_sm_limit(0)
        _NUM_SMS[dev] = int(_lib.lib().unet3d_num_sms())
    return _NUM_SMS[dev]


def set_sm_limit(limit: int, launches: int = 0):
    """Size the grids of the next `launches` library launches (current device) for `limit` SMs; 0 = back to all SMs.
    Used while a gradient all-reduce overlaps the backward pass: NCCL's CTAs occupy SMs the persistent GEMM kernels
    (one 227 KB CTA per SM, tiles assigned statically) would otherwise wait for."""
    global SM_LIMIT, _SM_LIMIT_UNTIL
    limit = int(limit) if launches > 0 else 0
    if limit != SM_LIMIT:
        _lib.check(_lib.lib().unet3d_set_sm_limit(limit), "unet3d_set_sm_limit")
    SM_LIMIT, _SM_LIMIT_UNTIL = limit, LAUNCHES + int(launches)


def limited_split(n_jobs: int, split: int, limit: int) -> int:
    """Split-K count of a weight-gradient launch (jobs x split CTAs, one per SM) while only `limit` SMs are free: the
    largest split that keeps one wave; unchanged without a limit, if it already fits, or if the jobs alone exceed it."""
    if limit and n_jobs <= limit < n_jobs * split:
        return limit // n_jobs
    return split


def _count(n: int = 1):
    global LAUNCHES
    LAUNCHES += n
    if SM_LIMIT and LAUNCHES > _SM_LIMIT_UNTIL:
        set_sm_limit(0)


class _Timed:
    def __init__(self, name, flops=0.0, nbytes=0.0, detail=""):
        self.name, self.flops, self.nbytes, self.detail = name, flops, nbytes, detail

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.name, self.flops, self.e0, e1, self.nbytes))
            if PROFILE_DETAIL is not None:
                PROFILE_DETAIL.append(self.detail() if callable(self.detail) else self.detail)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device_of_first_arg(fn):
    """Entry points that users call directly with tensors (case-level kernels): run with the tensor's device current, so
    that the launch goes to THAT device's stream.  (The network / loss / predict paths set the device once at their own
    entry: engine.UNetEngine.run, loss._SegLossFn, trainer.predict_*.)"""
    import functools

    @functools.wraps(fn)
    def wrapper(t, *a, **k):
        with torch.cuda.device(t.device):
            return fn(t, *a, **k)
    return wrapper


def err_word(device) -> torch.Tensor:
    idx = torch.device(device).index or 0
    if idx not in _err_words:
        _err_words[idx] = torch.zeros(1, dtype=torch.int32, device=device)
    return _err_words[idx]


def check_device_errors(device=None):
    """Synchronising check of the device error word the pipelined kernels write on a timeout."""
    for idx, w in _err_words.items():
        v = int(w.item())
        if v != 0:
            w.zero_()
            raise _lib.Unet3dError(f"device-side pipeline timeout, role code {v} (see csrc/conv_gemm.cu)")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


ACT_DTYPES = (torch.bfloat16, torch.float16)


def _f16(t: torch.Tensor) -> int:
    return int(t.dtype == torch.float16)


def make_src(t: torch.Tensor, parity=None) -> _lib.Src:
    assert t.dtype in ACT_DTYPES and t.is_contiguous() and t.dim() == 5, (t.dtype, t.shape)
    n, d, h, w, cp = t.shape
    sW = cp * 2
    sH, sD, sN = w * sW, h * w * sW, d * h * w * sW
    s = _lib.Src()
    if parity is None:
        s.ptr, s.C, s.W, s.H, s.D, s.N = t.data_ptr(), cp, w, h, d, n
        s.sW, s.sH, s.sD, s.sN = sW, sH, sD, sN
    else:
        pd, ph, pw = parity
        s.ptr = t.data_ptr() + pd * sD + ph * sH + pw * sW
        s.C, s.W, s.H, s.D, s.N = cp, (w - pw + 1) // 2, (h - ph + 1) // 2, (d - pd + 1) // 2, n
        s.sW, s.sH, s.sD, s.sN = 2 * sW, 2 * sH, 2 * sD, sN
    return s


class DeviceConvPlan:
    """A ConvPlan with its tables resident on one device."""

    def __init__(self, plan: P.ConvPlan, device):
        self.plan = plan
        self.device = device
        # the device copies of a (memoised, immutable) plan's tables are shared by all layers that use the plan
        dev_key = str(torch.device(device))
        cache = plan.__dict__.setdefault("_dev", {})
        if dev_key not in cache:
            cache[dev_key] = (torch.from_numpy(plan.tab).to(device), torch.from_numpy(plan.widx).to(device),
                              torch.from_numpy(P.bias_index(plan)).to(device))
        self.tab, self.widx, self.bidx = cache[dev_key]
        self._w_version = None
        self._w_packed = None
        self._packed_in_capture = False      # the engine's batched pack ran inside the current graph capture
        self._b_version = None
        self._b_packed = None

    @staticmethod
    def weight_key(w, dtype, key=None):
        if isinstance(w, (list, tuple)):
            return tuple((t.data_ptr(), t._version) for t in w) + (dtype, PACK_EPOCH)
        return (w.data_ptr(), w._version, dtype, PACK_EPOCH) if key is None else tuple(key) + (PACK_EPOCH,)

    def pack_buffer(self, dtype, device) -> torch.Tensor:
        """The persistent 16-bit tile-stream buffer of this plan (its address is what batched pack tables hold)."""
        if self._w_packed is None or self._w_packed.dtype != dtype:
            self._w_packed = torch.empty(self.widx.numel(), dtype=dtype, device=device)
            self._w_version = None
        return self._w_packed

    def packed_weight(self, w, dtype=torch.bfloat16, key=None) -> torch.Tensor:
        """16-bit tile stream of parameter `w` (re-gathered only when the parameter, the dtype or ops.PACK_EPOCH changed).
        `w` may be a list of parameters: the plan's gather index then addresses their concatenation.
        `key`: cache key to use when `w` is a temporary derived from parameters (its own address means nothing)."""
        k = self.weight_key(w, dtype, key)
        # a TRAINING graph must re-pack on every replay (the optimizer step is part of it); an inference graph
        # (no_grad) replays against the weights it was captured with
        capturing_train = torch.cuda.is_current_stream_capturing() and torch.is_grad_enabled()
        if self._w_version == k and (not capturing_train or self._packed_in_capture):
            return self._w_packed
        if isinstance(w, (list, tuple)):
            src = torch.cat([t.detach().reshape(-1).float() for t in w])
        else:
            src = w.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
        out = self.pack_buffer(dtype, src.device)
        _count()
        _lib.check(_lib.lib().unet3d_weight_pack(src.data_ptr(), self.widx.data_ptr(), out.data_ptr(), out.numel(),
                                                 int(dtype == torch.float16), _stream()), "unet3d_weight_pack")
        self._w_version = k
        return self._w_packed

    def packed_bias(self, b: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if b is None:
            return None
        key = (b.data_ptr(), b._version, PACK_EPOCH)
        if self._b_version != key or (torch.cuda.is_current_stream_capturing() and torch.is_grad_enabled()):
            flat = torch.cat([b.detach().reshape(-1).float(), b.new_zeros(1, dtype=torch.float32)])
            global BIAS_EPOCH
            if self._b_packed is not None:
                BIAS_EPOCH += 1            # a captured inference graph still reads the old buffer: it must be re-captured
            self._b_packed = flat.index_select(0, self.bidx).contiguous()
            self._b_version = key
        return self._b_packed


def conv_gemm(dp: DeviceConvPlan, inputs: Sequence[torch.Tensor], wpacked: torch.Tensor, outs: Sequence[torch.Tensor],
              grid: Tuple[int, int, int, int], bias: Optional[torch.Tensor] = None,
              addends: Optional[Sequence[Optional[torch.Tensor]]] = None, stats: Optional[torch.Tensor] = None,
              zero_last: bool = False, act: int = 0):
    """Launch the shifted-GEMM kernel for `dp` (see plan.make_conv_plan for the meaning of a plan)."""
    pl = dp.plan
    a = _lib.ConvArgs()
    a.n_src = len(pl.maps)
    for i, (ti, par) in enumerate(pl.maps):
        a.src[i] = make_src(inputs[ti], par)
    a.tab, a.w = dp.tab.data_ptr(), wpacked.data_ptr()
    o0 = outs[0]
    assert o0.dtype in ACT_DTYPES and o0.is_contiguous()
    for o in outs:
        assert o.shape == o0.shape and o.is_contiguous() and o.dtype == o0.dtype
    assert wpacked.dtype == inputs[0].dtype and all(t.dtype == inputs[0].dtype for t in inputs)
    if addends is not None:
        assert all(t is None or t.dtype == o0.dtype for t in addends)
    a.out = o0.data_ptr()
    a.out2 = outs[1].data_ptr() if len(outs) > 1 else None
    a.bias = _ptr(bias)
    if addends is not None:
        a.addend = _ptr(addends[0])
        a.addend2 = _ptr(addends[1]) if len(addends) > 1 else None
    a.stats = _ptr(stats)
    a.err = err_word(o0.device).data_ptr()
    a.N, a.D, a.H, a.W = grid
    a.Dt, a.n_nblk, a.nblk, a.G, a.n_cg, a.n_taps = pl.Dt, pl.n_nblk, pl.nblk, pl.G, pl.n_cg, len(pl.shifts)
    a.fuse = 3 if pl.fuse_kd else 1
    a.nbuf = pl.nbuf
    a.wT, a.w_stages, a.a_stages = pl.wT, pl.w_stages, pl.a_stages
    a.in_f16, a.out_f16 = _f16(inputs[0]), _f16(o0)
    n, d, h, w, cp = o0.shape
    a.out_sW, a.out_sH, a.out_sD, a.out_sN = cp, w * cp, h * w * cp, d * h * w * cp
    a.out_C = cp
    a.stats_C = stats.shape[1] if stats is not None else 0
    a.omul = pl.omul
    a.zD, a.zH, a.zW = (d - 1, h - 1, w - 1) if zero_last else (-1, -1, -1)
    a.act = act
    a.dense = int(pl.dense)
    a.dbg_out = DBG_OUT.data_ptr() if DBG_OUT is not None else None
    flops = pl.flops_per_voxel * grid[0] * grid[1] * grid[2] * grid[3]
    _count()
    with _Timed("conv_gemm_kernel", flops, 0.0,
                lambda: f"{pl.kind} k{pl.ks} s{pl.stride} {pl.in_C}->{pl.out_C} grid{tuple(grid)} nblk{pl.nblk} Dt{pl.Dt} G{pl.G} "
                        f"fuse{3 if pl.fuse_kd else 1} nbuf{pl.nbuf} dense{int(pl.dense)} items{a.N * -(-a.D // pl.Dt) * -(-a.H // P.HT) * -(-a.W // P.WT) * pl.n_nblk}"):
        _lib.check(_lib.lib().unet3d_conv_gemm(C.byref(a), _stream()), "unet3d_conv_gemm")


class GatherTable:
    """Device-resident job table of unet3d_gather_multi: many `out = src[idx]` gathers in one launch."""

    def __init__(self, jobs: Sequence[dict], device):
        """jobs: dicts with src0, idx (tensors), optional src1 (tensor), out (tensor, or an int byte offset that is
        added to the `out_base` given at launch), n0, mode (0 bf16, 1 fp16, 2 fp32 * scale)."""
        arr = (_lib.GatherJob * len(jobs))()
        first = [0]
        self._keep = []
        for i, j in enumerate(jobs):
            n = j["idx"].numel()
            assert j["idx"].dtype == torch.int32 and j["src0"].dtype == torch.float32 and j["src0"].is_contiguous()
            arr[i].src0 = j["src0"].data_ptr()
            arr[i].src1 = j["src1"].data_ptr() if j.get("src1") is not None else None
            arr[i].idx = j["idx"].data_ptr()
            arr[i].out = j["out"].data_ptr() if isinstance(j["out"], torch.Tensor) else int(j["out"])
            arr[i].n, arr[i].n0, arr[i].mode = n, int(j.get("n0", 0)), int(j["mode"])
            first.append(first[-1] + -(-n // 2048))
            self._keep.append((j["src0"], j.get("src1"), j["idx"], j["out"]))
        self.jobs_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self.first = torch.tensor(first, dtype=torch.int32, device=device)
        self.n_jobs, self.n_blocks = len(jobs), first[-1]

    def launch(self, scale: Optional[torch.Tensor] = None, out_base: Optional[torch.Tensor] = None):
        _count()
        _lib.check(_lib.lib().unet3d_gather_multi(self.jobs_dev.data_ptr(), self.first.data_ptr(), self.n_jobs, self.n_blocks,
                                                  _ptr(scale), _ptr(out_base), _stream()), "unet3d_gather_multi")


class UnpackTable:
    """Device-resident job table of unet3d_dw_unpack: every layer's dw[k3][Kp][Np] accumulator -> parameter layout."""

    def __init__(self, jobs: Sequence[dict], device):
        """jobs: dicts with dw (fp32 tensor), rowmap (int32 tensor [Kp]), out (int byte offset added to `out_base` at
        launch), k3, Kp, Np, ncols, col_stride."""
        arr = (_lib.UnpackJob * len(jobs))()
        first = [0]
        self._keep = []
        for i, j in enumerate(jobs):
            assert j["dw"].dtype == torch.float32 and j["rowmap"].dtype == torch.int32 and j["Np"] % 8 == 0
            assert j["dw"].data_ptr() % 16 == 0 and j["rowmap"].numel() == j["Kp"]
            arr[i].dw, arr[i].rowmap, arr[i].out = j["dw"].data_ptr(), j["rowmap"].data_ptr(), int(j["out"])
            arr[i].k3, arr[i].Kp, arr[i].Np, arr[i].Ncols = j["k3"], j["Kp"], j["Np"], j["ncols"]
            arr[i].col_stride = j["col_stride"]
            first.append(first[-1] + -(-(j["k3"] * j["Kp"] * (j["Np"] // 8)) // 256))
            self._keep.append((j["dw"], j["rowmap"]))
        self.jobs_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self.first = torch.tensor(first, dtype=torch.int32, device=device)
        self.n_jobs, self.n_blocks = len(jobs), first[-1]

    def launch(self, scale: Optional[torch.Tensor] = None, out_base: Optional[torch.Tensor] = None):
        _count()
        _lib.check(_lib.lib().unet3d_dw_unpack(self.jobs_dev.data_ptr(), self.first.data_ptr(), self.n_jobs, self.n_blocks,
                                               _ptr(scale), _ptr(out_base), _stream()), "unet3d_dw_unpack")


class DeviceWgradPlan:
    def __init__(self, plan: P.WgradPlan, device):
        self.plan = plan
        dev_key = str(torch.device(device))
        cache = plan.__dict__.setdefault("_dev", {})
        if dev_key not in cache:
            gidx = torch.from_numpy(plan.gidx).to(device)
            cache[dev_key] = (torch.from_numpy(plan.tab).to(device), gidx, gidx.to(torch.int32),
                              torch.from_numpy(plan.unpack["rowmap"]).to(device))
        self.tab, self.gidx, self.gidx32, self.rowmap = cache[dev_key]


def wgrad_gemm(dp: DeviceWgradPlan, xs: Sequence[torch.Tensor], dy: torch.Tensor, dw: torch.Tensor,
               grid: Tuple[int, int, int, int]):
    pl = dp.plan
    a = _lib.WgradArgs()
    srcs = ([(xs[ti], par, P.WT + 2, P.HT + 2, 8 * pl.wx) for (ti, par) in pl.x_maps] +
            [(dy, par, P.WT, P.HT, 8 * pl.wy) for par in pl.y_maps])
    a.n_src = len(srcs)
    for i, (t, par, bw, bh, bc) in enumerate(srcs):
        a.src[i] = make_src(t, par)
        a.box_w[i], a.box_h[i], a.box_c[i] = bw, bh, bc
    a.tab, a.dw, a.err = dp.tab.data_ptr(), dw.data_ptr(), err_word(dw.device).data_ptr()
    a.N, a.D, a.H, a.W = grid
    _count()
    split = limited_split(pl.n_jobs, pl.split, SM_LIMIT)
    a.n_jobs, a.job_stride, a.split = pl.n_jobs, pl.job_stride, split
    a.x_f16 = _f16(xs[0])
    assert all(x.dtype == dy.dtype for x in xs)        # one MMA cannot mix fp16 and bf16 operands
    assert dw.dtype == torch.float32 and dw.numel() >= pl.dw_numel
    with _Timed("wgrad_gemm_kernel", pl.flops_per_voxel * grid[0] * grid[1] * grid[2] * grid[3], 0.0,
                lambda: f"{pl.kind} x{[tuple(x.shape[1:]) for x in xs]} dy{tuple(dy.shape[1:])} grid{tuple(grid)} jobs{pl.n_jobs} "
                        f"split{split} wx{pl.wx} wy{pl.wy} dt{[j['dt'] for j in pl.jobs][:3]} ent{[len(j['units']) for j in pl.jobs][:4]}"):
        _lib.check(_lib.lib().unet3d_wgrad_gemm(C.byref(a), _stream()), "unet3d_wgrad_gemm")


# ---- memory-bound kernels ---------------------------------------------------------------------------
def in_finalize(stats: torch.Tensor, drop: Optional[torch.Tensor], table: torch.Tensor, count: int, eps: float = 1e-5):
    nc = stats.shape[0] * stats.shape[1]
    _count()
    _lib.check(_lib.lib().unet3d_in_finalize(stats.data_ptr(), _ptr(drop), table.data_ptr(), nc, float(count), eps,
                                             _stream()), "unet3d_in_finalize")


def in_apply(y: torch.Tensor, skip: Optional[torch.Tensor], out: torch.Tensor, table: torch.Tensor,
             shift: Optional[torch.Tensor] = None):
    n, d, h, w, cp = y.shape
    _count()
    assert out.dtype == y.dtype and (skip is None or skip.dtype == y.dtype)
    with _Timed("in_apply", 0.0, _alg_numel(y) * 2.0 * (3 if skip is not None else 2), f"{tuple(y.shape)} skip{int(skip is not None)}"):
        _lib.check(_lib.lib().unet3d_in_apply(y.data_ptr(), _ptr(skip), out.data_ptr(), table.data_ptr(), _ptr(shift), n, d * h * w, cp,
                                              _f16(y), _stream()), "unet3d_in_apply")


def in_bwd_reduce(dout, dout2, out, y, g, table, sums, shift=None):
    n, d, h, w, cp = y.shape
    _count()
    assert dout.dtype == y.dtype and (g is None or g.dtype == y.dtype) and (out is None or out.dtype == y.dtype)
    with _Timed("in_bwd_reduce", 0.0, _alg_numel(y) * 2.0 * (2 + (g is not None) + (dout2 is not None) + (out is not None)),
                f"{tuple(y.shape)} d2{int(dout2 is not None)} out{int(out is not None)} g{int(g is not None)}"):
        _lib.check(_lib.lib().unet3d_in_bwd_reduce(dout.data_ptr(), _ptr(dout2), _ptr(out), y.data_ptr(), _ptr(g),
                                                   table.data_ptr(), _ptr(shift), sums.data_ptr(), n, d * h * w, cp, _f16(y), _stream()),
                   "unet3d_in_bwd_reduce")


def in_bwd_apply(g, y, dy, table, sums, dsum=None, zero_last=False, coef=None, g_is_dout=False):
    """g_is_dout: `g` is the upstream gradient of a norm without residual input (pass 1 ran with g=None): the activation
    gradient is recomputed here from the sign of the normalised value instead of being stored and re-read."""
    n, d, h, w, cp = y.shape
    _count()
    assert g.dtype == y.dtype and dy.dtype == y.dtype and not (g_is_dout and coef is not None)
    with _Timed("in_bwd_apply", 0.0, _alg_numel(y) * 2.0 * 3, f"{tuple(y.shape)}"):
        _lib.check(_lib.lib().unet3d_in_bwd_apply(g.data_ptr(), y.data_ptr(), dy.data_ptr(), table.data_ptr(), sums.data_ptr(),
                                                  _ptr(coef), _ptr(dsum), n, d, h, w, cp, int(zero_last) | (int(g_is_dout) << 1),
                                                  _f16(y), _stream()),
                   "unet3d_in_bwd_apply")


def in_bwd_small(dout, dout2, out, y, g, dy, table, sums):
    """Both backward passes of an InstanceNorm in one launch (small per-sample slices); `sums` is written."""
    n, d, h, w, cp = y.shape
    _count()
    assert dout.dtype == y.dtype and dy.dtype == y.dtype and (g is None or g.dtype == y.dtype)
    assert g is not None or (out is None and dout2 is None)
    with _Timed("in_bwd_small", 0.0, _alg_numel(y) * 2.0 * (3 + (dout2 is not None) + (out is not None) + (g is not None)),
                f"{tuple(y.shape)} d2{int(dout2 is not None)} out{int(out is not None)} g{int(g is not None)}"):
        _lib.check(_lib.lib().unet3d_in_bwd_small(dout.data_ptr(), _ptr(dout2), _ptr(out), y.data_ptr(), _ptr(g), dy.data_ptr(),
                                                  table.data_ptr(), sums.data_ptr(), n, d * h * w, cp, _f16(y), _stream()),
                   "unet3d_in_bwd_small")


def channel_sum(x: torch.Tensor, dsum: torch.Tensor):
    n, d, h, w, cp = x.shape
    _count()
    with _Timed("channel_sum", 0.0, x.numel() * 2.0):
        _lib.check(_lib.lib().unet3d_channel_sum(x.data_ptr(), dsum.data_ptr(), n * d * h * w, cp, _stream()),
                   "unet3d_channel_sum")


def att_gate_fwd(xs: torch.Tensor, z: torch.Tensor, out: torch.Tensor):
    assert xs.dtype in ACT_DTYPES and z.dtype == xs.dtype and out.dtype == xs.dtype and xs.shape == z.shape == out.shape
    _count()
    with _Timed("att_gate_fwd", 0.0, xs.numel() * 6.0):
        _lib.check(_lib.lib().unet3d_att_gate_fwd(xs.data_ptr(), z.data_ptr(), out.data_ptr(), xs.numel(), _f16(xs), _stream()),
                   "unet3d_att_gate_fwd")


def att_gate_bwd(dout, xs, z, dxs, dz, sums):
    n, d, h, w, cp = xs.shape
    assert all(t.dtype == xs.dtype and t.shape == xs.shape for t in (dout, z, dxs, dz))
    assert sums.dtype == torch.float64 and sums.numel() == 2 * cp
    _count()
    with _Timed("att_gate_bwd", 0.0, xs.numel() * 10.0):
        _lib.check(_lib.lib().unet3d_att_gate_bwd(dout.data_ptr(), xs.data_ptr(), z.data_ptr(), dxs.data_ptr(), dz.data_ptr(),
                                                  sums.data_ptr(), n * d * h * w, cp, _f16(xs), _stream()), "unet3d_att_gate_bwd")


def att_mid_bwd(df, f, dxs, dpre, t, ssum):
    n, d, h, w, cp = f.shape
    assert all(x.dtype == f.dtype and x.shape == f.shape for x in (df, dxs, dpre, t))
    assert ssum.dtype == torch.float64 and ssum.numel() == cp
    _count()
    with _Timed("att_mid_bwd", 0.0, f.numel() * 10.0):
        _lib.check(_lib.lib().unet3d_att_mid_bwd(df.data_ptr(), f.data_ptr(), dxs.data_ptr(), dpre.data_ptr(), t.data_ptr(),
                                                 ssum.data_ptr(), n * d * h * w, cp, _f16(f), _stream()), "unet3d_att_mid_bwd")


def maxpool_fwd(x: torch.Tensor, out: torch.Tensor, code: torch.Tensor):
    """nn.MaxPool3d(2, 2) on (N, D, H, W, Cp); code (uint8, shape of out) = kd*4 + kh*2 + kw of the winner."""
    n, d, h, w, cp = x.shape
    assert x.dtype in ACT_DTYPES and out.dtype == x.dtype and code.dtype == torch.uint8
    assert tuple(out.shape) == (n, d // 2, h // 2, w // 2, cp) and code.shape == out.shape
    _count()
    with _Timed("maxpool_fwd", 0.0, x.numel() * 2.0 + out.numel() * 3.0):
        _lib.check(_lib.lib().unet3d_maxpool3d_fwd(x.data_ptr(), out.data_ptr(), code.data_ptr(), n, d, h, w, cp, _f16(x),
                                                   _stream()), "unet3d_maxpool3d_fwd")


def maxpool_bwd(dout: torch.Tensor, code: torch.Tensor, dx: torch.Tensor):
    n, d, h, w, cp = dx.shape
    assert dout.dtype == dx.dtype and code.dtype == torch.uint8 and code.shape == dout.shape
    _count()
    with _Timed("maxpool_bwd", 0.0, dx.numel() * 2.0 + dout.numel() * 3.0):
        _lib.check(_lib.lib().unet3d_maxpool3d_bwd(dout.data_ptr(), code.data_ptr(), dx.data_ptr(), n, d, h, w, cp, _stream()),
                   "unet3d_maxpool3d_bwd")


def maxpool_flat_index(code: torch.Tensor, c: int) -> torch.Tensor:
    """The int64 indices ``F.max_pool3d(..., return_indices=True)`` returns, (N, C, D/2, H/2, W/2), from the codes."""
    n, do, ho, wo, _ = code.shape
    k = code[..., :c].long()
    d = torch.arange(do, device=code.device).view(1, do, 1, 1, 1)
    h = torch.arange(ho, device=code.device).view(1, 1, ho, 1, 1)
    w = torch.arange(wo, device=code.device).view(1, 1, 1, wo, 1)
    flat = ((2 * d + (k >> 2)) * (2 * ho) + 2 * h + ((k >> 1) & 1)) * (2 * wo) + 2 * w + (k & 1)
    return flat.permute(0, 4, 1, 2, 3).contiguous()


def stem_fwd(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, out: torch.Tensor):
    """x fp32 (N, Cin, D, H, W) contiguous; w fp32 [Cin][27][Cp]; b fp32 [Cp]; out 16-bit (N, D, H, W, Cp)."""
    n, d, h, ww, cp = out.shape
    cin = x.shape[1]
    assert x.is_contiguous() and w.numel() == cin * 27 * cp
    _count()
    with _Timed("stem_fwd", 0.0, x.numel() * 4.0 + _alg_numel(out) * 2.0):
        _lib.check(_lib.lib().unet3d_stem_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, cin, d, h, ww, cp,
                                              _f16(out), _stream()), "unet3d_stem_fwd")


def stem_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor):
    """dw fp32 [Cin][28][Cp] (zeroed by the caller): per input channel its 27 taps and the bias row."""
    n, d, h, ww, cp = dy.shape
    cin = x.shape[1]
    assert x.is_contiguous() and dw.numel() == cin * 28 * cp
    vol = d * h * ww
    for ci in range(cin):
        _count()
        with _Timed("stem_wgrad", 0.0, x.numel() / cin * 4.0 + _alg_numel(dy) * 2.0):
            _lib.check(_lib.lib().unet3d_stem_wgrad(x.data_ptr() + 4 * ci * vol, dy.data_ptr(), dw.data_ptr() + 4 * ci * 28 * cp,
                                                    n, d, h, ww, cin * vol, cp, _f16(dy), _stream()), "unet3d_stem_wgrad")


def head_fwd(a: torch.Tensor, w: torch.Tensor, b: torch.Tensor, logits: torch.Tensor):
    n, d, h, ww, cp = a.shape
    k = logits.shape[1]
    _count()
    with _Timed("head_fwd", 0.0, _alg_numel(a) * 2.0 + logits.numel() * 4.0):
        _lib.check(_lib.lib().unet3d_head_fwd(a.data_ptr(), w.data_ptr(), b.data_ptr(), logits.data_ptr(), k, n, d * h * ww,
                                              cp, _f16(a), _stream()), "unet3d_head_fwd")


def head_bwd(dl: torch.Tensor, a: torch.Tensor, w: torch.Tensor, da: torch.Tensor, dw: torch.Tensor,
             grad_scale: Optional[torch.Tensor] = None):
    n, d, h, ww, cp = a.shape
    k = dl.shape[1]
    _count()
    assert da.dtype == a.dtype
    with _Timed("head_bwd", 0.0, dl.numel() * 4.0 + _alg_numel(a) * 4.0):
        _lib.check(_lib.lib().unet3d_head_bwd(dl.data_ptr(), a.data_ptr(), w.data_ptr(), da.data_ptr(), dw.data_ptr(),
                                              _ptr(grad_scale), k, n, d * h * ww, cp, _f16(a), _stream()), "unet3d_head_bwd")


def loss_fwd(logits, target, sums, gamma):
    n, k = logits.shape[:2]
    v = logits[0, 0].numel()
    _count()
    with _Timed("loss_fwd", 0.0, logits.numel() * 4.0 + target.numel() * 8.0):
        _lib.check(_lib.lib().unet3d_loss_fwd(logits.data_ptr(), target.data_ptr(), sums.data_ptr(), k, n, v, float(gamma),
                                              _stream()), "unet3d_loss_fwd")


def loss_bwd(logits, target, coef, gscale, dlogits, gamma, use_focal):
    n, k = logits.shape[:2]
    v = logits[0, 0].numel()
    _count()
    with _Timed("loss_bwd", 0.0, logits.numel() * 8.0 + target.numel() * 8.0):
        _lib.check(_lib.lib().unet3d_loss_bwd(logits.data_ptr(), target.data_ptr(), coef.data_ptr(), _ptr(gscale),
                                              dlogits.data_ptr(), k, n, v, float(gamma), int(use_focal), _stream()),
                   "unet3d_loss_bwd")


def sw_accumulate(logits, window, acc, origin, shape):
    """acc: int64 (n_slab, K + 1, Xs, Y, Z) fixed-point sums (see csrc/elementwise.cu: sw_accumulate_kernel); shape = the
    (padded) volume extent (X, Y, Z) the window origin refers to, X <= n_slab * Xs."""
    k, px, py, pz = logits.shape[-4:]
    n_slab, k1, Xs, Y, Z = acc.shape
    X = int(shape[0])
    assert acc.dtype == torch.int64 and acc.is_contiguous() and k1 == k + 1 and X <= n_slab * Xs
    assert (Y, Z) == (int(shape[1]), int(shape[2]))
    _count()
    with _Timed("sw_accumulate", 0.0, px * py * pz * (4.0 * k + 16.0 * (k + 1) + (4.0 if window is not None else 0.0))):
        _lib.check(_lib.lib().unet3d_sw_accumulate(logits.data_ptr(), _ptr(window), acc.data_ptr(), k, px, py, pz,
                                                   origin[0], origin[1], origin[2], X, Y, Z, Xs, _stream()),
                   "unet3d_sw_accumulate")


def sw_finalize(slab, labels, probs):
    """slab: int64 (K + 1, ...) sums of one slab; labels uint8 (...) and / or probs fp32 (..., K)."""
    k = slab.shape[0] - 1
    n = slab[0].numel()
    assert slab.dtype == torch.int64 and slab.is_contiguous()
    _count()
    with _Timed("sw_finalize", 0.0, n * (8.0 * (k + 1) + (1.0 if labels is not None else 0.0) + (4.0 * k if probs is not None else 0.0))):
        _lib.check(_lib.lib().unet3d_sw_finalize(slab.data_ptr(), _ptr(labels), _ptr(probs), k, n, _stream()),
                   "unet3d_sw_finalize")


# ------------------------------------------------------------------------------------------------
# case-level resampling (transform.py:32-100 = scipy.ndimage.zoom order 1) -- csrc/resample.cu
# ------------------------------------------------------------------------------------------------
def _zoom_ws(out_shape, device):
    n = _lib.lib().unet3d_zoom_workspace_bytes(*[int(v) for v in out_shape])
    return torch.empty((n + 15) // 16 * 2, dtype=torch.float64, device=device), n


def _c_arr(ctype, values):
    return (ctype * len(values))(*[int(v) for v in values])


@_on_device_of_first_arg
def zoom_linear(src: torch.Tensor, dst: torch.Tensor, norm: Optional[Sequence[Sequence[float]]] = None):
    """dst[x', y', z', c] = zoom(src[..., c]); src / dst are 4-D (x, y, z, channel) VIEWS (any strides) of float32 or
    uint8 CUDA tensors; the output shape is dst's.  norm: per channel (pct_00_5, pct_99_5, mean, std + 1e-8)."""
    assert src.is_cuda and dst.is_cuda and src.dim() == 4 and dst.dim() == 4 and src.shape[3] == dst.shape[3]
    assert src.dtype in (torch.float32, torch.uint8) and dst.dtype in (torch.float32, torch.uint8)
    ws, nbytes = _zoom_ws(dst.shape[:3], src.device)
    nm = None
    if norm is not None:
        nm = (C.c_float * (4 * len(norm)))(*[float(np.float32(v)) for row in norm for v in row])
    vox = float(np.prod(dst.shape[:3])) * dst.shape[3]
    _count(2)
    with _Timed("zoom_linear", 0.0, vox * (src.element_size() * float(np.prod(src.shape[:3])) / float(np.prod(dst.shape[:3]))
                                           + dst.element_size())):
        _lib.check(_lib.lib().unet3d_zoom_linear(
            src.data_ptr(), int(src.dtype == torch.uint8), dst.data_ptr(), int(dst.dtype == torch.uint8), int(src.shape[3]),
            _c_arr(C.c_int, src.shape[:3]), _c_arr(C.c_longlong, src.stride()), _c_arr(C.c_int, dst.shape[:3]),
            _c_arr(C.c_longlong, dst.stride()), nm, ws.data_ptr(), nbytes, _stream()), "unet3d_zoom_linear")


@_on_device_of_first_arg
def zoom_label(src: torch.Tensor, dst: torch.Tensor):
    """Labels with >= 3 classes: per-class one-hot zoom + argmax (transform.py:72-78); 3-D uint8 views."""
    assert src.is_cuda and dst.is_cuda and src.dim() == 3 and dst.dim() == 3
    assert src.dtype == torch.uint8 and dst.dtype == torch.uint8
    ws, nbytes = _zoom_ws(dst.shape, src.device)
    _count(2)
    with _Timed("zoom_label", 0.0, float(src.numel() + dst.numel())):
        _lib.check(_lib.lib().unet3d_zoom_label(
            src.data_ptr(), dst.data_ptr(), _c_arr(C.c_int, src.shape), _c_arr(C.c_longlong, src.stride()),
            _c_arr(C.c_int, dst.shape), _c_arr(C.c_longlong, dst.stride()), ws.data_ptr(), nbytes, _stream()),
            "unet3d_zoom_label")


# ------------------------------------------------------------------------------------------------
# cascade glue (data.regions_crop_case, trainer.cascade_predict_case merge) -- csrc/regions.cu
# ------------------------------------------------------------------------------------------------
@_on_device_of_first_arg
def connected_components(mask: torch.Tensor):
    """6-connected components of a uint8 CUDA volume (X, Y, Z): scipy.ndimage.label's numbering.
    Returns (labels int32 (X, Y, Z): root index or -1, roots int32 [n] sorted, stats int32 [n][8] =
    {voxels, xmin, xmax, ymin, ymax, zmin, zmax, 0})."""
    assert mask.is_cuda and mask.dtype == torch.uint8 and mask.dim() == 3 and mask.is_contiguous()
    X, Y, Z = (int(v) for v in mask.shape)
    labels = torch.empty((X, Y, Z), dtype=torch.int32, device=mask.device)
    is_root = torch.empty((X, Y, Z), dtype=torch.uint8, device=mask.device)
    _count(3)
    with _Timed("ccl_label", 0.0, mask.numel() * (1.0 + 4.0 + 1.0)):
        _lib.check(_lib.lib().unet3d_ccl_label(mask.data_ptr(), labels.data_ptr(), is_root.data_ptr(), X, Y, Z, _stream()),
                   "unet3d_ccl_label")
    roots = torch.nonzero(is_root.view(-1)).view(-1).to(torch.int32)          # sorted = raster order of the first voxels
    n = int(roots.numel())
    init = torch.tensor([0, 2 ** 31 - 1, -1, 2 ** 31 - 1, -1, 2 ** 31 - 1, -1, 0], dtype=torch.int32, device=mask.device)
    stats = init.repeat(max(n, 1), 1).contiguous()
    if n:
        _count()
        with _Timed("ccl_stats", 0.0, mask.numel() * 4.0):
            _lib.check(_lib.lib().unet3d_ccl_stats(labels.data_ptr(), roots.data_ptr(), n, stats.data_ptr(), X, Y, Z,
                                                   _stream()), "unet3d_ccl_stats")
    return labels, roots, stats[:n]


@_on_device_of_first_arg
def region_accumulate(pred: torch.Tensor, result: torch.Tensor, count: torch.Tensor, src0, dst0, box):
    """result[dst0 + i] += pred[src0 + i] over `box` voxels; pred float32 (rx, ry, rz, K) view, result float64
    (X, Y, Z, K) contiguous, count int32 (X, Y, Z) contiguous."""
    assert pred.dtype == torch.float32 and result.dtype == torch.float64 and count.dtype == torch.int32
    assert result.is_contiguous() and count.is_contiguous() and pred.stride(3) == 1 and pred.shape[3] == result.shape[3]
    if min(box) < 1:
        return
    view = pred[src0[0]:src0[0] + box[0], src0[1]:src0[1] + box[1], src0[2]:src0[2] + box[2]]
    _count()
    with _Timed("region_accumulate", 0.0, float(np.prod(box)) * (pred.shape[3] * 20.0 + 8.0)):
        _lib.check(_lib.lib().unet3d_region_accumulate(
            view.data_ptr(), result.data_ptr(), count.data_ptr(), int(pred.shape[3]), _c_arr(C.c_int, box),
            _c_arr(C.c_longlong, view.stride()[:3]), _c_arr(C.c_int, dst0), int(result.shape[1]), int(result.shape[2]),
            _stream()), "unet3d_region_accumulate")


@_on_device_of_first_arg
def merge_finalize(result: torch.Tensor, count: torch.Tensor) -> torch.Tensor:
    labels = torch.empty(count.shape, dtype=torch.uint8, device=count.device)
    _count()
    with _Timed("merge_finalize", 0.0, count.numel() * (8.0 * result.shape[3] + 5.0)):
        _lib.check(_lib.lib().unet3d_merge_finalize(result.data_ptr(), count.data_ptr(), labels.data_ptr(),
                                                    int(result.shape[3]), count.numel(), _stream()), "unet3d_merge_finalize")
    return labels


@_on_device_of_first_arg
def overlap_counts(pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """int64 (3, 256): per label value c {|pred == c and label == c|, |pred == c|, |label == c|} of two uint8 CUDA volumes."""
    assert pred.is_cuda and label.is_cuda and pred.dtype == torch.uint8 and label.dtype == torch.uint8
    assert pred.shape == label.shape
    pred, label = pred.contiguous(), label.contiguous()
    counts = torch.zeros(3, 256, dtype=torch.int64, device=pred.device)
    _count()
    with _Timed("overlap_counts", 0.0, 2.0 * pred.numel()):
        _lib.check(_lib.lib().unet3d_overlap_counts(pred.data_ptr(), label.data_ptr(), pred.numel(), counts.data_ptr(),
                                                    _stream()), "unet3d_overlap_counts")
    return counts

"""Host-side plans for the tcgen05 shifted-GEMM kernels (csrc/conv_gemm.cu, csrc/wgrad_gemm.cu).

A plan turns one layer call of the reference network (nn.Conv3d / nn.ConvTranspose3d forward or one
of their gradients; network.py:312-314,394-403,541-547) into

  * the A-source views (whole tensors, concat halves, stride-2 parity sub-lattices),
  * the int32 device table the kernel walks (channel groups, tap shifts, active-tap masks, ...),
  * a gather index that packs the fp32 PyTorch-layout parameter into the bf16 tile stream the
    kernel's weight ring consumes (one index_select per layer per step).

Everything here is integer bookkeeping done once per (layer, shape); it is plain numpy so that the
CPU test-suite can check it against torch.nn.functional without a GPU (tests/emulate.py executes a
plan with torch CPU ops).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

HT, WT = 16, 8                 # output tile (rows of one UMMA): 16 x 8 voxels
CHUNK_PITCH = 2944             # weight-gradient x bricks: one (plane, 8-channel chunk) box, padded to 128 B


def plane_pitch(g: int) -> int:
    """conv_gemm A slab: bytes of one plane of a channel group of g chunks (18 x 10 rows of 16 g bytes, swizzled),
    padded to the 1 KiB swizzle-atom alignment (csrc/conv_gemm.cuh: cg_plane_pitch)."""
    return -(-(18 * 10 * 16 * g) // 1024) * 1024
W_STAGES = 16
FUSE_KD = True                 # fuse the three d-taps of a (kh,kw) into one wider UMMA where N <= 64
SMEM_LIMIT = 227 * 1024


_PAD64 = os.environ.get("U3D_PAD64", "0") == "1"


def pad_channels(c: int) -> int:
    """Physical channel count of a 16-bit NDHWC activation: a multiple of 16 (UMMA K step).

    Tuning knob U3D_PAD64=1 (off): wide tensors go up to the next multiple of 64 when that costs at most 7 %
    (240 -> 256, 480 -> 512), so that the conv kernels can use 32- / 64-channel groups where 240 = 15 x 16 forces
    16-channel ones.  Measured at cfg-2 (tools/bench_layers.py, round 2): the strided 240 -> 480 layers gain (forward
    130 -> 95 us, transposed data gradient 126 -> 92 us), the 8^3-grid weight gradients too (50 -> 40 us x 9), but the
    level-4 forward / data gradient lose (58 -> 61, 54 -> 56 us x 9): about -0.1 ms per 18.7 ms step for 6.7 % more
    activation memory on the deep levels -- not taken."""
    p16 = (c + 15) // 16 * 16
    p64 = (c + 63) // 64 * 64
    if _PAD64 and c > 128 and p64 * 100 <= c * 107:
        return p64
    return p16


def a_slabs(dt: int, g: int) -> int:
    """A slabs (channel groups) in flight: 2.  The kernel supports up to 4 (env U3D_A_STAGES, for sweeps); measured at
    cfg-2 (gpurun_out, sweep with 2/3/4 slabs) a third slab never helps and costs weight-ring space: the A loads are
    not what the MMA warp waits for."""
    forced = os.environ.get("U3D_A_STAGES")
    slab = (dt + 2) * plane_pitch(g)
    n = int(forced) if forced else 2
    return max(2, min(4, n, (SMEM_LIMIT - 2048 - 16384) // slab))


def weight_ring(dt: int, g: int, nblk: int, fuse: int, n_taps: int = 27) -> Tuple[int, int]:
    """(taps per stage, stages) of the weight ring: batch small tiles up to ~28 KB per stage (one barrier round trip
    and one issue dispatch per batch), ring of ~56 KB when the A slabs leave that much.  Measured (gpurun_out/
    layers6.log): smaller batches in a deeper ring (16 KB x 7..11 stages) are SLOWER on every layer (L0 +6 %, L4 +30 %):
    the per-batch cost on the single MMA-issuing thread outweighs the extra latency cover."""
    tile = g * fuse * nblk * 16
    taps = max(1, -(-n_taps // fuse))
    avail = SMEM_LIMIT - 2048 - a_slabs(dt, g) * (dt + 2) * plane_pitch(g)
    wt = max(1, min(taps, 16, 28672 // tile))
    while wt > 1 and avail // (wt * tile) < 2:
        wt -= 1
    ring_bytes = int(os.environ.get("U3D_W_RING", "57344"))
    stages = min(W_STAGES, max(2, ring_bytes // (wt * tile)), avail // (wt * tile))
    return wt, stages


def conv_smem_bytes(dt: int, g: int, nblk: int, fuse: int = 1, n_taps: int = 27) -> int:
    """Shared memory of one conv_gemm CTA; > SMEM_LIMIT when not even a 2-stage weight ring fits."""
    wt, stages = weight_ring(dt, g, nblk, fuse, n_taps)
    if stages < 2:
        return SMEM_LIMIT + 1
    return 2048 + a_slabs(dt, g) * (dt + 2) * plane_pitch(g) + stages * wt * g * fuse * nblk * 16


def choose_nblk(cp_out: int) -> Tuple[int, int]:
    best = None
    for nblk in (128, 96, 64, 32):
        n = -(-cp_out // nblk)
        waste = n * nblk - cp_out
        if best is None or waste < best[0]:
            best = (waste, nblk, n)
    return best[1], best[2]


def choose_dt_g(nblk: int, chunk_counts: Sequence[int], depth: int, fuse: int = 1) -> Tuple[int, int]:
    """Planes per segment and chunks per channel group: biggest Dt (halo amortisation), then biggest G."""
    gs = [g for g in (4, 8, 2) if all(c % g == 0 for c in chunk_counts)]
    if not gs:
        raise ValueError(f"channel chunk counts {chunk_counts} need an even common divisor")
    dt_max = max(1, min(256 // nblk, 8, depth))
    for dt in (8, 4, 2, 1):                    # the kernel's MMA issue code is unrolled for these
        if dt > dt_max:
            continue
        for g in gs:
            if conv_smem_bytes(dt, g, nblk, fuse) <= SMEM_LIMIT:
                return dt, g
    raise ValueError("no (Dt, G) fits shared memory")


NUM_SMS = 148
_L2_LATENCY_CLK = 1.0               # loaded L2 -> SM round trip the weight ring has to cover (~2 us)
_L2_BYTES_PER_CLK = 3700.0          # whole-chip L2 -> SM bandwidth the cost model assumes (~7 TB/s at 1.9 GHz)


def _mma_cycles(n: int) -> float:
    """Sustained cycles of one M=128, K=16 tcgen05.mma of width n with both operands in shared memory, measured back
    to back on a B200 (tools/probe/mma_rate.cu): 47 / 51 / 56 / 64 / 96 / 128 for n = 32 / 64 / 96 / 128 / 192 / 256 --
    the tensor floor n/2 above n = 128, a ~45-cycle operand-read floor below it, independent of layout and swizzle."""
    return max(n / 2.0, 42.0 + n / 7.0)


def choose_config(out_cp: int, chunk_counts: Sequence[int], grid, taps_per_cg: float, n_out_par: int, can_fuse: bool):
    """Pick (nblk, Dt, G, nbuf, fuse) for one conv launch with a small analytic model of the kernel:
    per work item  t = max(MMA cycles, L2 bytes / share of L2 bandwidth) [+ exposed epilogue if TMEM is single
    buffered], total = waves * t.  Wide N blocks and few planes starve small grids; narrow ones re-stream weights."""
    N_, D_, H_, W_ = grid
    best = None
    forced = os.environ.get("U3D_CONV_CFG")         # tuning sweeps only: "nblk,dt,g,nbuf,fuse"
    if forced:
        nblk, dt, g, nbuf, fuse = (int(v) for v in forced.split(","))
        if not can_fuse:
            fuse = 1
        return nblk, dt, g, nbuf, fuse
    gs = [g for g in (4, 8, 2) if all(c % g == 0 for c in chunk_counts)]
    if not gs:
        raise ValueError(f"channel chunk counts {chunk_counts} need an even common divisor")
    n_chunks = sum(chunk_counts)
    for nblk in (128, 96, 64, 32):
        n_nb = -(-out_cp // nblk) * n_out_par
        waste = (-(-out_cp // nblk) * nblk) / float(out_cp)
        for nbuf in (2, 1):
            for dt in (8, 4, 2, 1):
                if dt * nblk * nbuf > 512 or dt > max(1, D_):
                    continue
                fuse = 3 if (can_fuse and nblk <= 64 and FUSE_KD) else 1
                for g in gs:
                    if conv_smem_bytes(dt, g, nblk, fuse) > SMEM_LIMIT:
                        continue
                    n_cg = n_chunks // g
                    g2 = g // 2
                    tiles = N_ * (-(-D_ // dt)) * (-(-H_ // HT)) * (-(-W_ // WT))
                    items = tiles * n_nb
                    if fuse == 3:
                        per_tap = (max(dt - 2, 0) * _mma_cycles(3 * nblk) + min(2, dt) * _mma_cycles(2 * nblk if dt >= 2 else nblk)
                                   + 2 * _mma_cycles(nblk)) * g2
                        mma = n_cg * (taps_per_cg / 3.0) * per_tap
                    else:
                        mma = n_cg * taps_per_cg * dt * g2 * _mma_cycles(nblk)
                    if not (fuse == 3 and nblk in (32, 64)):
                        mma *= 1.3          # generic issue path: per-tap table / mask bookkeeping on the issuing thread
                    a_bytes = n_cg * (dt + 2) * g * 2880 * (2.0 if g == 1 else 1.0)     # rows of 16 g >= 32 bytes: whole sectors
                    w_bytes = n_cg * taps_per_cg * g * nblk * 16
                    active = min(items, NUM_SMS)
                    bw = min(48.0, _L2_BYTES_PER_CLK / active)
                    epi = dt * (nblk / 32.0) * 500.0
                    wt, stages = weight_ring(dt, g, nblk, fuse)
                    ring_rate = stages * wt * g * fuse * nblk * 16 / _L2_LATENCY_CLK     # bytes in flight per L2 round trip
                    t_item = (max(mma, (a_bytes + w_bytes) / bw, w_bytes / min(bw, ring_rate))
                              + (epi if nbuf == 1 else 0.0) + 3000.0)
                    waves = max(1.0, items / float(NUM_SMS))
                    if waves < 6:
                        waves = float(-(-items // NUM_SMS))
                    cost = waves * t_item * (1.0 + 0.0 * waste)
                    if best is None or cost < best[0]:
                        best = (cost, nblk, dt, g, nbuf, fuse)
    if best is None:
        raise ValueError("no conv configuration fits shared memory / TMEM")
    return best[1:]


# per-dimension (parity, shift) -> kernel index for a stride-2 k3 p1 conv read through parity views:
# in = 2*o + k - 1;  parity-0 view index q = o holds in = 2q (k = 1), brick shift 1 (offset 0);
# parity-1 view: in = 2q + 1 -> k = 0 at q = o - 1 (shift 0), k = 2 at q = o (shift 1).
_S2_K = {(0, 1): 1, (1, 0): 0, (1, 1): 2}
# transposed: out[2q + p] += in[q + o] * W[k], k = p + 1 - 2 o; brick shift = o + 1
_TR_K = {(0, 1): 1, (1, 1): 2, (1, 2): 0}


@dataclass
class ConvPlan:
    kind: str                       # conv_fwd | conv_dgrad | convT_fwd | convT_dgrad
    ks: int
    stride: int
    pattern: str                    # direct | transposed
    in_C: List[int]                 # real channels of each A input tensor (concat order)
    in_Cp: List[int]
    out_C: List[int]                # real channels of each output tensor
    out_Cp: List[int]
    maps: List[Tuple[int, Optional[Tuple[int, int, int]]]]   # (input tensor index, parity or None)
    G: int
    Dt: int
    nblk: int
    cg_map: List[int]
    cg_ch: List[int]
    shifts: List[Tuple[int, int, int]]
    nb_sel: List[int]
    nb_coff: List[int]
    nb_ooff: List[Tuple[int, int, int]]
    nb_real0: List[int]             # first real N-channel (concatenated over outputs) of the block
    masks: np.ndarray               # [n_nblk, n_cg] uint32
    wbase: List[int]
    tab: np.ndarray                 # int32 table for the kernel
    widx: np.ndarray                # int64 gather index into cat(W.flatten(), [0])
    omul: int
    flops_per_voxel: float = 0.0    # algorithmic FLOPs per tile-grid voxel (unpadded channels)
    wT: int = 1                     # taps per weight-ring stage
    w_stages: int = 6               # weight-ring depth
    a_stages: int = 2               # A slabs in flight
    dense: bool = False             # fused, 9 taps in (kh,kw) order, all active everywhere: the kernel's unrolled path
    nbuf: int = 2                   # TMEM accumulator buffers (1: Dt * nblk <= 512, epilogue not overlapped)
    n_tiles_w: int = 0              # number of weight tiles
    fuse_kd: bool = False           # one weight tile = the 3 d-taps of a (kh,kw), rows ordered sd = 2,1,0

    @property
    def n_nblk(self) -> int:
        return len(self.nb_sel)

    @property
    def n_cg(self) -> int:
        return len(self.cg_map)


def _kidx_and_valid(plan_kind: str, ks: int, stride: int, pattern: str, shift, parity_in, parity_out):
    """Kernel tap (kd,kh,kw) used by brick shift `shift` for the given source parity (direct stride 2)
    or output parity (transposed); None if that combination contributes nothing."""
    k = []
    for d in range(3):
        s = shift[d]
        if ks == 1:
            if s != 1:
                return None
            if pattern == "direct" and stride == 2 and parity_in[d] != 0:
                return None
            if pattern == "transposed" and parity_out[d] != 0:
                return None
            k.append(0)
        elif pattern == "direct" and stride == 1:
            k.append(s)
        elif pattern == "direct":
            kk = _S2_K.get((parity_in[d], s))
            if kk is None:
                return None
            k.append(kk)
        else:
            kk = _TR_K.get((parity_out[d], s))
            if kk is None:
                return None
            k.append(kk)
    return tuple(k)


_PLAN_CACHE: Dict[tuple, object] = {}
_TUNING_ENV = ("U3D_CONV_CFG", "U3D_A_STAGES", "U3D_WG_WAVES", "U3D_WG_NOSW", "U3D_WG_FORCESW", "U3D_WG_PAIR", "U3D_W_RING",
               "U3D_WG_WSHIFT")


def make_conv_plan(kind: str, ks: int, stride: int, in_C: Sequence[int], out_C: Sequence[int], depth: int,
                   grid: Optional[Tuple[int, int, int, int]] = None, skip_k1: bool = False) -> ConvPlan:
    """Memoised front end of _make_conv_plan: a network repeats the same layer shape many times (the nine level-4
    blocks of the default net share two plans), and building the gather index of a 480x480x27 weight costs ~0.1 s of
    numpy work.  Plans are immutable after construction."""
    key = ("conv", kind, ks, stride, tuple(in_C), tuple(out_C), depth, None if grid is None else tuple(grid), skip_k1,
           tuple(os.environ.get(k) for k in _TUNING_ENV))
    if key not in _PLAN_CACHE:
        _PLAN_CACHE[key] = _make_conv_plan(kind, ks, stride, in_C, out_C, depth, grid, skip_k1)
    return _PLAN_CACHE[key]


def _make_conv_plan(kind: str, ks: int, stride: int, in_C: Sequence[int], out_C: Sequence[int], depth: int,
                    grid: Optional[Tuple[int, int, int, int]] = None, skip_k1: bool = False) -> ConvPlan:
    """kind:
         conv_fwd    Conv3d forward (weight (Cout, Cin, k,k,k)); inputs may be a concat (len(in_C) > 1)
         conv_dgrad  its data gradient; outputs may be a concat split (len(out_C) > 1)
         convT_fwd   ConvTranspose3d(k3,s2,p1) forward + zero pad plane (weight (Cin, Cout, k,k,k))
         convT_dgrad its data gradient
       in_C / out_C are the REAL channel counts of the A-side / output-side tensors of this call
       (for a gradient, in_C is the channel count of dy).  `depth` = D extent of the tile grid.

       skip_k1 (conv_dgrad only): the data gradient of a whole residual block input in ONE launch,
         dx = dgrad_conv1(dy1) + dgrad_skip_conv(g2)        (network.py:405-416: both convs read the block input)
       A sources = [dy1, g2] (same channel count); the first uses the k3 weight with every tap, the second the k1
       skip weight with the centre tap only.  The packed stream is gathered from cat(W_conv1.flatten(), W_skip.flatten())."""
    in_C, out_C = list(in_C), list(out_C)
    if skip_k1:
        assert kind == "conv_dgrad" and ks == 3 and len(in_C) == 2 and in_C[0] == in_C[1]
    in_Cp = [pad_channels(c) for c in in_C]
    out_Cp = [pad_channels(c) for c in out_C]
    if kind == "conv_fwd":
        pattern = "direct"
    elif kind == "conv_dgrad":
        pattern = "direct" if stride == 1 else "transposed"
    elif kind == "convT_fwd":
        pattern, stride, ks = "transposed", 2, 3
    elif kind == "convT_dgrad":
        pattern, stride, ks = "direct", 2, 3
    else:
        raise ValueError(kind)
    if (pattern == "transposed" or stride == 2) and ((len(in_C) != 1 and not skip_k1) or len(out_C) != 1):
        raise ValueError("strided / transposed convs take one input and one output")

    # ---- A maps and channel groups
    parities = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    if pattern == "direct" and stride == 2:
        par_list = parities if ks == 3 else [(0, 0, 0)]
        maps = [(0, p) for p in par_list]
    else:
        maps = [(i, None) for i in range(len(in_C))]
    chunk_counts = [in_Cp[m[0]] // 8 for m in maps]

    # ---- configuration (N block width, planes per segment, chunks per channel group, TMEM buffering, d-tap fusion)
    if grid is None:
        grid = (2, max(1, depth), 64, 64)
    n_shift = 1 if ks == 1 else (27 if (pattern == "direct" and stride == 1) else 8)
    if ks == 1:
        taps_avg = 1.0
    elif pattern == "direct" and stride == 1:
        taps_avg = 27.0
    else:
        taps_avg = 27.0 / 8.0
    can_fuse = pattern == "direct" and stride == 1 and ks == 3
    n_out_par = 8 if (pattern == "transposed" and ks == 3) else 1
    nblk, Dt, G, nbuf, fuse = choose_config(max(out_Cp), chunk_counts, grid, taps_avg, n_out_par, can_fuse)
    if len(out_Cp) > 1 and any(cp % nblk and cp > nblk for cp in out_Cp):
        nblk = 32
        Dt, G = choose_dt_g(nblk, chunk_counts, depth, 3 if can_fuse and FUSE_KD else 1)
        nbuf, fuse = 2, (3 if can_fuse and FUSE_KD else 1)
    fuse_kd = fuse == 3

    # ---- N blocks
    nb_sel, nb_coff, nb_ooff, nb_real0 = [], [], [], []
    if pattern == "transposed":
        n = -(-out_Cp[0] // nblk)
        out_par = parities if ks == 3 else [(0, 0, 0)]
        for p in out_par:
            for j in range(n):
                nb_sel.append(0); nb_coff.append(j * nblk); nb_ooff.append(p); nb_real0.append(j * nblk)
    else:
        real0 = 0
        for t, cp in enumerate(out_Cp):
            for j in range(-(-cp // nblk)):
                nb_sel.append(t); nb_coff.append(j * nblk); nb_ooff.append((0, 0, 0)); nb_real0.append(real0 + j * nblk)
            real0 += out_C[t]

    cg_map, cg_ch = [], []
    for mi, cnt in enumerate(chunk_counts):
        for i in range(cnt // G):
            cg_map.append(mi); cg_ch.append(i * G * 8)

    # ---- taps
    if ks == 1:
        shifts = [(1, 1, 1)]
    elif pattern == "direct" and stride == 1:
        shifts = [(a, b, c) for a in range(3) for b in range(3) for c in range(3)]
        if fuse_kd:
            shifts = [(0, b, c) for b in range(3) for c in range(3)]
    elif pattern == "direct":
        shifts = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    else:
        shifts = [(a, b, c) for a in (1, 2) for b in (1, 2) for c in (1, 2)]

    # real channel index of every K slot: concat offset of the source tensor + channel, -1 for padding
    in_off = np.concatenate([[0], np.cumsum(in_C)]).astype(np.int64)
    Ktot, Ntot = int(sum(in_C)), int(sum(out_C))
    skip_woff = 0
    if skip_k1:
        Ktot = in_C[0]
        skip_woff = Ktot * Ntot * ks ** 3          # W_skip starts right after W_conv1 in the concatenated source
    out_real_end = np.cumsum(out_C)

    n_nb, n_cg = len(nb_sel), len(cg_map)
    masks = np.zeros((n_nb, n_cg), dtype=np.uint32)
    wbase, pieces = [], []
    n_tiles = 0
    k3 = ks ** 3
    for nb in range(n_nb):
        wbase.append(n_tiles)
        t_out = nb_sel[nb]
        ncol = nb_real0[nb] + np.arange(nblk, dtype=np.int64)                 # real N channel (global)
        n_hi = int(out_real_end[t_out])
        n_ok = (nb_coff[nb] + np.arange(nblk)) < out_C[t_out]
        ncol = np.where(n_ok & (ncol < n_hi), ncol, -1)
        for cg in range(n_cg):
            ti, par_in = maps[cg_map[cg]]
            ch = cg_ch[cg] + np.arange(G * 8, dtype=np.int64)                 # channel within the source tensor
            krow = np.where(ch < in_C[ti], in_off[ti] + ch, -1)
            is_skip_src = skip_k1 and ti == 1
            if skip_k1:
                krow = np.where(ch < in_C[ti], ch, -1)                        # each source has its own weight tensor
            for t, sh0 in enumerate(shifts):
                sub = []
                for sd in ((2, 1, 0) if fuse_kd else (sh0[0],)):
                    sh = (sd, sh0[1], sh0[2])
                    if is_skip_src:
                        # k1 skip conv: centre tap only; in a fused tile the other two d-taps are structural zeros
                        kk1 = _kidx_and_valid(kind, 1, stride, pattern, sh, par_in or (0, 0, 0), nb_ooff[nb])
                        if kk1 is None:
                            if fuse_kd and (sh0[1], sh0[2]) == (1, 1):
                                sub.append(np.full((G, nblk, 8), -1, np.int64))
                                continue
                            sub = None
                            break
                        K = krow.reshape(G, 1, 8)
                        Nn = ncol.reshape(1, nblk, 1)
                        flat = skip_woff + (K * Ntot + Nn)                      # W_skip[cout=K][cin=N][0]
                        sub.append(np.where((K >= 0) & (Nn >= 0), flat, -1))
                        continue
                    kk = _kidx_and_valid(kind, ks, stride, pattern, sh, par_in or (0, 0, 0), nb_ooff[nb])
                    if kk is None:
                        sub = None
                        break
                    if kind == "conv_dgrad" and pattern == "direct":
                        kk = tuple(ks - 1 - v for v in kk)                        # flipped kernel
                    kflat = (kk[0] * ks + kk[1]) * ks + kk[2]
                    K = krow.reshape(G, 1, 8)
                    Nn = ncol.reshape(1, nblk, 1)
                    if kind == "conv_fwd":            # W[cout=N][cin=K][k]
                        flat = (Nn * Ktot + K) * k3 + kflat
                    elif kind == "conv_dgrad":        # W[cout=K][cin=N][k]
                        flat = (K * Ntot + Nn) * k3 + kflat
                    elif kind == "convT_fwd":         # Wt[cin=K][cout=N][k]
                        flat = (K * Ntot + Nn) * k3 + kflat
                    else:                             # convT_dgrad: Wt[cin=N][cout=K][k]
                        flat = (Nn * Ktot + K) * k3 + kflat
                    sub.append(np.where((K >= 0) & (Nn >= 0), flat, -1))
                if sub is None:
                    continue
                masks[nb, cg] |= np.uint32(1 << t)
                pieces.append(np.concatenate(sub, axis=1).reshape(-1))        # [G][fuse * nblk][8]
                n_tiles += 1
    widx = np.concatenate(pieces) if pieces else np.zeros(0, np.int64)
    if kind in ("conv_fwd", "convT_dgrad"):
        numel = Ntot * Ktot * k3
    else:
        numel = Ktot * Ntot * k3
    if skip_k1:
        numel += Ktot * Ntot
    assert numel < 2 ** 31
    widx = widx.astype(np.int32)                    # -1 = structural zero (channel padding)

    ring = weight_ring(Dt, G, nblk, 3 if fuse_kd else 1, 27 if ks == 3 else 1)
    assert conv_smem_bytes(Dt, G, nblk, 3 if fuse_kd else 1, 27 if ks == 3 else 1) <= SMEM_LIMIT
    tab = np.concatenate([
        np.asarray(cg_map, np.int32), np.asarray(cg_ch, np.int32),
        np.asarray([s[0] | (s[1] << 8) | (s[2] << 16) for s in shifts], np.int32),
        masks.astype(np.int64).astype(np.int32).reshape(-1),
        np.asarray(wbase, np.int32),
        np.asarray([c | (s << 30) for c, s in zip(nb_coff, nb_sel)], np.int32),
        np.asarray([o[0] | (o[1] << 8) | (o[2] << 16) for o in nb_ooff], np.int32),
    ]).astype(np.int32)

    return ConvPlan(kind=kind, ks=ks, stride=stride, pattern=pattern, in_C=in_C, in_Cp=in_Cp, out_C=out_C, out_Cp=out_Cp,
                    maps=maps, G=G, Dt=Dt, nblk=nblk, cg_map=cg_map, cg_ch=cg_ch, shifts=shifts, nb_sel=nb_sel,
                    nb_coff=nb_coff, nb_ooff=nb_ooff, nb_real0=nb_real0, masks=masks, wbase=wbase, tab=tab, widx=widx,
                    omul=2 if pattern == "transposed" else 1, n_tiles_w=n_tiles, fuse_kd=fuse_kd, nbuf=nbuf, wT=ring[0], w_stages=ring[1], a_stages=a_slabs(Dt, G),
                    dense=bool(fuse_kd and len(shifts) == 9 and nblk in (32, 64) and (masks == 0x1ff).all()),
                    flops_per_voxel=(2.0 * in_C[0] * sum(out_C) * (k3 + 1)) if skip_k1 else 2.0 * sum(in_C) * sum(out_C) * k3)


def bias_vector(plan: ConvPlan, bias: np.ndarray) -> np.ndarray:
    """fp32 [n_nblk * nblk] bias laid out per N block (zero in padded columns)."""
    out = np.zeros(plan.n_nblk * plan.nblk, np.float32)
    for nb in range(plan.n_nblk):
        for n in range(plan.nblk):
            c = plan.nb_coff[nb] + n
            if c < plan.out_C[plan.nb_sel[nb]]:
                out[nb * plan.nblk + n] = bias[plan.nb_real0[nb] + n]
    return out


def bias_index(plan: ConvPlan) -> np.ndarray:
    """gather index into cat(bias, [0]) producing the per-block bias vector."""
    nreal = int(sum(plan.out_C))
    idx = np.full(plan.n_nblk * plan.nblk, nreal, np.int64)
    for nb in range(plan.n_nblk):
        n = np.arange(plan.nblk)
        ok = (plan.nb_coff[nb] + n) < plan.out_C[plan.nb_sel[nb]]
        idx[nb * plan.nblk:(nb + 1) * plan.nblk] = np.where(ok, plan.nb_real0[nb] + n, nreal)
    return idx


# =====================================================================================================
# weight-gradient plans (csrc/wgrad_gemm.cuh)
# =====================================================================================================
WG_MAX_G = 32
WG_J_XLIST = 8
WG_J_YLIST = WG_J_XLIST + 2 * WG_MAX_G
WG_J_ENT = WG_J_YLIST + 2 * WG_MAX_G
WG_E_SIZE = 18 + WG_MAX_G
WG_DY_BOX = HT * WT * 16


@dataclass
class WgradPlan:
    kind: str
    x_maps: List[Tuple[int, Optional[Tuple[int, int, int]]]]     # (x tensor index, parity)
    y_maps: List[Optional[Tuple[int, int, int]]]                 # parity of the dy view or None
    tab: np.ndarray
    jobs: list
    n_jobs: int
    job_stride: int
    split: int
    dw_numel: int
    ld: int
    gidx: np.ndarray            # gather: param_grad.flatten() = cat(dw, [0])[gidx]
    wx: int = 1                 # chunks per x TMA box (1 = 16-byte rows, no swizzle; 2/4/8 = SWIZZLE_32B/64B/128B rows)
    wy: int = 1                 # chunks per dy TMA box
    flops_per_voxel: float = 0.0
    # analytic form of gidx for unet3d_dw_unpack: parameter element = rowmap[row] + tap + col * col_stride
    unpack: Optional[dict] = None


WG_ENT_MAX = 16
_WG_DT_CANDIDATES = (8, 6, 5, 4, 3, 2, 1)


def make_wgrad_plan(kind: str, ks: int, stride: int, x_C: Sequence[int], y_C: int, dims: Tuple[int, int, int, int],
                    num_sms: int = 148) -> WgradPlan:
    """Memoised front end of _make_wgrad_plan (see make_conv_plan)."""
    key = ("wgrad", kind, ks, stride, tuple(x_C), y_C, tuple(dims), num_sms, tuple(os.environ.get(k) for k in _TUNING_ENV))
    if key not in _PLAN_CACHE:
        _PLAN_CACHE[key] = _make_wgrad_plan(kind, ks, stride, x_C, y_C, dims, num_sms)
    return _PLAN_CACHE[key]


def _make_wgrad_plan(kind: str, ks: int, stride: int, x_C: Sequence[int], y_C: int, dims: Tuple[int, int, int, int],
                     num_sms: int = 148) -> WgradPlan:
    """Weight gradient of
         kind='conv' : Conv3d weight (Cout=y_C, Cin=sum(x_C), k,k,k), stride 1 or 2 (x may be a concat)
         kind='convT': ConvTranspose3d(k3,s2,p1) weight (Cin=x_C[0], Cout=y_C, 3,3,3); dy lives on the fine grid
       dims = (N, D, H, W) of the TILE grid (= the coarse grid for strided / transposed layers).
       The kernel accumulates dw[kflat][Kp][Np] fp32 (K = x channels in padded concat order, N = dy channels)."""
    x_C = list(x_C)
    x_Cp = [pad_channels(c) for c in x_C]
    y_Cp = pad_channels(y_C)
    N_, D_, H_, W_ = dims
    k3 = ks ** 3
    parities = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    if kind == "convT":
        ks, stride, k3 = 3, 2, 27
    if kind == "conv" and stride == 2:
        x_maps = [(0, p) for p in (parities if ks == 3 else [(0, 0, 0)])]
    else:
        x_maps = [(i, None) for i in range(len(x_C))]
    y_maps = list(parities) if kind == "convT" else [None]
    xp_off = np.concatenate([[0], np.cumsum(x_Cp)]).astype(np.int64)
    XL = [dict(map=mi, ch=c * 8, k0=int(xp_off[ti]) + c * 8, par=p)
          for mi, (ti, p) in enumerate(x_maps) for c in range(x_Cp[ti] // 8)]
    YL = [dict(map=len(x_maps) + mi, ch=c * 8, n0=c * 8, par=p) for mi, p in enumerate(y_maps) for c in range(y_Cp // 8)]
    Kp, Np = int(sum(x_Cp)), y_Cp
    ld = Np
    dw_numel = k3 * Kp * Np
    dw_origin = 0

    if ks == 1:
        d_shifts, hw_shifts = [1], [(1, 1)]
    elif kind == "conv" and stride == 1:
        d_shifts, hw_shifts = [0, 1, 2], [(a, b) for a in range(3) for b in range(3)]
    elif kind == "conv":
        d_shifts, hw_shifts = [0, 1], [(a, b) for a in (0, 1) for b in (0, 1)]
    else:
        d_shifts, hw_shifts = [1, 2], [(a, b) for a in (1, 2) for b in (1, 2)]
    min_sd, span = min(d_shifts), max(d_shifts) - min(d_shifts) + 1

    def kflat_of(shift, p_x, p_y):
        kk = []
        for d in range(3):
            s = shift[d]
            if ks == 1:
                if s != 1 or (p_x is not None and p_x[d]) or (p_y is not None and p_y[d]):
                    return None
                kk.append(0)
            elif kind == "conv" and stride == 1:
                kk.append(s)
            elif kind == "conv":
                v = _S2_K.get((p_x[d], s))
                if v is None:
                    return None
                kk.append(v)
            else:
                v = _TR_K.get((p_y[d], s))
                if v is None:
                    return None
                kk.append(v)
        return (kk[0] * ks + kk[1]) * ks + kk[2]

    case_a = kind == "conv" and stride == 1 and ks == 3 and len(XL) <= 16 and 16 % len(XL) == 0
    # ---- w-shift mode (3x3x3 stride-1 layers whose sources and dy all have 17..32 channels: the level-0 layers, 41 % of
    # the net's FLOPs).  With N = 32 an M = 128, K = 16 MMA costs ~47 cycles whatever it computes (operand-read floor), so
    # nine (kh, kw) entries of N = 32 run at a third of the tensor peak.  Substituting u = v + (kw - 1) e_w in
    #     dW[kd,kh,kw] = sum_v x[v + (kd-1, kh-1, kw-1)] dy[v]  =  sum_u x[u + (kd-1, kh-1, 0)] dy[u - (kw-1) e_w]
    # moves the kw tap from the x operand to the dy operand: three copies of the dy tile, TMA-loaded at w0 + 1, w0, w0 - 1
    # (out-of-volume voxels are zero-filled, exactly the terms that must vanish), sit side by side as three MN atoms of
    # one swizzled B operand, so ONE MMA of N = 96 per kh does the work of three; the kd taps stay in M (four x planes).
    wshift = (case_a and len(YL) == 4 and all(cp == 32 for cp in x_Cp) and D_ >= 1
              and os.environ.get("U3D_WG_WSHIFT", "1") == "1" and not os.environ.get("U3D_WG_NOSW")
              and os.environ.get("U3D_WG_PAIR", "0") != "1")            # (the round-1 pair-mode experiment overrides it)
    if wshift:
        m_blocks = [(4 * i, 4) for i in range(len(x_Cp))]          # one job family per source: M = 4 planes x 32 channels
        ppm = 4
        units = [(0, sh, 1) for sh in range(3)]                     # kw is carried by the dy copies; x stays at the centre
        planes_extra = 4
    elif case_a:
        m_blocks = [(0, len(XL))]
        ppm = 16 // len(XL)
        units = [(pg, sh, sw) for pg in range(0, span, ppm) for (sh, sw) in hw_shifts]
        planes_extra = -(-span // ppm) * ppm
    else:
        m_blocks = [(m0, min(16, len(XL) - m0)) for m0 in range(0, len(XL), 16)]
        ppm = 1
        units = [(sd - min_sd, sh, sw) for sd in d_shifts for (sh, sw) in hw_shifts]
        planes_extra = span
    y_blocks = [(y0, min(16, len(YL) - y0)) for y0 in range(0, len(YL), 16)]

    # ---- pair mode (16/32-channel 3x3x3 layers on the 16-byte-row layout): an M = 128, K = 16 MMA costs max(N/2, ~50)
    # cycles, so N = 32 runs at a third of the tensor peak.  Two consecutive dy planes side by side make N = 64 at the
    # same cost per MMA: row block i (x plane d-1+i) against column block j (dy plane d+j) is tap kd = i - j.  The four
    # row blocks x two column blocks hold kd = -1..3; -1 and 3 land in two scratch slots in front of / behind the 27 real
    # taps (dw_origin), everything else accumulates where the single-plane layout puts it.  Twice the TMEM columns per
    # (kh,kw) entry, so the nine entries split over two jobs (5 + 4) that read the same tiles at the same time.
    pair = False       # decided below, once box_width exists

    def group_ok(lst, w, exact):
        if exact and len(lst) % w:
            return False
        for i in range(0, len(lst), w):
            grp = lst[i:i + w]
            if any(c["map"] != grp[0]["map"] or c["ch"] != grp[0]["ch"] + 8 * k for k, c in enumerate(grp)):
                return False
        return True

    pair = (case_a and D_ >= 2 and len(YL) <= 4 and len(XL) <= 8 and not (len(XL) == 8 and group_ok(XL, 8, True))
            and not os.environ.get("U3D_WG_FORCESW") and os.environ.get("U3D_WG_PAIR", "0") == "1")

    if pair:
        dw_origin = 9 * Kp * Np                  # scratch slot for kd = -1 in front, one for kd = 3 behind
        dw_numel = 45 * Kp * Np

    jobs = []
    p0_values = sorted(set(u[0] for u in units))
    for (y0, ycnt) in y_blocks:
        gy = -(-ycnt // 4) * 4 * (3 if wshift else 1)
        max_ent = min(512 // (gy * 8 * (2 if pair else 1)), WG_ENT_MAX)
        for (m0, gx) in m_blocks:
            for p0 in p0_values:          # one job never spans plane groups: keeps the x stage small
                us = [u for u in units if u[0] == p0]
                n_j = -(-len(us) // max_ent)
                per = -(-len(us) // n_j)
                for u0 in range(0, len(us), per):
                    jobs.append(dict(y0=y0, ycnt=ycnt, gy=gy, m0=m0, gx=gx, p0=p0, units=us[u0:u0 + per]))
    planes_extra = ppm

    # ---- swizzled whole-row boxes: w consecutive 8-channel chunks of one source are loaded as ONE TMA box with rows of
    # 16 w bytes (SWIZZLE_32B/64B/128B = the swizzled MN-major UMMA layouts; MN atoms of 8 w channels at the box pitch,
    # which also continues across planes).  16-byte-row boxes cost a 32-byte L2 sector and a shared-memory write
    # wavefront per row: every layer but the 30-channel ones was load-bound by 1.4-2x (profiles/r01_notes.md).
    def box_width(lists, exact):
        longest = max(len(lst) for lst in lists)
        for w in (8, 4, 2):
            if w > longest:                 # never pad a short channel list up to a wider box
                continue
            ok = True
            for lst in lists:
                if exact and len(lst) % w:
                    ok = False
                for i in range(0, len(lst), w):
                    grp = lst[i:i + w]
                    if any(c["map"] != grp[0]["map"] or c["ch"] != grp[0]["ch"] + 8 * k for k, c in enumerate(grp)):
                        ok = False
                if not ok:
                    break
            if ok:
                return w
        return 1

    xls = [[XL[j["m0"] + i] for i in range(j["gx"])] for j in jobs]
    yls = [[YL[j["y0"] + i] for i in range(j["ycnt"])] for j in jobs]
    wx = box_width(xls, case_a)
    wy = box_width(yls, False)
    if os.environ.get("U3D_WG_NOSW"):
        wx = wy = 1
    use_sw = wx > 1 and wy > 1 and not pair
    if wshift:
        assert wx == 4 and wy == 4
    elif case_a and wx < 8 and not os.environ.get("U3D_WG_FORCESW"):
        # 16/32-channel 3x3x3 layers are bound by the tensor core's shared-memory operand reads, not by the loads, and a
        # 32/64-byte MN-major row fills only part of a 128-byte read wavefront: measured 0.436 vs 0.412 ms (30->30) and
        # 0.951 vs 0.853 ms (60->30 concat) at 2x128^3 -- they keep the 16-byte-row layout (8 rows x 16 B = one wavefront)
        use_sw = False
    x_pitch = -(-((HT + 2) * (WT + 2) * 16 * wx) // (128 * wx)) * (128 * wx)       # whole swizzle atoms (8 rows)
    y_pitch = WG_DY_BOX * wy

    job_stride = WG_J_ENT + WG_E_SIZE * WG_ENT_MAX
    tab = np.zeros((len(jobs), job_stride), np.int64)
    for ji, j in enumerate(jobs):
        gx, gy = j["gx"], j["gy"]
        if use_sw:
            nbx, nby = -(-gx // wx), -(-gy // wy)
            margin = 0 if case_a else (16 // wx - nbx) * x_pitch
            x_plane, y_plane = nbx * x_pitch, nby * y_pitch
        else:
            margin = 0 if (case_a or gx == 16) else (16 - gx) * CHUNK_PITCH
            x_plane, y_plane = gx * CHUNK_PITCH, gy * WG_DY_BOX
        dt = None
        for cand in _WG_DT_CANDIDATES:
            if cand > max(1, D_) + (D_ % 2 if pair else 0) or (pair and cand % 2):
                continue
            px = cand - (2 if pair else 1) + planes_extra
            if 2048 + 2 * (px * x_plane + cand * y_plane) + margin <= SMEM_LIMIT:
                dt = cand
                break
        if dt is None:
            raise ValueError("wgrad stage does not fit shared memory")
        px = dt - (2 if pair else 1) + planes_extra
        j["dt"], j["px"] = dt, px
        row = tab[ji]
        row[0:7] = [dt, px, min_sd - 1 + j["p0"], gx, gy, len(j["units"]), ld]
        xl = [XL[j["m0"] + i] for i in range(gx)]
        yl = [YL[j["y0"] + i] if i < j["ycnt"] else YL[j["y0"]] for i in range(gy // (3 if wshift else 1))]
        j["xl"], j["yl"] = xl, yl
        if use_sw:
            row[7] = wx | (wy << 8) | (nbx << 16) | (nby << 24)
            for b in range(nbx):
                row[WG_J_XLIST + 2 * b], row[WG_J_XLIST + 2 * b + 1] = xl[b * wx]["map"], xl[b * wx]["ch"]
            for b in range(nby):
                if wshift:              # box b = the dy tile for kw = b, loaded at w0 + (1 - kw): code 3 / 2 / 1 in bits 16-17
                    row[WG_J_YLIST + 2 * b], row[WG_J_YLIST + 2 * b + 1] = YL[0]["map"], YL[0]["ch"] | ((3 - b) << 16)
                    continue
                c = YL[j["y0"] + b * wy] if b * wy < j["ycnt"] else YL[j["y0"]]
                row[WG_J_YLIST + 2 * b], row[WG_J_YLIST + 2 * b + 1] = c["map"], c["ch"]
        else:
            if pair:
                row[7] = 1 << 30          # wx = 0 (16-byte rows) + pair flag
            for i, c in enumerate(xl):
                row[WG_J_XLIST + 2 * i], row[WG_J_XLIST + 2 * i + 1] = c["map"], c["ch"]
            for i, c in enumerate(yl):
                row[WG_J_YLIST + 2 * i], row[WG_J_YLIST + 2 * i + 1] = c["map"], c["ch"]
        col = 0
        for e, (p0, sh, sw) in enumerate(j["units"]):
            ent = row[WG_J_ENT + e * WG_E_SIZE: WG_J_ENT + (e + 1) * WG_E_SIZE]
            ent[0] = (sh * (WT + 2) + sw) * 16
            ent[1] = col
            col += gy * 8 * (2 if pair else 1)
            ent[2:] = -1
            if wshift:
                for s in range(16):
                    sd = p0 + s // gx
                    if sd <= 2:
                        ent[2 + s] = (((sd * 3 + sh) * 3 + 0) * Kp + xl[s % gx]["k0"]) * ld
                for kw in range(3):
                    for h in range(4):
                        ent[18 + kw * 4 + h] = kw * Kp * ld + yl[h]["n0"]
                continue
            if pair:
                # rows: x plane shift sd = 0..3 (3 pairs only with the second dy plane); columns: (dy plane j, chunk h)
                for s in range(16):
                    sd = p0 + s // gx
                    if 0 <= sd <= 3:
                        ent[2 + s] = (((sd * 3 + sh) * 3 + sw) * Kp + xl[s % gx]["k0"]) * ld
                for jp in (0, 1):
                    for h in range(j["ycnt"]):
                        ent[18 + jp * gy + h] = dw_origin + yl[h]["n0"] - jp * 9 * Kp * ld
                continue
            for s in range(16):
                plane_rel = p0 + s // gx if case_a else p0
                if not case_a and s >= gx:
                    continue
                sd = min_sd + plane_rel
                if sd not in d_shifts:
                    continue
                c = xl[s % gx]
                if kind == "convT":
                    ent[2 + s] = c["k0"] * ld
                else:
                    kf = kflat_of((sd, sh, sw), c["par"], None)
                    if kf is not None:
                        ent[2 + s] = (kf * Kp + c["k0"]) * ld
            for h in range(gy):
                if h >= j["ycnt"]:
                    continue
                c = yl[h]
                if kind == "convT":
                    kf = kflat_of((min_sd + p0, sh, sw), None, c["par"])
                    if kf is not None:
                        ent[18 + h] = kf * Kp * ld + c["n0"]
                else:
                    ent[18 + h] = c["n0"]
        assert col <= 512
    assert tab.max() < 2 ** 31

    # gather index from dw[kflat][Kp][Np] back to the PyTorch parameter layout
    Ktot, Ntot = int(sum(x_C)), int(y_C)
    kreal = np.concatenate([int(xp_off[t]) + np.arange(c) for t, c in enumerate(x_C)])     # padded K index of real channel
    if kind == "conv":      # W[cout][cin][k]
        co, ci, kf = np.meshgrid(np.arange(Ntot), np.arange(Ktot), np.arange(k3), indexing="ij")
    else:                   # Wt[cin][cout][k]
        ci, co, kf = np.meshgrid(np.arange(Ktot), np.arange(Ntot), np.arange(k3), indexing="ij")
    gidx = dw_origin + (kf * Kp + kreal[ci]) * Np + co
    rowmap = np.full(Kp, -1, np.int32)
    rowmap[kreal] = np.arange(Ktot) * (k3 if kind == "conv" else Ntot * k3)
    unpack = dict(k3=k3, Kp=Kp, Np=Np, ncols=Ntot, col_stride=Ktot * k3 if kind == "conv" else k3, rowmap=rowmap,
                  origin=dw_origin)
    n_tiles_min = N_ * (-(-max(1, D_) // 8)) * (-(-H_ // HT)) * (-(-W_ // WT))
    n_tiles_max = N_ * max(1, D_) * (-(-H_ // HT)) * (-(-W_ // WT))
    # CTAs = jobs x split, one CTA per SM at a time (227 KB of shared memory): never spill a few CTAs into an extra
    # wave (6 jobs x 50 = 300 CTAs ran 40 % slower than 6 x 74 = 444 = exactly three waves), so round DOWN to whole
    # waves.  One wave unless there are more jobs than SMs (measured: fewer, longer CTAs win -- less split-K
    # reduction traffic and pipeline fill).  Env U3D_WG_WAVES overrides (sweeps).
    waves = int(os.environ.get("U3D_WG_WAVES", "1"))
    if len(jobs) > waves * num_sms:
        waves = -(-len(jobs) // num_sms)
    split = max(1, min(n_tiles_max, (waves * num_sms) // len(jobs)))
    return WgradPlan(kind=kind, x_maps=x_maps, y_maps=y_maps, tab=tab.reshape(-1).astype(np.int32), jobs=jobs,
                     n_jobs=len(jobs), job_stride=job_stride, split=split, dw_numel=dw_numel, ld=ld,
                     gidx=gidx.reshape(-1).astype(np.int64), flops_per_voxel=2.0 * Ktot * Ntot * k3,
                     wx=wx if use_sw else 1, wy=wy if use_sw else 1, unpack=unpack)

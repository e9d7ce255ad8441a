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
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

HT, WT = 16, 8                 # output tile (rows of one UMMA): 16 x 8 voxels
CHUNK_PITCH = 2944
W_STAGES = 6
SMEM_LIMIT = 227 * 1024


def pad_channels(c: int) -> int:
    """Physical channel count of a bf16 NDHWC activation: multiple of 16 (UMMA K step)."""
    return (c + 15) // 16 * 16


def conv_smem_bytes(dt: int, g: int, nblk: int) -> int:
    return 2048 + 2 * (dt + 2) * g * CHUNK_PITCH + W_STAGES * g * nblk * 16


def choose_nblk(cp_out: int) -> Tuple[int, int]:
    best = None
    for nblk in (128, 96, 64, 32):
        n = -(-cp_out // nblk)
        waste = n * nblk - cp_out
        if best is None or waste < best[0]:
            best = (waste, nblk, n)
    return best[1], best[2]


def choose_dt_g(nblk: int, chunk_counts: Sequence[int], depth: int) -> Tuple[int, int]:
    """Planes per segment and chunks per channel group: biggest Dt (halo amortisation), then biggest G."""
    gs = [g for g in (4, 6, 2) if all(c % g == 0 for c in chunk_counts)]
    if not gs:
        raise ValueError(f"channel chunk counts {chunk_counts} need an even common divisor")
    dt_max = max(1, min(256 // nblk, 8, depth))
    for dt in range(dt_max, 0, -1):
        for g in gs:
            if conv_smem_bytes(dt, g, nblk) <= SMEM_LIMIT:
                return dt, g
    raise ValueError("no (Dt, G) fits shared memory")


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
    n_tiles_w: int = 0              # number of weight tiles

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


def make_conv_plan(kind: str, ks: int, stride: int, in_C: Sequence[int], out_C: Sequence[int], depth: int) -> ConvPlan:
    """kind:
         conv_fwd    Conv3d forward (weight (Cout, Cin, k,k,k)); inputs may be a concat (len(in_C) > 1)
         conv_dgrad  its data gradient; outputs may be a concat split (len(out_C) > 1)
         convT_fwd   ConvTranspose3d(k3,s2,p1) forward + zero pad plane (weight (Cin, Cout, k,k,k))
         convT_dgrad its data gradient
       in_C / out_C are the REAL channel counts of the A-side / output-side tensors of this call
       (for a gradient, in_C is the channel count of dy).  `depth` = D extent of the tile grid."""
    in_C, out_C = list(in_C), list(out_C)
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
    if (pattern == "transposed" or stride == 2) and (len(in_C) != 1 or len(out_C) != 1):
        raise ValueError("strided / transposed convs take one input and one output")

    # ---- A maps and channel groups
    parities = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    if pattern == "direct" and stride == 2:
        par_list = parities if ks == 3 else [(0, 0, 0)]
        maps = [(0, p) for p in par_list]
    else:
        maps = [(i, None) for i in range(len(in_C))]
    chunk_counts = [in_Cp[m[0]] // 8 for m in maps]

    # ---- N blocks
    nb_sel, nb_coff, nb_ooff, nb_real0 = [], [], [], []
    nblk = None
    if pattern == "transposed":
        nblk, n = choose_nblk(out_Cp[0])
        out_par = parities if ks == 3 else [(0, 0, 0)]
        for p in out_par:
            for j in range(n):
                nb_sel.append(0); nb_coff.append(j * nblk); nb_ooff.append(p); nb_real0.append(j * nblk)
    else:
        # one blocking for every output tensor (dgrad of a concat): use the blocking of the widest
        nblk, _ = choose_nblk(max(out_Cp))
        if any(cp % nblk and cp > nblk for cp in out_Cp) and len(out_Cp) > 1:
            nblk = 32
        real0 = 0
        for t, cp in enumerate(out_Cp):
            for j in range(-(-cp // nblk)):
                nb_sel.append(t); nb_coff.append(j * nblk); nb_ooff.append((0, 0, 0)); nb_real0.append(real0 + j * nblk)
            real0 += out_C[t]
    Dt, G = choose_dt_g(nblk, chunk_counts, depth)

    cg_map, cg_ch = [], []
    for mi, cnt in enumerate(chunk_counts):
        for i in range(cnt // G):
            cg_map.append(mi); cg_ch.append(i * G * 8)

    # ---- taps
    if ks == 1:
        shifts = [(1, 1, 1)]
    elif pattern == "direct" and stride == 1:
        shifts = [(a, b, c) for a in range(3) for b in range(3) for c in range(3)]
    elif pattern == "direct":
        shifts = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    else:
        shifts = [(a, b, c) for a in (1, 2) for b in (1, 2) for c in (1, 2)]

    # real channel index of every K slot: concat offset of the source tensor + channel, -1 for padding
    in_off = np.concatenate([[0], np.cumsum(in_C)]).astype(np.int64)
    Ktot, Ntot = int(sum(in_C)), int(sum(out_C))
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
            for t, sh in enumerate(shifts):
                kk = _kidx_and_valid(kind, ks, stride, pattern, sh, par_in or (0, 0, 0), nb_ooff[nb])
                if kk is None:
                    continue
                if kind == "conv_dgrad" and pattern == "direct":
                    kk = tuple(ks - 1 - v for v in kk)                        # flipped kernel
                kflat = (kk[0] * ks + kk[1]) * ks + kk[2]
                masks[nb, cg] |= np.uint32(1 << t)
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
                flat = np.where((K >= 0) & (Nn >= 0), flat, -1)
                pieces.append(flat.reshape(-1))
                n_tiles += 1
    widx = np.concatenate(pieces) if pieces else np.zeros(0, np.int64)
    if kind in ("conv_fwd", "convT_dgrad"):
        numel = Ntot * Ktot * k3
    else:
        numel = Ktot * Ntot * k3
    widx = np.where(widx < 0, numel, widx)          # slot `numel` of cat(W.flatten(), [0]) is the zero

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
                    omul=2 if pattern == "transposed" else 1, n_tiles_w=n_tiles)


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
    y_maps: List[Optional[Tuple[int, int, int]]]                 # parity of dy view or None
    tab: np.ndarray
    n_jobs: int
    job_stride: int
    split: int
    dw_numel: int
    gidx: np.ndarray            # gather: param_grad.flatten() = cat(dw, [0])[gidx]
    n_ent_max: int = 0


def make_wgrad_plan(kind: str, ks: int, stride: int, x_C: Sequence[int], y_C: int, dims: Tuple[int, int, int, int],
                    num_sms: int = 148) -> WgradPlan:
    """Weight gradient of
         conv  (kind='conv',  weight (Cout=y_C, Cin=sum(x_C), k,k,k), stride 1 or 2; x may be a concat)
         convT (kind='convT', weight (Cin=x_C[0], Cout=y_C, 3,3,3), stride 2): dy lives on the fine grid.
       dims = (N, D, H, W) of the TILE grid (the coarse grid for strided / transposed layers).
       The kernel accumulates dw[kflat][K_pad][N_pad] (K = x channels, N = dy channels, fp32)."""
    x_C = list(x_C)
    x_Cp = [pad_channels(c) for c in x_C]
    y_Cp = pad_channels(y_C)
    N_, D_, H_, W_ = dims
    k3 = ks ** 3
    parities = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    # ---- operand chunk lists: (map id, channel, real channel index or -1, parity)
    if kind == "conv" and stride == 2:
        par = parities if ks == 3 else [(0, 0, 0)]
        x_maps = [(0, p) for p in par]
    else:
        x_maps = [(i, None) for i in range(len(x_C))]
    if kind == "convT":
        y_maps = list(parities)
    else:
        y_maps = [None]
    x_off = np.concatenate([[0], np.cumsum(x_C)]).astype(np.int64)
    xch = []          # per x chunk: (map, ch, real0, nreal, parity)
    for mi, (ti, p) in enumerate(x_maps):
        for c in range(x_Cp[ti] // 8):
            real = [x_off[ti] + c * 8 + j if c * 8 + j < x_C[ti] else -1 for j in range(8)]
            xch.append((mi, c * 8, real, p))
    ych = []
    for mi, p in enumerate(y_maps):
        for c in range(y_Cp // 8):
            real = [c * 8 + j if c * 8 + j < y_C else -1 for j in range(8)]
            ych.append((len(x_maps) + mi, c * 8, real, p))
    Ktot, Ntot = int(sum(x_C)), int(y_C)
    Kp = ((Ktot + 7) // 8) * 8
    Np = ((Ntot + 7) // 8) * 8
    ld = Np
    dw_numel = k3 * Kp * Np

    def kflat_of(shift, p_x, p_y):
        kk = []
        for d in range(3):
            s = shift[d]
            if ks == 1:
                if s != 1 or (p_x and p_x[d]) or (p_y and p_y[d]):
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

    gx_plane = len(xch) if not (kind == "conv" and stride == 1 and len(xch) < 16) else len(xch)
    # in-plane (kh, kw) shifts and the d shifts
    if ks == 1:
        hw_shifts, d_shifts = [(1, 1)], [1]
    elif kind == "conv" and stride == 1:
        hw_shifts, d_shifts = [(a, b) for a in range(3) for b in range(3)], [0, 1, 2]
    elif kind == "conv":
        hw_shifts, d_shifts = [(a, b) for a in (0, 1) for b in (0, 1)], [0, 1]
    else:
        hw_shifts, d_shifts = [(a, b) for a in (1, 2) for b in (1, 2)], [1, 2]

    # ---- jobs.  x chunks of one plane are cut into M blocks of 16 chunk slots; when a plane has
    # fewer than 16 chunks an M block runs on into the following planes (d shifts for free).
    n_xc = len(xch)
    jobs = []
    # N blocks: up to 32 dy chunks (256 columns), multiple of 4 chunks
    y_blocks = []
    yb = 0
    max_gy = 16 if n_xc >= 16 else 32
    while yb < len(ych):
        cnt = min(max_gy, len(ych) - yb)
        y_blocks.append((yb, cnt))
        yb += cnt
    if n_xc >= 16:
        m_blocks = [(mb, min(16, n_xc - mb)) for mb in range(0, n_xc, 16)]     # (first chunk, valid chunks)
        planes_per_m = 1
    else:
        m_blocks = [(0, n_xc)]
        planes_per_m = 16 // n_xc if 16 % n_xc == 0 else 1
        if 16 % n_xc != 0:
            raise ValueError("x chunk count must divide 16 or be >= 16")
    for (y0, ycnt) in y_blocks:
        gy = ((ycnt + 3) // 4) * 4
        nblk = gy * 8
        max_ent = 512 // nblk
        for (m0, mcnt) in m_blocks:
            if n_xc >= 16:
                # one plane per MMA: entries = (d shift, hw shift)
                units = [(ds, hw) for ds in d_shifts for hw in hw_shifts]
                gx_job, x_first = 16, m0
                # chunk list padded to 16 with repeats of a valid chunk (rows discarded)
            else:
                ppm = planes_per_m
                d_groups = sorted(set(ds // ppm * ppm for ds in d_shifts)) if ppm > 1 else d_shifts
                # with ppm planes per M block a unit covers d shifts [dg, dg + ppm)
                units = [(dg, hw) for dg in (range(min(d_shifts), max(d_shifts) + 1, ppm)) for hw in hw_shifts]
                gx_job, x_first = n_xc, 0
            for u0 in range(0, len(units), max_ent):
                jobs.append(dict(y0=y0, ycnt=ycnt, gy=gy, m0=m0, mcnt=mcnt, gx=gx_job, x_first=x_first,
                                 units=units[u0:u0 + max_ent]))

    # ---- per job: Dt (planes of dy per stage) limited by shared memory
    n_ent_max = max(len(j["units"]) for j in jobs)
    job_stride = WG_J_ENT + WG_E_SIZE * n_ent_max
    tab = np.zeros((len(jobs), job_stride), np.int32)
    for ji, j in enumerate(jobs):
        gx, gy = j["gx"], j["gy"]
        ppm = (16 // gx) if gx < 16 else 1
        dmin = min(u[0] for u in j["units"])
        dmax = max(u[0] for u in j["units"]) + ppm - 1          # highest plane offset touched (incl. junk rows)
        if gx < 16:
            dmax = max(dmax, dmin + ppm - 1)
        span = dmax - dmin + 1
        dt = None
        for cand in (8, 6, 4, 3, 2, 1):
            if cand > max(1, D_):
                continue
            px = cand + span - 1
            if 2048 + 2 * (px * gx * CHUNK_PITCH + cand * gy * WG_DY_BOX) <= SMEM_LIMIT:
                dt = cand
                break
        if dt is None:
            raise ValueError("wgrad stage does not fit shared memory")
        px = dt + span - 1
        row = tab[ji]
        row[0], row[1], row[2], row[3], row[4], row[5], row[6] = dt, px, dmin - 1, gx, gy, len(j["units"]), ld
        # x chunk list
        xl = [xch[(j["x_first"] + i) if (j["x_first"] + i) < n_xc and i < (j["mcnt"] if gx == 16 else gx) else j["x_first"]]
              for i in range(gx)]
        x_valid = [i < (j["mcnt"] if gx == 16 else gx) for i in range(gx)]
        for i, c in enumerate(xl):
            row[WG_J_XLIST + 2 * i], row[WG_J_XLIST + 2 * i + 1] = c[0], c[1]
        yl = [ych[j["y0"] + i] if i < j["ycnt"] else ych[j["y0"]] for i in range(gy)]
        y_valid = [i < j["ycnt"] for i in range(gy)]
        for i, c in enumerate(yl):
            row[WG_J_YLIST + 2 * i], row[WG_J_YLIST + 2 * i + 1] = c[0], c[1]
        col = 0
        for e, (ds0, hw) in enumerate(j["units"]):
            ent = row[WG_J_ENT + e * WG_E_SIZE: WG_J_ENT + (e + 1) * WG_E_SIZE]
            ent[0] = (ds0 - dmin) * gx * CHUNK_PITCH + (hw[0] * (WT + 2) + hw[1]) * 16
            ent[1] = col
            col += gy * 8
            # rows: 16 chunk slots, slot s -> plane offset ds0 + s // gx, x chunk s % gx
            for s in range(16):
                ds = ds0 + s // gx
                ci = s % gx
                ok = ds in d_shifts and x_valid[ci]
                ent[2 + s] = -1
                if ok:
                    ent[2 + s] = 1 << 30          # resolved below with the column parity (needs kflat)
            ent[18:18 + WG_MAX_G] = -1
            # dW element = row_off[g] + (r % 8) * ld + col_off[h] + c % 8 with dw[kflat][K][N]:
            #   when kflat depends on the x parity (strided conv) it goes into row_off; when it depends on
            #   the dy parity (convT) it goes into col_off; otherwise into row_off.
            for s in range(16):
                if ent[2 + s] < 0:
                    continue
                ds = ds0 + s // gx
                c = xl[s % gx]
                real0 = c[2][0]
                if real0 < 0:
                    ent[2 + s] = -1
                    continue
                if kind == "convT":
                    ent[2 + s] = real0 * ld
                else:
                    kf = kflat_of((ds, hw[0], hw[1]), c[3], None)
                    ent[2 + s] = -1 if kf is None else (kf * Kp + real0) * ld
            for h in range(gy):
                if not y_valid[h]:
                    continue
                c = yl[h]
                real0 = c[2][0]
                if real0 < 0:
                    continue
                if kind == "convT":
                    kf = kflat_of((ds0, hw[0], hw[1]), None, c[3])
                    ent[18 + h] = -1 if kf is None else kf * Kp * ld + real0
                else:
                    ent[18 + h] = real0
        assert col <= 512

    # ---- gather index from dw[kflat][Kp][Np] back to the parameter layout
    if kind == "conv":      # W[cout][cin][k]
        co, ci, kf = np.meshgrid(np.arange(Ntot), np.arange(Ktot), np.arange(k3), indexing="ij")
        gidx = (kf * Kp + ci) * Np + co
    else:                   # Wt[cin][cout][k]
        ci, co, kf = np.meshgrid(np.arange(Ktot), np.arange(Ntot), np.arange(k3), indexing="ij")
        gidx = (kf * Kp + ci) * Np + co
    segs_min = 1
    n_tiles = N_ * max(1, D_) * (-(-H_ // HT)) * (-(-W_ // WT))
    split = max(1, min(n_tiles, (2 * num_sms) // max(1, len(jobs))))
    return WgradPlan(kind=kind, x_maps=x_maps, y_maps=y_maps, tab=tab.reshape(-1), n_jobs=len(jobs),
                     job_stride=job_stride, split=split, dw_numel=dw_numel, gidx=gidx.reshape(-1).astype(np.int64),
                     n_ent_max=n_ent_max)

"""CPU execution of the host plans (plan.py) with torch ops -- TEST INFRASTRUCTURE.

Runs exactly the arithmetic the tcgen05 kernels are told to do by a plan (same tables, same packed
weight stream, same source views / shifts / masks), ignoring only the tiling.  Lets the CPU suite
prove the integer bookkeeping (tap tables, parity views, weight packing) against torch.nn.functional
without a GPU; the GPU suite then only has to prove the kernels execute a plan faithfully.
"""
import numpy as np
import torch


def pack_weights(plan, w: torch.Tensor) -> torch.Tensor:
    flat = torch.cat([w.reshape(-1).float(), torch.zeros(1)])          # index -1 (structural zero) -> the last slot
    return flat[torch.from_numpy(plan.widx.astype(np.int64))]


def run_conv_plan(plan, inputs, wpacked, grid, out_dims, bias_vec=None, zero_last=False):
    """inputs: list of (N, D, H, W, Cp) float tensors; grid = (N, D, H, W) tile-grid extents;
    out_dims = (D, H, W) of the output tensors.  Returns list of (N, *out_dims, Cp_out) tensors."""
    N, D, H, W = grid
    outs = [torch.zeros(N, *out_dims, cp) for cp in plan.out_Cp]
    G, nblk = plan.G, plan.nblk
    fuse = 3 if plan.fuse_kd else 1
    tile_elems = G * fuse * nblk * 8
    for nb in range(plan.n_nblk):
        acc = torch.zeros(N, D, H, W, nblk)
        tptr = plan.wbase[nb]
        for cg in range(plan.n_cg):
            mask = int(plan.masks[nb, cg])
            if mask == 0:
                continue
            ti, par = plan.maps[plan.cg_map[cg]]
            v = inputs[ti]
            if par is not None:
                v = v[:, par[0]::2, par[1]::2, par[2]::2, :]
            ch = plan.cg_ch[cg]
            v = v[..., ch:ch + G * 8]
            # brick coordinate s <-> view index o + s - 1; pad generously with zeros (TMA OOB fill)
            vp = torch.zeros(N, D + 2, H + 2, W + 2, G * 8)
            d1, h1, w1 = min(v.shape[1], D + 1), min(v.shape[2], H + 1), min(v.shape[3], W + 1)
            vp[:, 1:1 + d1, 1:1 + h1, 1:1 + w1] = v[:, :d1, :h1, :w1]
            for t, (sd, sh, sw) in enumerate(plan.shifts):
                if not (mask >> t) & 1:
                    continue
                tile = wpacked[tptr * tile_elems:(tptr + 1) * tile_elems].reshape(G, fuse, nblk, 8)
                tptr += 1
                for f in range(fuse):                      # fused tiles hold the d-shifts 2, 1, 0 in that order
                    sdd = (2 - f) if plan.fuse_kd else sd
                    a = vp[:, sdd:sdd + D, sh:sh + H, sw:sw + W]
                    acc += torch.einsum("bdhwk,kn->bdhwn", a, tile[:, f].permute(0, 2, 1).reshape(G * 8, nblk))
        if bias_vec is not None:
            acc += bias_vec[nb * nblk:(nb + 1) * nblk]
        o = outs[plan.nb_sel[nb]]
        od, oh, ow = plan.nb_ooff[nb]
        m = plan.omul
        c0 = plan.nb_coff[nb]
        c1 = min(c0 + nblk, o.shape[-1])
        o[:, od::m, oh::m, ow::m, c0:c1] = acc[..., :c1 - c0]
    if zero_last:
        for o in outs:
            o[:, -1] = 0
            o[:, :, -1] = 0
            o[:, :, :, -1] = 0
    return outs


def to_ndhwc(x: torch.Tensor, cp: int) -> torch.Tensor:
    """(N, C, D, H, W) -> (N, D, H, W, Cp) zero padded."""
    n, c = x.shape[:2]
    out = torch.zeros(n, *x.shape[2:], cp)
    out[..., :c] = x.permute(0, 2, 3, 4, 1)
    return out


def from_ndhwc(x: torch.Tensor, c: int) -> torch.Tensor:
    return x[..., :c].permute(0, 4, 1, 2, 3).contiguous()


def run_wgrad_plan(plan, xs, dy, grid):
    """Execute a WgradPlan's job table on CPU.  xs: list of (N, D, H, W, Cp) tensors (the x views are taken
    from them by parity), dy: (N, D', H', W', Cp).  Returns the flat fp32 dw buffer (+1 zero slot)."""
    import importlib
    P = importlib.import_module("unet3d_b200.plan")
    N, D, H, W = grid
    dw = torch.zeros(plan.dw_numel + 1, dtype=torch.float64)
    tab = torch.from_numpy(plan.tab.astype(np.int64)).reshape(plan.n_jobs, plan.job_stride)

    def view(t, par):
        return t if par is None else t[:, par[0]::2, par[1]::2, par[2]::2, :]

    def shifted(v, ch, sd, sh, sw):
        # brick shift s <-> view index o + s - 1, zero outside (TMA OOB fill); result on the tile grid
        vp = torch.zeros(N, D + 3, H + 2, W + 2, 8, dtype=torch.float64)      # pair mode shifts by up to 3 planes
        d1, h1, w1 = min(v.shape[1], D + 1), min(v.shape[2], H + 1), min(v.shape[3], W + 1)
        if ch < v.shape[-1]:           # channels past the tensor are TMA out-of-bounds zeros (partial boxes)
            vp[:, 1:1 + d1, 1:1 + h1, 1:1 + w1] = v[:, :d1, :h1, :w1, ch:ch + 8]
        return vp[:, sd:sd + D, sh:sh + H, sw:sw + W]

    maps = [(xs[ti], par) for (ti, par) in plan.x_maps] + [(dy, par) for par in plan.y_maps]
    for ji in range(plan.n_jobs):
        row = tab[ji]
        dt, px, xd0, gx, gy, n_ent, ld = (int(v) for v in row[0:7])
        sw_word = int(row[7])
        wx, wy, nbx, nby = sw_word & 0xff, (sw_word >> 8) & 0xff, (sw_word >> 16) & 0xff, (sw_word >> 24) & 0xff
        pair = (sw_word >> 30) & 1       # two dy planes side by side in N: column group = (plane j, chunk h)
        if wx:      # swizzled whole-row boxes: the lists hold one (map, first channel) per box of wx / wy chunks
            bx = [(int(row[P.WG_J_XLIST + 2 * b]), int(row[P.WG_J_XLIST + 2 * b + 1])) for b in range(nbx)]
            by = [(int(row[P.WG_J_YLIST + 2 * b]), int(row[P.WG_J_YLIST + 2 * b + 1])) for b in range(nby)]
            xl = [(bx[i // wx][0], bx[i // wx][1] + 8 * (i % wx)) for i in range(nbx * wx)]
            # bits 16-17 of a dy box's channel word: w shift code (w-shift mode: 3 = +1, 2 = 0, 1 = -1; 0 = no shift)
            yl = [(by[i // wy][0], (by[i // wy][1] & 0xffff) + 8 * (i % wy), ((by[i // wy][1] >> 16) & 3)) for i in range(nby * wy)]
            gx_mem = nbx * wx          # chunks per plane as laid out in shared memory (the MMA's M index runs over them)
        else:
            xl = [(int(row[P.WG_J_XLIST + 2 * i]), int(row[P.WG_J_XLIST + 2 * i + 1])) for i in range(gx)]
            yl = [(int(row[P.WG_J_YLIST + 2 * i]), int(row[P.WG_J_YLIST + 2 * i + 1]), 0) for i in range(gy)]
            gx_mem = gx
        for e in range(n_ent):
            ent = row[P.WG_J_ENT + e * P.WG_E_SIZE: P.WG_J_ENT + (e + 1) * P.WG_E_SIZE]
            a_off = int(ent[0])
            slot0, rem = divmod(a_off, P.CHUNK_PITCH)
            sh, sw = divmod(rem // 16, P.WT + 2)
            for s in range(16):
                ro = int(ent[2 + s])
                if ro < 0:
                    continue
                slot = slot0 + s
                plane_rel, ci = divmod(slot, gx_mem)
                assert plane_rel < px - dt + 1 + 0 or True
                sd = xd0 + 1 + plane_rel
                mx, chx = xl[ci]
                a = shifted(view(*maps[mx]), chx, sd, sh, sw)              # (N, D, H, W, 8)
                for hh in range(gy * (2 if pair else 1)):
                    co = int(ent[18 + hh])
                    if co < 0:
                        continue
                    jp, h = divmod(hh, gy)
                    my, chy, code = yl[h]
                    vy = view(*maps[my])
                    b = torch.zeros(N, D + 2, H, W, 8, dtype=torch.float64)
                    d1, h1, w1 = min(vy.shape[1], D), min(vy.shape[2], H), min(vy.shape[3], W)
                    if chy < vy.shape[-1]:
                        b[:, :d1, :h1, :w1] = vy[:, :d1, :h1, :w1, chy:chy + 8]
                    if code:        # the dy box was loaded at w0 + (code - 2): tile voxel w holds dy[w + code - 2], zero outside
                        wsh = code - 2
                        shifted_b = torch.zeros_like(b)
                        if wsh > 0:
                            shifted_b[:, :, :, :W - wsh] = b[:, :, :, wsh:W]
                        elif wsh < 0:
                            shifted_b[:, :, :, -wsh:W] = b[:, :, :, :W + wsh]
                        else:
                            shifted_b = b
                        b = shifted_b
                    if pair:       # the MMA of dy planes (d, d+1), d even: x plane d + shift against dy plane d + jp
                        assert dt % 2 == 0
                        n_pairs = (D + 1) // 2
                        a_use = torch.zeros(N, 2 * n_pairs, H, W, 8, dtype=torch.float64)
                        a_use[:, :D] = a
                        blk = torch.einsum("bdhwr,bdhwc->rc", a_use[:, 0::2], b[:, jp:jp + 2 * n_pairs:2])
                    else:
                        blk = torch.einsum("bdhwr,bdhwc->rc", a, b[:, :D])
                    for r in range(8):
                        dw[ro + r * ld + co: ro + r * ld + co + 8] += blk[r]
    return dw

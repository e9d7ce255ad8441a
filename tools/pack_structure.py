"""Next-round groundwork for the weight PACK kernel (gather_multi, 0.74 ms per step: one 4-byte word per 32-byte sector).

Checks on the CPU that every conv plan's gather index `widx` (dest element -> source element of the fp32 parameter) has the
separable form a sector-efficient kernel needs:

    dest  = tile * TE + g * (fuse * nblk * 8) + (f * nblk + n) * 8 + k8          TE = G * fuse * nblk * 8
    src   = ncol[tile][n] * sN + krow[tile][g * 8 + k8] * sK + ktap[tile][f]     (or -1 = structural zero)

with per-tile tables (ncol: nblk, krow: G*8, ktap: fuse entries) and two per-plan strides.  Then ONE thread can own a
(tile column n, 8-channel chunk g) pair across all taps: it reads a contiguous block of the parameter once and writes one
16-byte piece per (tap, f), coalesced across the warp -- the mirror image of csrc/elementwise.cu:dw_unpack_kernel.
    python tools/pack_structure.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import unet3d_b200.plan as P


def analyse(kind, ks, stride, in_C, out_C, dims, **kw):
    plan = P.make_conv_plan(kind, ks, stride, in_C, out_C, dims[0], grid=(2, *dims), **kw)
    G, nblk, fuse = plan.G, plan.nblk, 3 if plan.fuse_kd else 1
    TE = G * fuse * nblk * 8
    w = plan.widx.astype(np.int64).reshape(-1, G, fuse, nblk, 8)          # [tile][g][f][n][k8]
    ok_all, strides = True, set()
    for t in range(w.shape[0]):
        tile = w[t]
        valid = tile >= 0
        if not valid.any():
            continue
        # candidate strides from first valid differences
        g0, f0, n0, k0 = np.argwhere(valid)[0]
        base = tile[g0, f0, n0, k0]
        dn = [tile[g0, f0, n, k0] - base for n in range(nblk) if valid[g0, f0, n, k0] and n != n0]
        dk = [tile[g, f0, n0, k] - base for g in range(G) for k in range(8) if valid[g, f0, n0, k] and (g, k) != (g0, k0)]
        sN = min((abs(v) for v in dn if v), default=0)
        sK = min((abs(v) for v in dk if v), default=0)
        strides.add((sN, sK))
        # separability: tile[g,f,n,k] - tile[g0,f0,n0,k0] == a[n] + b[g,k] + c[f] wherever valid
        a = np.where(valid[g0, f0, :, k0], tile[g0, f0, :, k0] - base, 0)
        b = np.where(valid[:, f0, n0, :], tile[:, f0, n0, :] - base, 0)
        c = np.array([(tile[g0, f, n0, k0] - base) if valid[g0, f, n0, k0] else 0 for f in range(fuse)])
        pred = base + a[None, None, :, None] + b[:, None, None, :] + c[None, :, None, None]
        # a column / row / sub-tile that is structurally zero must be so everywhere
        n_ok, gk_ok, f_ok = valid.any(axis=(0, 1, 3)), valid.any(axis=(1, 2)), valid.any(axis=(0, 2, 3))
        want_valid = n_ok[None, None, :, None] & gk_ok[:, None, None, :] & f_ok[None, :, None, None]
        ok = np.array_equal(valid, want_valid) and np.array_equal(np.where(valid, tile, 0), np.where(valid, pred, 0))
        ok_all &= bool(ok)
    return ok_all, w.shape[0], TE, strides


if __name__ == "__main__":
    D2 = (128, 128, 128)
    cases = [("conv_fwd", 3, 1, [30], [30], D2), ("conv_dgrad", 3, 1, [30], [30], D2), ("conv_fwd", 3, 1, [30, 30], [30], D2),
             ("conv_dgrad", 3, 1, [30], [30, 30], D2), ("conv_fwd", 3, 2, [30], [60], (64, 64, 64)),
             ("conv_dgrad", 3, 2, [60], [30], (64, 64, 64)), ("conv_fwd", 3, 1, [480], [480], (8, 8, 8)),
             ("conv_dgrad", 3, 1, [480], [480], (8, 8, 8)), ("conv_fwd", 1, 1, [60, 60], [60], (64, 64, 64)),
             ("convT_fwd", 3, 2, [60], [30], (64, 64, 64)), ("convT_dgrad", 3, 2, [30], [60], (64, 64, 64))]
    for c in cases:
        try:
            ok, tiles, te, strides = analyse(*c)
            print(f"{c[0]:12s} k{c[1]} s{c[2]} {c[3]}->{c[4]}: separable={ok}, {tiles} tiles of {te} elements, (sN, sK) seen: {sorted(strides)[:4]}")
        except Exception as e:          # plan signature differences are reported, not fatal: this is a study tool
            print(f"{c[0]:12s} {c[3]}->{c[4]}: {type(e).__name__}: {e}")

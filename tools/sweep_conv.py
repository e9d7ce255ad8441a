"""Sweep the conv_gemm launch configuration (nblk, Dt, G, nbuf, fuse) of one layer on the GPU; prints the timings
sorted.  Used to calibrate plan.choose_config.  Usage: python tools/sweep_conv.py <layer-substring> [fwd|dgrad]"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops, plan as P

dev = "cuda"
N = 2
layers = {
    "L0": ("conv", 3, 1, [30], 30, (128, 128, 128)),
    "L0cat": ("conv", 3, 1, [30, 30], 30, (128, 128, 128)),
    "L0k1": ("conv", 1, 1, [30, 30], 30, (128, 128, 128)),
    "P0": ("conv", 3, 2, [30], 60, (64, 64, 64)),
    "L1": ("conv", 3, 1, [60], 60, (64, 64, 64)),
    "L1cat": ("conv", 3, 1, [60, 60], 60, (64, 64, 64)),
    "P1": ("conv", 3, 2, [60], 120, (32, 32, 32)),
    "L2": ("conv", 3, 1, [120], 120, (32, 32, 32)),
    "L2cat": ("conv", 3, 1, [120, 120], 120, (32, 32, 32)),
    "L3": ("conv", 3, 1, [240], 240, (16, 16, 16)),
    "L3cat": ("conv", 3, 1, [240, 240], 240, (16, 16, 16)),
    "L4": ("conv", 3, 1, [480], 480, (8, 8, 8)),
    "U0": ("convT", 3, 2, [60], 30, (64, 64, 64)),
    "U1": ("convT", 3, 2, [120], 60, (32, 32, 32)),
    "U3": ("convT", 3, 2, [480], 240, (8, 8, 8)),
}
name = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
kind, ks, stride, cins, cout, g = layers[name]
grid = (N, *g)
flops = 2.0 * N * g[0] * g[1] * g[2] * sum(cins) * cout * ks ** 3


def act(dims, c):
    return (torch.randn(N, *dims, P.pad_channels(c), device=dev) * 0.5).to(torch.bfloat16)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if kind == "conv":
    xs = [act(tuple(v * stride for v in g), c) for c in cins]
    y = act(g, cout)
    w = torch.randn(cout, sum(cins), ks, ks, ks, device=dev) * 0.05
else:
    xs = [act(g, cins[0])]
    y = act(tuple(2 * v for v in g), cout)
    w = torch.randn(cins[0], cout, 3, 3, 3, device=dev) * 0.05
dxs = [torch.zeros_like(x) for x in xs]
st = torch.zeros(N, y.shape[-1], 2, device=dev, dtype=torch.float64)
res = []
for nblk, dt, gg, nbuf, fuse in [(None,) * 5] + list(itertools.product((32, 64, 96, 128), (8, 4, 2, 1), (2, 4, 8), (2, 1), (3, 1))):
    if nblk is None:
        os.environ.pop("U3D_CONV_CFG", None)
    else:
        if dt * nblk * nbuf > 512 or fuse * nblk > 256 or dt > g[0]:
            continue
        os.environ["U3D_CONV_CFG"] = f"{nblk},{dt},{gg},{nbuf},{fuse}"
    try:
        if which == "fwd":
            pk = "conv_fwd" if kind == "conv" else "convT_fwd"
            pl = P.make_conv_plan(pk, ks, stride, cins, [cout], g[0], grid)
        else:
            pk = "conv_dgrad" if kind == "conv" else "convT_dgrad"
            pl = P.make_conv_plan(pk, ks, stride, [cout], cins, g[0], grid)
        if nblk is not None and (pl.nblk, pl.Dt, pl.G, pl.nbuf, 3 if pl.fuse_kd else 1) != (nblk, dt, gg, nbuf, fuse):
            continue
        dp = ops.DeviceConvPlan(pl, dev)
        wp = dp.packed_weight(w)
        if which == "fwd":
            t = timeit(lambda: ops.conv_gemm(dp, xs, wp, [y], grid, stats=st, zero_last=(kind == "convT")))
        else:
            t = timeit(lambda: ops.conv_gemm(dp, [y], wp, dxs, grid))
        ops.check_device_errors()
    except Exception as e:      # configuration not valid for this layer
        continue
    tag = "DEFAULT" if nblk is None else ""
    res.append((t, f"nblk{pl.nblk}x{pl.n_nblk} Dt{pl.Dt} G{pl.G} buf{pl.nbuf} fuse{3 if pl.fuse_kd else 1} wT{pl.wT} st{pl.w_stages} {tag}"))
res.sort()
print(f"{name} {which}: {flops / 1e9:.1f} GF")
for t, d in res[:12]:
    print(f"   {t * 1e3:8.1f} us  {flops / t / 1e9:7.1f} TF/s  {d}")
d = [r for r in res if "DEFAULT" in r[1]]
if d:
    print(f"   default: {d[0][0] * 1e3:.1f} us  {d[0][1]}")

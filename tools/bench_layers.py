"""Per-layer timing of the tensor-core kernels at the cfg-2 shapes (2 x 128^3, default net).
Prints ms and algorithmic TFLOP/s for forward, data-gradient and weight-gradient of each distinct conv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops, plan as P

dev = "cuda"
N = 2
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(os.environ.get("REPS", "5"))


def timeit(fn):
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


def act(dims, c):
    return (torch.randn(N, *dims, P.pad_channels(c), device=dev) * 0.5).to(torch.bfloat16)


layers = [  # name, kind, ks, stride, cins, cout, out dims (tile grid), count per fwd
    ("L0 30->30 k3", "conv", 3, 1, [30], 30, (128, 128, 128), 3),
    ("L0 60->30 k3 cat", "conv", 3, 1, [30, 30], 30, (128, 128, 128), 1),
    ("L0 60->30 k1 cat", "conv", 1, 1, [30, 30], 30, (128, 128, 128), 1),
    ("P0 30->60 k3 s2", "conv", 3, 2, [30], 60, (64, 64, 64), 1),
    ("P0 30->60 k1 s2", "conv", 1, 2, [30], 60, (64, 64, 64), 1),
    ("L1 60->60 k3", "conv", 3, 1, [60], 60, (64, 64, 64), 4),
    ("L1 120->60 k3 cat", "conv", 3, 1, [60, 60], 60, (64, 64, 64), 1),
    ("L2 120->120 k3", "conv", 3, 1, [120], 120, (32, 32, 32), 6),
    ("L3 240->240 k3", "conv", 3, 1, [240], 240, (16, 16, 16), 8),
    ("L4 480->480 k3", "conv", 3, 1, [480], 480, (8, 8, 8), 9),
    ("P3 240->480 k3 s2", "conv", 3, 2, [240], 480, (8, 8, 8), 1),
    ("P2 120->240 k3 s2", "conv", 3, 2, [120], 240, (16, 16, 16), 1),
    ("D3 480->240 k3 cat", "conv", 3, 1, [240, 240], 240, (16, 16, 16), 1),
    ("U0 60->30 convT", "convT", 3, 2, [60], 30, (64, 64, 64), 1),
    ("U3 480->240 convT", "convT", 3, 2, [480], 240, (8, 8, 8), 1),
]
tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
for name, kind, ks, stride, cins, cout, g, cnt in layers:
    if only and only not in name:
        continue
    grid = (N, *g)
    vox = N * g[0] * g[1] * g[2]
    flops = 2.0 * vox * sum(cins) * cout * ks ** 3
    if kind == "conv":
        in_dims = tuple(v * stride for v in g)
        xs = [act(in_dims, c) for c in cins]
        y = act(g, cout)
        w = torch.randn(cout, sum(cins), ks, ks, ks, device=dev) * 0.05
        fp = ops.DeviceConvPlan(P.make_conv_plan("conv_fwd", ks, stride, cins, [cout], g[0], grid), dev)
        dp = ops.DeviceConvPlan(P.make_conv_plan("conv_dgrad", ks, stride, [cout], cins, g[0], grid), dev)
        dxs = [torch.zeros_like(x) for x in xs]
    else:
        xs = [act(g, cins[0])]
        y = act(tuple(2 * v for v in g), cout)
        w = torch.randn(cins[0], cout, 3, 3, 3, device=dev) * 0.05
        fp = ops.DeviceConvPlan(P.make_conv_plan("convT_fwd", 3, 2, cins, [cout], g[0], grid), dev)
        dp = ops.DeviceConvPlan(P.make_conv_plan("convT_dgrad", 3, 2, [cout], cins, g[0], grid), dev)
        dxs = [torch.zeros_like(xs[0])]
    wp = ops.DeviceWgradPlan(P.make_wgrad_plan(kind, ks, stride, cins, cout, grid, 148), dev)
    st = torch.zeros(N, y.shape[-1], 2, device=dev, dtype=torch.float64)
    wf, wd = fp.packed_weight(w), dp.packed_weight(w)
    dw = torch.zeros(wp.plan.dw_numel + 1, device=dev)
    t_f = timeit(lambda: ops.conv_gemm(fp, xs, wf, [y], grid, stats=st, zero_last=(kind == "convT")))
    t_d = timeit(lambda: ops.conv_gemm(dp, [y], wd, dxs, grid))
    t_w = timeit(lambda: ops.wgrad_gemm(wp, xs, y, dw, grid))
    ops.check_device_errors()
    print(f"{name:20s} x{cnt}  GF {flops/1e9:7.1f} | fwd {t_f:7.3f} ms {flops/t_f/1e9:7.1f} TF/s (Dt{fp.plan.Dt} G{fp.plan.G} nblk{fp.plan.nblk}x{fp.plan.n_nblk} buf{fp.plan.nbuf} f{int(fp.plan.fuse_kd)}; dgrad Dt{dp.plan.Dt} nblk{dp.plan.nblk}x{dp.plan.n_nblk} buf{dp.plan.nbuf})"
          f" | dgrad {t_d:7.3f} ms {flops/t_d/1e9:7.1f} TF/s | wgrad {t_w:7.3f} ms {flops/t_w/1e9:7.1f} TF/s"
          f" (jobs {wp.plan.n_jobs} split {wp.plan.split})", flush=True)
    tot["fwd"] += t_f * cnt; tot["dgrad"] += t_d * cnt; tot["wgrad"] += t_w * cnt
    del xs, y, dxs
print("weighted totals (ms):", {k: round(v, 2) for k, v in tot.items()})

"""ncu targets the round-1 review asked for: one level-4 launch (conv 480->480 k3 @ 2x8^3), one level-3 launch
(240->240 @ 2x16^3) and one strided launch (convT 60->30 forward, 64^3 -> 128^3), 3 launches each.
Usage: python tools/prof_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops, plan as P
dev = "cuda"


def run(kind, ks, stride, cin, cout, grid, in_dims, out_dims):
    pl = P.make_conv_plan(kind, ks, stride, [cin], [cout], grid[1], grid)
    dp = ops.DeviceConvPlan(pl, dev)
    x = torch.randn(grid[0], *in_dims, P.pad_channels(cin), device=dev).to(torch.bfloat16)
    w = (torch.randn(cin, cout, 3, 3, 3, device=dev) if kind.startswith("convT") else torch.randn(cout, cin, ks, ks, ks, device=dev)) * 0.05
    out = torch.empty(grid[0], *out_dims, P.pad_channels(cout), device=dev, dtype=torch.bfloat16)
    st = torch.zeros(grid[0], P.pad_channels(cout), 2, device=dev, dtype=torch.float64)
    pw = dp.packed_weight(w)
    for _ in range(3):
        ops.conv_gemm(dp, [x], pw, [out], grid, stats=st, zero_last=kind == 'convT_fwd')
    torch.cuda.synchronize()


run("conv_fwd", 3, 1, 480, 480, (2, 8, 8, 8), (8, 8, 8), (8, 8, 8))
run("conv_fwd", 3, 1, 240, 240, (2, 16, 16, 16), (16, 16, 16), (16, 16, 16))
run("convT_fwd", 3, 2, 60, 30, (2, 64, 64, 64), (64, 64, 64), (128, 128, 128))
ops.check_device_errors()
print("ok")

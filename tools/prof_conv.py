"""ncu targets: one level-0 conv (30->30 k3 @ 2x128^3) forward and its weight gradient, one level-0 InstanceNorm
forward / backward pair, launched 3 times each.  Usage: python tools/prof_conv.py [fwd|wgrad|norm|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops, plan as P
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
grid = (2, 128, 128, 128)
pl = P.make_conv_plan("conv_fwd", 3, 1, [30], [30], 128, grid)
dp = ops.DeviceConvPlan(pl, dev)
x = torch.randn(2, 128, 128, 128, 32, device=dev).to(torch.bfloat16)
w = torch.randn(30, 30, 3, 3, 3, device=dev) * 0.1
out = torch.empty_like(x)
g = torch.empty_like(x)
dy = torch.empty_like(x)
st = torch.zeros(2, 32, 2, device=dev, dtype=torch.float64)
table = torch.empty(2, 32, 2, device=dev)
sums = torch.zeros(2, 32, 2, device=dev, dtype=torch.float64)
wpl = ops.DeviceWgradPlan(P.make_wgrad_plan("conv", 3, 1, [30], 30, grid, 148), dev)
dw = torch.zeros(wpl.plan.dw_numel + 1, device=dev)
for _ in range(3):
    if which in ("fwd", "all"):
        st.zero_()
        ops.conv_gemm(dp, [x], dp.packed_weight(w), [out], grid, stats=st)
    if which in ("wgrad", "all"):
        ops.wgrad_gemm(wpl, [x], out, dw, grid)
    if which in ("norm", "all"):
        ops.in_finalize(st, None, table, 128 ** 3)
        ops.in_apply(out, None, g, table)
        ops.in_bwd_reduce(x, None, g, out, dy, table, sums)
        ops.in_bwd_apply(dy, out, g, table, sums)
torch.cuda.synchronize()
ops.check_device_errors()
print("ok")

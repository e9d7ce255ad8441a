"""ncu target: the level-2 weight gradient (120->120 k3 @ 2x32^3), the swizzled-operand path.  python tools/prof_wgrad_l2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops, plan as P
dev = "cuda"
grid = (2, 32, 32, 32)
x = torch.randn(2, 32, 32, 32, 128, device=dev).to(torch.bfloat16)
dy = torch.randn(2, 32, 32, 32, 128, device=dev).to(torch.bfloat16)
wpl = ops.DeviceWgradPlan(P.make_wgrad_plan("conv", 3, 1, [120], 120, grid, 148), dev)
dw = torch.zeros(wpl.plan.dw_numel + 1, device=dev)
for _ in range(3):
    ops.wgrad_gemm(wpl, [x], dy, dw, grid)
torch.cuda.synchronize()
ops.check_device_errors()
print("ok")

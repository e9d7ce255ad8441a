"""One level-0 conv (30->30 k3 @ 2x128^3) launched 3 times: the ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops, plan as P
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
pl = P.make_conv_plan("conv_fwd", 3, 1, [30], [30], 128)
dp = ops.DeviceConvPlan(pl, dev)
x = torch.randn(2, 128, 128, 128, 32, device=dev).to(torch.bfloat16)
w = torch.randn(30, 30, 3, 3, 3, device=dev) * 0.1
out = torch.empty_like(x)
st = torch.zeros(2, 32, 2, device=dev, dtype=torch.float64)
wpl = ops.DeviceWgradPlan(P.make_wgrad_plan("conv", 3, 1, [30], 30, (2, 128, 128, 128), 148), dev)
dw = torch.zeros(wpl.plan.dw_numel + 1, device=dev)
for _ in range(3):
    if which == "fwd":
        ops.conv_gemm(dp, [x], dp.packed_weight(w), [out], (2, 128, 128, 128), stats=st)
    else:
        ops.wgrad_gemm(wpl, [x], out, dw, (2, 128, 128, 128))
torch.cuda.synchronize()
ops.check_device_errors()
print("ok")

"""ncu target: the stem kernels at cfg-2 size (2 x 1 x 128^3 -> 32 channels), mma.sync path.  python tools/prof_stem.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops
dev = "cuda"
x = torch.randn(2, 1, 128, 128, 128, device=dev)
w0 = torch.randn(1, 27, 32, device=dev) * 0.2
b0 = torch.randn(32, device=dev)
out = torch.empty(2, 128, 128, 128, 32, device=dev, dtype=torch.bfloat16)
dy = torch.randn(2, 128, 128, 128, 32, device=dev).to(torch.bfloat16)
dw = torch.zeros(1, 28, 32, device=dev)
for i in range(3):
    if i == 2:
        torch.cuda.cudart().cudaProfilerStart()
    ops.stem_fwd(x, w0, b0, out)
    ops.stem_wgrad(x, dy, dw)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
ops.check_device_errors()
print("ok")

"""Cycle accounting of the conv_gemm MMA warp (CTA 0) for one layer: total, and the time spent waiting for a free
accumulator buffer / an A slab / a weight stage.  Usage: python tools/conv_times.py [L0|L1|...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops, plan as P
dev = "cuda"
layers = {"L0": ([30], 30, 128), "L0cat": ([30, 30], 30, 128), "L1": ([60], 60, 64), "L2": ([120], 120, 32), "L3": ([240], 240, 16), "L4": ([480], 480, 8)}
name = sys.argv[1] if len(sys.argv) > 1 else "L0"
cins, cout, e = layers[name]
grid = (2, e, e, e)
pl = P.make_conv_plan("conv_fwd", 3, 1, cins, [cout], e, grid)
dp = ops.DeviceConvPlan(pl, dev)
xs = [(torch.randn(2, e, e, e, P.pad_channels(c), device=dev) * 0.5).to(torch.bfloat16) for c in cins]
w = torch.randn(cout, sum(cins), 3, 3, 3, device=dev) * 0.05
out = torch.empty(2, e, e, e, P.pad_channels(cout), device=dev, dtype=torch.bfloat16)
st = torch.zeros(2, out.shape[-1], 2, device=dev, dtype=torch.float64)
ops.DBG_OUT = torch.zeros(8, dtype=torch.int64, device=dev)
wp = dp.packed_weight(w)
for _ in range(3):
    ops.conv_gemm(dp, xs, wp, [out], grid, stats=st)
torch.cuda.synchronize()
t = ops.DBG_OUT.tolist()
n = max(1, t[4])
print(f"{name}: Dt{pl.Dt} G{pl.G} nblk{pl.nblk}x{pl.n_nblk} dense={pl.dense} wT{pl.wT}: items {t[4]}, per item: total {t[0]/n:.0f} clk, "
      f"wait acc_empty {t[1]/n:.0f}, wait a_full {t[2]/n:.0f}, wait w_full {t[3]/n:.0f}, issue+other {(t[0]-t[1]-t[2]-t[3])/n:.0f}")

"""ncu target: the case-level kernels around the window loop at cfg-4 size -- zoom + clip + z-score of a 512x512x128 case
onto 512x512x256, the label map zoomed back, connected components of a blob mask.  python tools/prof_prepost.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops, transform as T
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
raw = torch.randn(512, 512, 128, 1, device=dev, generator=g)
up = torch.empty(512, 512, 256, 1, device=dev)
table = T.normalize_table({"mean": 0.1, "std": 0.9, "pct_00_5": -2.0, "pct_99_5": 2.0})
smooth = torch.nn.functional.interpolate(torch.rand(1, 1, 64, 64, 32, device=dev, generator=g), size=(512, 512, 256),
                                         mode="trilinear")[0, 0]
lab = ((smooth > 0.5).to(torch.uint8) + (smooth > 0.58).to(torch.uint8)).contiguous()      # blocky 3-class label map
down = torch.empty(512, 512, 128, dtype=torch.uint8, device=dev)
mask = (smooth > 0.55).to(torch.uint8).contiguous()
torch.cuda.synchronize()
for i in range(2):
    if i == 1:
        torch.cuda.cudart().cudaProfilerStart()
    ops.PROFILE = []
    T.rescale_device(raw, (1.0, 1.0, 2.0), multi_class=True, out=up, norm=table)
    T.rescale_device(lab, (1.0, 1.0, 0.5), is_label=True, num_classes=3, out=down)
    labels, roots, stats = ops.connected_components(mask)
    torch.cuda.synchronize()
    print({n: round(a.elapsed_time(b), 4) for n, _, a, b, _ in ops.PROFILE}, "components", roots.numel())
    ops.PROFILE = None
torch.cuda.cudart().cudaProfilerStop()
ops.check_device_errors()
print("ok")

"""Back-to-back timing (CUDA graph of R launches, so no host launch overhead) of the InstanceNorm kernels at the five
level shapes of cfg-2: in_apply (+skip), in_bwd_reduce (+out), in_bwd_apply (plain / zero_last + dsum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops

dev = "cuda"
R = 20
shapes = [(2, 128, 128, 128, 32), (2, 64, 64, 64, 64), (2, 32, 32, 32, 128), (2, 16, 16, 16, 240), (2, 8, 8, 8, 480)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_time(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R):
            fn()
    ts = []
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / R * 1e3)
    return min(ts)


for shp in shapes:
    n, d, h, w, cp = shp
    y = torch.randn(shp, device=dev).to(torch.bfloat16)
    s = torch.randn(shp, device=dev).to(torch.bfloat16)
    o = torch.empty_like(y); g = torch.empty_like(y); dy = torch.empty_like(y)
    table = torch.rand(n, cp, 2, device=dev) + 0.5
    sums = torch.zeros(n, cp, 2, device=dev, dtype=torch.float64)
    dsum = torch.zeros(cp, device=dev, dtype=torch.float64)
    mb = y.numel() * 2 / 1e6
    res = {
        "apply": (graph_time(lambda: ops.in_apply(y, None, o, table)), 2),
        "apply+skip": (graph_time(lambda: ops.in_apply(y, s, o, table)), 3),
        "bwd_reduce": (graph_time(lambda: ops.in_bwd_reduce(s, None, None, y, g, table, sums)), 3),
        "bwd_reduce+out": (graph_time(lambda: ops.in_bwd_reduce(s, None, o, y, g, table, sums)), 4),
        "bwd_apply": (graph_time(lambda: ops.in_bwd_apply(g, y, dy, table, sums)), 3),
        "bwd_apply zl+dsum": (graph_time(lambda: ops.in_bwd_apply(g, y, dy, table, sums, dsum, True)), 3),
    }
    print(f"{shp}: " + "  ".join(f"{k} {t:6.1f} us ({mb * f / t * 1e-3 * 1e3 / 1e3:5.2f} TB/s)" for k, (t, f) in res.items()), flush=True)

"""Does the mma.sync stem forward depend on how the batch is partitioned?  (DESIGN.md open question of round 1: 4e-3 of
noise between two ranks x 2 samples and one process x 4 samples.)  Runs the kernel on N = 4 and on the two halves and
compares bit for bit, and against the CUDA-core kernel (other process: the switch is a static getenv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from unet3d_b200 import ops, plan as P

dev = "cuda"
torch.manual_seed(0)
for dims in ((16, 16, 16), (16, 20, 24), (32, 32, 32)):
    x = torch.randn(4, 1, *dims, device=dev)
    w = torch.randn(1, 27, 32, device=dev) * 0.3
    w[:, :, 30:] = 0
    b = torch.randn(32, device=dev)
    b[30:] = 0
    def run(xx):
        out = torch.full((xx.shape[0], *dims, 32), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.stem_fwd(xx.contiguous(), w, b, out)
        torch.cuda.synchronize()
        return out
    full = run(x)
    halves = torch.cat([run(x[:2]), run(x[2:])])
    again = run(x)
    ref = torch.nn.functional.conv3d(x, w[0, :, :30].t().reshape(30, 1, 3, 3, 3), b[:30], padding=1)
    got = full[..., :30].permute(0, 4, 1, 2, 3).float()
    print(dims, "mma" if os.environ.get("U3D_STEM_FWD_MMA") else "fma", "N=4 vs 2+2 identical:", torch.equal(full, halves),
          "| run twice identical:", torch.equal(full, again), "| nan:", bool(torch.isnan(full.float()).any()),
          "| rel vs fp32 conv:", float((got - ref).norm() / ref.norm()),
          "| differing elements:", int((full != halves).sum()))

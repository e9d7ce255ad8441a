"""Bisect a gradient difference between the eager data-parallel step and GraphedTrainStep's one-graph mode.
torchrun --nproc-per-node 2 tools/diag/dp_graph_diag.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import unet3d_b200  # noqa: E402
from unet3d_b200 import parallel  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)


class EvalStep(unet3d_b200.GraphedTrainStep):
    def _fwd_bwd(self, image, label):
        self.optimizer.zero_grad(set_to_none=True)
        logits = self.model(image)
        loss = self.loss_fn(logits, label)
        loss.backward()
        return loss, logits


def run(kind, steps=5, **kw):
    torch.manual_seed(7)
    m = unet3d_b200.ResUnet3D(num_pool=2, num_features=16, out_channels=3).to(dev).eval()
    o = torch.optim.SGD(m.parameters(), lr=0.0)
    lf = unet3d_b200.DiceLoss()
    gg = torch.Generator().manual_seed(50 + rank)
    xb = torch.randn(2, 1, 32, 32, 32, generator=gg).to(dev)
    yb = torch.randint(0, 3, (2, 32, 32, 32), generator=gg).to(dev)
    stepper = EvalStep(m, lf, o, warmup=2, **kw) if kind == "graph" else None
    if kind == "eager-prescale":
        o.zero_grad(set_to_none=True); lf(m(xb), yb).backward(); parallel.all_reduce_gradients(m)
        parallel.prescale_gradients(m)
        if kw.get("overlap"):
            parallel.overlap_gradient_all_reduce(m)
    out = []
    for _ in range(steps):
        if stepper is not None:
            stepper(xb, yb)
        else:
            o.zero_grad(set_to_none=True)
            lf(m(xb), yb).backward()
            parallel.all_reduce_gradients(m)
            o.step()
        torch.cuda.synchronize()
        out.append({n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None})
    return out, (stepper.mode if stepper else kind)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def report(tag, got, ref):
    errs = sorted(((rel(got[n], ref[n]), n) for n in ref if n in got), reverse=True)
    if rank == 0:
        print(f"{tag:46s} " + "  ".join(f"{e:.1e} {n.replace('net.', '')}" for e, n in errs[:3]), flush=True)


base, _ = run("eager")
for i in range(1, 5):
    report(f"eager step {i} vs eager step 0", base[i], base[0])
again, _ = run("eager")
report("eager (second model) vs eager", again[-1], base[-1])
for kw in (dict(), dict(overlap=True)):
    got, mode = run("eager-prescale", **kw)
    report(f"eager prescale {kw}", got[-1], base[-1])
for kw in (dict(capture_collectives=False), dict(overlap=False), dict()):
    got, mode = run("graph", **kw)
    for i in range(5):
        report(f"graph {kw} [{mode}] step {i}", got[i], base[-1])
    del got
    import gc; gc.collect(); torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
os._exit(0)

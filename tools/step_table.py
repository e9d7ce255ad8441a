"""Per-launch table of one training step of cfg-2 (or --patch P / --net 5): every library launch with its layer shape,
plan parameters, CUDA-event time and algorithmic TFLOP/s or GB/s, sorted by time; plus totals per kernel and per shape.
Eager launches (event pair around each): the times are per kernel, the step total is larger than a graph replay."""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
from unet3d_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--patch", type=int, default=128)
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--infer", action="store_true", help="a no-grad forward of --batch windows instead of a training step")
a = ap.parse_args()
dev = "cuda"
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(4, 30, 1, 3).to(dev)
model.precision = a.precision
model.train(not a.infer)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
loss_fn = unet3d_b200.DiceLoss()
x = torch.randn(a.batch, 1, a.patch, a.patch, a.patch, device=dev)
y = torch.randint(0, 3, (a.batch, a.patch, a.patch, a.patch), device=dev)
ops.REAL_CHANNELS = {32: 30, 64: 60, 128: 120}


def step():
    if a.infer:
        with torch.no_grad():
            model(x)
        return
    opt.zero_grad(set_to_none=True)
    loss_fn(model(x), y).backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
acc = collections.OrderedDict()
for r in range(a.reps):
    ops.PROFILE, ops.PROFILE_DETAIL = [], []
    step()
    torch.cuda.synchronize()
    for i, ((name, flops, e0, e1, nbytes), det) in enumerate(zip(ops.PROFILE, ops.PROFILE_DETAIL)):
        k = (i, name, det)
        t = acc.setdefault(k, [0.0, flops, nbytes])
        t[0] += e0.elapsed_time(e1) / a.reps
ops.PROFILE = ops.PROFILE_DETAIL = None
rows = [(t[0], name, det, t[1], t[2], i) for (i, name, det), t in acc.items()]
tot = sum(r[0] for r in rows)
print(f"{len(rows)} library launches, {tot:.3f} ms (event-timed, eager)")
by_kernel = collections.defaultdict(lambda: [0.0, 0])
by_shape = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for ms, name, det, fl, nb, i in rows:
    by_kernel[name][0] += ms; by_kernel[name][1] += 1
    s = by_shape[(name, det)]
    s[0] += ms; s[1] += 1; s[2] += fl; s[3] += nb
print("\n== per kernel")
for name, (ms, n) in sorted(by_kernel.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:8.3f} ms  x{n:3d}  {name}")
print("\n== per (kernel, shape), sorted by total time")
for (name, det), (ms, n, fl, nb) in sorted(by_shape.items(), key=lambda kv: -kv[1][0]):
    rate = f"{fl / ms / 1e9:7.1f} TF/s" if fl > 0 else (f"{nb / ms / 1e6:7.1f} GB/s" if nb > 0 else "")
    print(f"{ms:8.3f} ms  x{n:2d}  ({ms / n * 1e3:7.1f} us each) {rate}  {name[:18]:18s} {det}")

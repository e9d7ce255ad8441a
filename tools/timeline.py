"""Kernel-time breakdown of one training step with torch.profiler (CUPTI): per-kernel totals, GPU-busy time
vs wall time (host-bound gaps).  Usage: python tools/timeline.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import unet3d_b200

dev = "cuda"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(out_channels=3).to(dev).train()
model.precision = precision
loss_fn = unet3d_b200.DiceLoss()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
x = torch.randn(2, 1, 128, 128, 128, device=dev)
y = torch.randint(0, 3, (2, 128, 128, 128), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = loss_fn(model(x), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / steps
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
rows = []
tot = 0.0
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0)
    if t > 0 and ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        rows.append((t / steps / 1e3, ev.count // steps, ev.key))
        tot += t / steps / 1e3
rows.sort(reverse=True)
print(f"precision {precision}: wall {wall:.2f} ms/step (unprofiled), sum of kernel time {tot:.2f} ms/step")
for t, n, k in rows[:28]:
    print(f"{t:8.3f} ms  x{n:4d}  {k[:90]}")

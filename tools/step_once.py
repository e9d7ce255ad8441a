"""One training step of cfg-2 inside a cudaProfilerStart/Stop window (after warm-up): the ncu launch-list target.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python tools/step_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200

dev = "cuda"
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(out_channels=3).to(dev).train()
model.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
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
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")

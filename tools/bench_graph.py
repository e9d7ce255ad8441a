"""cfg-2 training step eager vs replayed as one CUDA graph (GraphedTrainStep).  python tools/bench_graph.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200

dev = "cuda"
torch.manual_seed(0)
x = torch.randn(2, 1, 128, 128, 128, device=dev)
y = torch.randint(0, 3, (2, 128, 128, 128), device=dev)
for mode in ("eager", "graph"):
    torch.manual_seed(0)
    model = unet3d_b200.ResUnet3D(out_channels=3).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    loss_fn = unet3d_b200.DiceLoss()
    if mode == "graph":
        stepper = unet3d_b200.GraphedTrainStep(model, loss_fn, opt, warmup=3)
        step = lambda: stepper(x, y)[0]
    else:
        def step():
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(model(x), y)
            loss.backward()
            opt.step()
            return loss
    for _ in range(5):
        l = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        l = step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{mode}: {e0.elapsed_time(e1) / 10:.2f} ms/step, loss {l.item():.4f}", flush=True)
    del model, opt
    torch.cuda.empty_cache()

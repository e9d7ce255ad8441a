"""Gradient reproducibility probe: the same shard must give the same gradients whatever was computed in between
(other batch sizes re-plan the layers and rebuild the gradient arenas).  python tools/det_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, unet3d_b200
dev = "cuda"
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3).to(dev).eval()
g = torch.Generator().manual_seed(5)
x_all = torch.randn(4, 1, 16, 16, 16, generator=g).to(dev)
y_all = torch.randint(0, 3, (4, 16, 16, 16), generator=g).to(dev)
loss_fn = unet3d_b200.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1)


def grads(x, y):
    model.zero_grad(set_to_none=True)
    loss_fn(model(x), y).backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def worst(a, b):
    return sorted(((rel(a[n], b[n]), n) for n in a), reverse=True)[:2]


a1 = grads(x_all[:2], y_all[:2])
a2 = grads(x_all[:2], y_all[:2])
print("shard 0 twice                :", worst(a2, a1))
b1 = grads(x_all, y_all)
a3 = grads(x_all[:2], y_all[:2])
print("shard 0 after a batch of 4   :", worst(a3, a1))
b2 = grads(x_all, y_all)
print("batch of 4 twice             :", worst(b2, b1))
c1 = grads(x_all[2:], y_all[2:])
a4 = grads(x_all[:2], y_all[:2])
print("shard 0 after shard 1        :", worst(a4, a1))

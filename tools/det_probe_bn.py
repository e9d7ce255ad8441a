"""Reproducibility probe for the BatchNorm variant: gradients of a batch must not depend on what the engine computed before
(the 2-GPU check showed one process's single-process reference off by 0.95 after a pass with another batch size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, unet3d_b200
dev = "cuda"
torch.manual_seed(1)
bn = unet3d_b200.ResAttrBNUnet3D(num_pool=1, num_features=8, out_channels=3).to(dev).train()
for mod in bn.modules():
    if isinstance(mod, torch.nn.Dropout3d):
        mod.p = 0.0
sd0 = {k: v.clone() for k, v in bn.state_dict().items()}
g = torch.Generator().manual_seed(5)
x_all = torch.randn(4, 1, 16, 16, 16, generator=g).to(dev)
y_all = torch.randint(0, 3, (4, 16, 16, 16), generator=g).to(dev)
loss_fn = unet3d_b200.DiceLoss()


def grads(x, y):
    bn.load_state_dict(sd0)
    bn.zero_grad(set_to_none=True)
    loss_fn(bn(x), y).backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in bn.named_parameters() if p.grad is not None}


def worst(a, b):
    big = max(v.norm() for v in b.values())
    return sorted((((a[n] - b[n]).norm() / b[n].norm().clamp_min(1e-20)).item(), n) for n in b if b[n].norm() > 1e-6 * big)[-2:]


b1 = grads(x_all, y_all)
b2 = grads(x_all, y_all)
print("batch 4 twice (fresh engine)   :", worst(b2, b1))
for which in (slice(0, 2), slice(2, 4)):
    grads(x_all[which], y_all[which])
    b3 = grads(x_all, y_all)
    print(f"batch 4 after shard {which.start // 2}          :", worst(b3, b1))

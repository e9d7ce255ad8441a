"""What does PyTorch's own mixed precision do on the parity test?  Runs the oracle's functional network
(cuDNN / ATen on the GPU) in fp32 and under torch.autocast(bf16 / fp16) on the default net at 1x32^3 and prints the
same numbers tests/test_model_gpu.py checks for this repo: logits rel-L2 / argmax agreement vs fp32 and per-layer
gradient rel-L2.  Context for the tolerances in DESIGN.md (not a test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import unet3d_b200
from oracle import unet3d_oracle as O

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(out_channels=3)
sd0 = {k: v.clone().to(dev) for k, v in model.state_dict().items()}
x = torch.randn(1, 1, 32, 32, 32, generator=torch.Generator().manual_seed(1234)).to(dev)
y = torch.randint(0, 3, (1, 32, 32, 32), generator=torch.Generator().manual_seed(4321)).to(dev)


def run(dtype):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    if dtype is None:
        lg = O.resunet3d_forward(sd, x)
    else:
        with torch.autocast("cuda", dtype=dtype):
            lg = O.resunet3d_forward(sd, x)
    lg = lg.float()
    p = torch.softmax(lg, 1)
    g = torch.nn.functional.one_hot(y, 3).permute(0, 4, 1, 2, 3).float()
    tp = (p * g).sum((0, 2, 3, 4)); fn = ((1 - p) * g).sum((0, 2, 3, 4)); fp = (p * (1 - g)).sum((0, 2, 3, 4))
    loss = (1 - (tp + 1e-7) / (tp + 0.5 * fn + 0.5 * fp + 1e-7)).mean()
    loss.backward()
    return lg.detach(), {k: v.grad for k, v in sd.items()}


ref, gref = run(None)
for name, dt in (("bf16 autocast", torch.bfloat16), ("fp16 autocast", torch.float16)):
    lg, gr = run(dt)
    rel = ((lg - ref).norm() / ref.norm()).item()
    agree = (lg.argmax(1) == ref.argmax(1)).float().mean().item()
    errs = []
    for k, g in gr.items():
        if g is None or gref[k] is None or (k.endswith("bias") and ("conv1" in k or "conv2" in k)):
            continue
        errs.append((((g.float() - gref[k]).norm() / gref[k].norm().clamp_min(1e-30)).item(), k))
    errs.sort()
    print(f"{name}: logits rel-L2 {rel:.3e}, argmax agreement {agree:.5f}; grad rel-L2 median {errs[len(errs)//2][0]:.3f} "
          f"max {errs[-1][0]:.3f} ({errs[-1][1]}), net.conv.weight "
          f"{[e for e, k in errs if k == 'net.conv.weight'][0]:.3f}, decode_blocks.0.conv2.weight "
          f"{[e for e, k in errs if k == 'net.decode_blocks.0.conv2.weight'][0]:.3f}")

"""Does any kernel of the training step read memory it (or a producer) never wrote?  Run the same step on the same
weights with the caching allocator's free memory pre-filled with different byte patterns; every gradient and every
captured intermediate must be bit-identical.   python tools/diag/stale_memory_diag.py [features] [size]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import unet3d_b200  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
torch.manual_seed(7)
m = unet3d_b200.ResUnet3D(num_pool=2, num_features=F, out_channels=3).to(dev).eval()
lf = unet3d_b200.DiceLoss()
gg = torch.Generator().manual_seed(50)
xb = torch.randn(2, 1, S, S, S, generator=gg).to(dev)
yb = torch.randint(0, 3, (2, S, S, S), generator=gg).to(dev)


def poison(byte):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    big = torch.empty(6 << 30, dtype=torch.uint8, device=dev).fill_(byte)
    small = [torch.empty(256 << 10, dtype=torch.uint8, device=dev).fill_(byte) for _ in range(64)]
    tiny = [torch.empty(2048, dtype=torch.uint8, device=dev).fill_(byte) for _ in range(512)]
    torch.cuda.synchronize()
    del big, small, tiny


def step(byte, capture):
    poison(byte)
    m.zero_grad(set_to_none=True)
    eng = m.net._engine
    if capture:
        eng.capture = []
    loss = lf(m(xb), yb)
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    cap = None
    if capture:
        cap = []
        for e in eng.capture:
            for k, v in e.items():
                if torch.is_tensor(v):
                    cap.append((e["key"], k, v.detach().clone()))
                elif isinstance(v, (list, tuple)) and v and torch.is_tensor(v[0]):
                    cap += [((e["key"]), f"{k}[{i}]", t.detach().clone()) for i, t in enumerate(v)]
                elif isinstance(v, dict):
                    cap += [(e["key"], f"{k}.{kk}", t.detach().clone()) for kk, t in v.items() if torch.is_tensor(t)]
        eng.capture = None
    return float(loss), grads, cap


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


lf(m(xb), yb).backward()          # plans, tables
for capture in (False, True):
    l0, g0, c0 = step(0x00, capture)
    for byte in (0x00, 0x70, 0x3c):
        l1, g1, c1 = step(byte, capture)
        bad = [(rel(g1[n], g0[n]), n) for n in g0 if not torch.equal(g1[n], g0[n])]
        print(f"capture={capture} fill=0x{byte:02x}: loss {l1:.7f} vs {l0:.7f}; {len(bad)} of {len(g0)} gradients differ "
              + "  ".join(f"{e:.1e} {n}" for e, n in sorted(bad, reverse=True)[:6]), flush=True)
        if capture and byte != 0:
            for (k0, n0, t0), (k1, n1, t1) in zip(c0, c1):
                assert k0 == k1 and n0 == n1
                if t0.shape != t1.shape or not torch.equal(t0, t1):
                    t0f, t1f = t0.float(), t1.float()
                    real = None
                    if t0.dim() == 5 and t0.shape[-1] % 16 == 0 and t0.shape[-1] >= F:
                        real = rel(t1f[..., :F], t0f[..., :F])
                    print(f"    {k0} {n0} {tuple(t0.shape)} differs: rel {rel(t1f, t0f):.2e}"
                          + (f" (real channels only: {real:.2e})" if real is not None else ""), flush=True)

"""Diagnostic for tests/test_block_parity_gpu.py: where inside a residual block does the CUDA backward leave the
stage-forced oracle block?  Prints rel-L2 of every backward intermediate (g2, dy2, da1, g1, dy1, d inputs), the
InstanceNorm tables against statistics of the stored tensors, and how concentrated the error is."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import unet3d_b200
from oracle import unet3d_oracle as O, bf16_model as Q
from test_block_parity_gpu import _model_and_batch, _captured_step, ncdhw, rel

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["dec/0"]
force_tab = len(sys.argv) > 3 and sys.argv[3] == "tab"
dt = torch.float16 if prec == "fp16" else torch.bfloat16
model, x, y = _model_and_batch(lambda: unet3d_b200.ResUnet3D(4, 30, 1, 3), (1, 1, 32, 32, 32), prec)
cap, gscale, _ = _captured_step(model, x, y, unet3d_b200.DiceLoss(), train=False)
names = {m: n for n, m in model.named_modules()}
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
q = lambda t: t.to(dt).float()
force = Q._force
for e in cap:
    tag = "/".join(str(v) for v in e["key"])
    if e["kind"] != "res" or tag not in which:
        continue
    blk = e["blk"]; pre = names[blk] + "."
    cin, cout, stride = blk.in_channels, blk.out_channels, blk.stride
    gpu = {k: ncdhw(v, cout) for k, v in e["fwd"].items()}
    tabs = {k: v[:, :cout].float().cpu() for k, v in e["tab"].items()}       # (N, C, 2): mean, scale
    xin = torch.cat([ncdhw(t, c) for t, c in zip(e["inputs"], e["in_C"])], 1).requires_grad_(True)
    w1, w2 = q(sd[pre + "conv1.weight"]), q(sd[pre + "conv2.weight"])
    skip = q(F.conv3d(xin, q(sd[pre + "skip_conv.weight"]), sd[pre + "skip_conv.bias"], stride=stride)) if blk.uses_skip_conv else xin
    y1 = force(q(F.conv3d(xin, w1, None, stride=stride, padding=1)), gpu["y1"]); y1.retain_grad()
    n1 = O._inorm(y1)
    def tab_norm(yv, t):
        return (yv - t[:, :, 0].view(1, -1, 1, 1, 1)) * t[:, :, 1].view(1, -1, 1, 1, 1)
    print(f"[{prec}] block {tag}: cin {cin} cout {cout} stride {stride} voxels {y1[0,0].numel()} gscale {gscale} force_tab {force_tab}")
    print(f"  norm1 from the CUDA table vs instance_norm(stored y1): rel {rel(tab_norm(gpu['y1'], tabs['t1']), n1.detach()):.3e}; "
          f"sign mismatches {int(((tab_norm(gpu['y1'], tabs['t1']) > 0) != (n1.detach() > 0)).sum())} of {n1.numel()}")
    if force_tab:
        n1 = force(n1, tab_norm(gpu["y1"], tabs["t1"]))
    n1.retain_grad()
    a1 = force(q(O._lrelu(n1)), gpu["a1"]); a1.retain_grad()
    y2 = force(q(F.conv3d(a1, w2, None, padding=1)), gpu["y2"]); y2.retain_grad()
    pre_act = O._inorm(y2) + skip; pre_act.retain_grad()
    slope = torch.where(gpu["out"] > 0, torch.ones_like(pre_act), torch.full_like(pre_act, 0.01))
    out = pre_act * slope
    dout = ncdhw(e["dout"], cout) + (ncdhw(e["dout2"], cout) if e["dout2"] is not None else 0)
    dout = dout / gscale
    out.backward(dout)
    b = e["bwd"]
    for nm, got, want in (("g2", b["g2"], pre_act.grad), ("dy2", b["dy2"], y2.grad), ("da1", b["da1"], a1.grad),
                          ("g1", b["g1"], n1.grad), ("dy1", b["dy1"], y1.grad)):
        gg = ncdhw(got, cout) / gscale
        d = (gg - want)
        top = d.abs().flatten().topk(10)
        print(f"  bwd {nm:4s} rel {rel(gg, want):.3e}   |want| max {want.abs().max():.3e} rms {want.pow(2).mean().sqrt():.3e}; "
              f"err energy in top-10 elements {float((top.values ** 2).sum() / (d ** 2).sum()):.3f}")
        if nm in ("g1", "dy2"):
            for i in top.indices[:4]:
                i = int(i)
                print(f"       elem {i}: got {gg.flatten()[i]:.4e} want {want.flatten()[i]:.4e}")
    off = 0
    for i, (t, c) in enumerate(zip(e["dins"], e["in_C"])):
        gg = ncdhw(t, c) / gscale
        want = xin.grad[:, off:off + c]
        print(f"  d(input {i}) rel {rel(gg, want):.3e}")
        off += c

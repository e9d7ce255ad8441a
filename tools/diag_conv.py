"""GPU diagnostic for the shifted-GEMM kernel: runs a ladder of cases from trivial to full and
prints where (which rows / channels / taps) a mismatch sits.  Not a test; used while bringing the
kernel up through gpurun."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import unet3d_b200  # noqa
from unet3d_b200 import ops, plan as P

dev = "cuda"
torch.manual_seed(0)


def to_ndhwc(x, cp):
    n, c = x.shape[:2]
    out = torch.zeros(n, *x.shape[2:], cp, device=x.device, dtype=torch.bfloat16)
    out[..., :c] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return out


def run(kind, ks, stride, cins, couts, dims, bias=False, stats=False, addend=False, label=""):
    N, D, H, W = dims
    bf = lambda t: t.to(torch.bfloat16).float()
    if kind == "conv_fwd":
        xs = [bf(torch.randn(N, c, D, H, W, device=dev)) for c in cins]
        w = bf(torch.randn(couts[0], sum(cins), ks, ks, ks, device=dev) * 0.1)
        b = torch.randn(couts[0], device=dev) if bias else None
        ref = [F.conv3d(torch.cat(xs, 1), w, b, stride=stride, padding=ks // 2)]
        grid = (N, D // stride, H // stride, W // stride)
        inputs = [to_ndhwc(x, P.pad_channels(x.shape[1])) for x in xs]
        depth = D // stride
    elif kind == "conv_dgrad":
        cin_tot = sum(couts)
        x = torch.randn(N, cin_tot, D, H, W, device=dev, requires_grad=True)
        w = bf(torch.randn(cins[0], cin_tot, ks, ks, ks, device=dev) * 0.1)
        y = F.conv3d(x, w, None, stride=stride, padding=ks // 2)
        dy = bf(torch.randn_like(y))
        y.backward(dy)
        ref, off = [], 0
        for c in couts:
            ref.append(x.grad[:, off:off + c]); off += c
        grid = (N, D // stride, H // stride, W // stride)
        inputs = [to_ndhwc(dy, P.pad_channels(cins[0]))]
        b = None
        depth = D // stride
    elif kind == "convT_fwd":
        x = bf(torch.randn(N, cins[0], D, H, W, device=dev))
        w = bf(torch.randn(cins[0], couts[0], 3, 3, 3, device=dev) * 0.1)
        b = torch.randn(couts[0], device=dev) if bias else None
        ref = [F.pad(F.conv_transpose3d(x, w, b, stride=2, padding=1), (0, 1, 0, 1, 0, 1))]
        grid = (N, D, H, W)
        inputs = [to_ndhwc(x, P.pad_channels(cins[0]))]
        depth = D
    else:
        raise ValueError(kind)
    pl = P.make_conv_plan(kind, ks, stride, cins, couts, depth)
    dp = ops.DeviceConvPlan(pl, dev)
    wp = dp.packed_weight(w)
    bp = dp.packed_bias(b)
    outs = [torch.full((N, *r.shape[2:], P.pad_channels(r.shape[1])), float("nan"), device=dev, dtype=torch.bfloat16)
            for r in ref]
    if kind == "conv_dgrad" and stride == 2 and ks == 1:
        for o in outs:
            o.zero_()
    adds = None
    if addend:
        adds = [torch.zeros_like(o) for o in outs]
        for a_, r_ in zip(adds, ref):
            a_[..., :r_.shape[1]] = torch.randn(*a_.shape[:-1], r_.shape[1], device=dev)
    st = torch.zeros(N, outs[0].shape[-1], 2, device=dev, dtype=torch.float64) if stats else None
    ops.conv_gemm(dp, inputs, wp, outs, grid, bias=bp, addends=adds, stats=st, zero_last=(kind == "convT_fwd"))
    torch.cuda.synchronize()
    ops.check_device_errors()
    ok = True
    for i, (o, r) in enumerate(zip(outs, ref)):
        got = o[..., :r.shape[1]].permute(0, 4, 1, 2, 3).float()
        want = r
        if addend:
            want = want + adds[i][..., :r.shape[1]].permute(0, 4, 1, 2, 3).float()
        if torch.isnan(got).any():
            print(f"   NaN count {torch.isnan(got).sum().item()} of {got.numel()}")
        err = (got - want).norm() / want.norm()
        padmax = o[..., r.shape[1]:].float().abs().max().item() if o.shape[-1] > r.shape[1] else 0.0
        print(f"[{label}] {kind} ks{ks} s{stride} {cins}->{couts} dims{dims} Dt{pl.Dt} G{pl.G} nblk{pl.nblk}x{pl.n_nblk}"
              f" out{i}: rel-L2 {err.item():.3e} padmax {padmax:.2e}")
        if not (err < 1e-2) or padmax != 0.0:
            ok = False
            d = (got - want).abs()
            print("   err by channel :", d.amax(dim=(0, 2, 3, 4))[:12].tolist())
            print("   err by d       :", d.amax(dim=(0, 1, 3, 4)).tolist()[:12])
            print("   err by h       :", d.amax(dim=(0, 1, 2, 4)).tolist()[:24])
            print("   err by w       :", d.amax(dim=(0, 1, 2, 3)).tolist()[:24])
            print("   got[0,0,0,0,:8]", got[0, 0, 0, 0, :8].tolist())
            print("   want           ", want[0, 0, 0, 0, :8].tolist())
    if stats:
        v = outs[0].float()
        s1 = v.sum(dim=(1, 2, 3)).double()
        s2 = (v * v).sum(dim=(1, 2, 3)).double()
        e1 = (st[..., 0] - s1).abs().max().item() / (s1.abs().max().item() + 1e-9)
        e2 = (st[..., 1] - s2).abs().max().item() / (s2.abs().max().item() + 1e-9)
        print(f"   stats rel err sum {e1:.2e} sumsq {e2:.2e}")
        ok = ok and e1 < 2e-2 and e2 < 2e-2
    return ok


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "SMs", _ := ops._lib.lib().unet3d_num_sms())
    cases = [
        ("conv_fwd", 1, 1, [16], [32], (1, 1, 16, 8), {}),
        ("conv_fwd", 3, 1, [16], [32], (1, 1, 16, 8), {}),
        ("conv_fwd", 3, 1, [30], [30], (1, 4, 16, 16), dict(bias=True)),
        ("conv_fwd", 3, 1, [30], [30], (2, 5, 20, 12), dict(bias=True, stats=True)),
        ("conv_fwd", 3, 1, [30, 30], [30], (1, 8, 32, 32), dict(stats=True)),
        ("conv_fwd", 3, 1, [60], [60], (1, 8, 16, 16), dict(addend=True)),
        ("conv_fwd", 3, 1, [120], [120], (1, 4, 16, 16), {}),
        ("conv_fwd", 3, 1, [240], [240], (2, 4, 16, 16), dict(stats=True)),
        ("conv_fwd", 3, 1, [480], [480], (2, 2, 8, 8), dict(stats=True)),
        ("conv_fwd", 3, 1, [240, 240], [240], (1, 4, 16, 16), {}),
        ("conv_fwd", 1, 1, [60, 60], [60], (1, 4, 16, 16), dict(bias=True)),
        ("conv_fwd", 3, 2, [30], [60], (1, 8, 32, 32), dict(stats=True)),
        ("conv_fwd", 1, 2, [30], [60], (1, 8, 32, 32), dict(bias=True)),
        ("conv_dgrad", 3, 1, [30], [30], (1, 4, 16, 16), dict(addend=True)),
        ("conv_dgrad", 3, 1, [30], [30, 30], (1, 4, 16, 16), {}),
        ("conv_dgrad", 1, 1, [60], [60, 60], (1, 4, 16, 16), {}),
        ("conv_dgrad", 3, 2, [60], [30], (1, 8, 32, 32), {}),
        ("conv_dgrad", 1, 2, [60], [30], (1, 8, 32, 32), {}),
        ("convT_fwd", 3, 2, [60], [30], (1, 4, 16, 16), dict(bias=True, stats=True)),
        ("convT_fwd", 3, 2, [480], [240], (1, 2, 4, 4), dict(bias=True)),
    ]
    n_ok = 0
    for i, (kind, ks, s, ci, co, dims, kw) in enumerate(cases):
        try:
            ok = run(kind, ks, s, ci, co, dims, label=str(i), **kw)
        except Exception as e:  # noqa
            print(f"[{i}] EXCEPTION {type(e).__name__}: {e}")
            ok = False
            if "timeout" in str(e) or "CUDA" in str(e):
                break
        n_ok += bool(ok)
    print(f"{n_ok}/{len(cases)} cases ok")
    # quick timing of the level-0 conv at the benchmark shape
    if n_ok == len(cases):
        pl = P.make_conv_plan("conv_fwd", 3, 1, [30], [30], 128)
        dp = ops.DeviceConvPlan(pl, dev)
        x = torch.randn(2, 128, 128, 128, 32, device=dev).to(torch.bfloat16)
        w = torch.randn(30, 30, 3, 3, 3, device=dev) * 0.1
        wp = dp.packed_weight(w)
        out = torch.empty_like(x)
        st = torch.zeros(2, 32, 2, device=dev, dtype=torch.float64)
        for _ in range(3):
            ops.conv_gemm(dp, [x], wp, [out], (2, 128, 128, 128), stats=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv_gemm(dp, [x], wp, [out], (2, 128, 128, 128), stats=st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"level-0 conv 30->30 @2x128^3: {ms:.3f} ms  -> {203.84 / ms:.1f} TFLOP/s algorithmic")
        ops.check_device_errors()

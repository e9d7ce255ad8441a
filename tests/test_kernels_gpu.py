"""GPU parity of the individual kernels (through the C ABI) against fp32 PyTorch on the same inputs."""
import importlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import unet3d_b200  # noqa: E402,F401
from unet3d_b200 import ops, plan as P  # noqa: E402

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False          # the fp32 references must be real fp32
torch.backends.cuda.matmul.allow_tf32 = False
TOL_BF16 = 5e-3     # rel-L2 of one bf16 conv against the fp32 result on the same bf16-rounded inputs


def bf(t):
    return t.to(torch.bfloat16).float()


def to_ndhwc(x, cp=None):
    n, c = x.shape[:2]
    cp = cp or P.pad_channels(c)
    out = torch.zeros(n, *x.shape[2:], cp, device=x.device, dtype=torch.bfloat16)
    out[..., :c] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return out


def from_ndhwc(x, c):
    return x[..., :c].permute(0, 4, 1, 2, 3).float()


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


CONV_CASES = [
    ("conv_fwd", 1, 1, [16], [32], (1, 1, 16, 8)),
    ("conv_fwd", 3, 1, [30], [30], (2, 5, 20, 12)),
    ("conv_fwd", 3, 1, [30, 30], [30], (1, 8, 32, 32)),
    ("conv_fwd", 3, 1, [60], [60], (1, 8, 16, 16)),
    ("conv_fwd", 3, 1, [240], [240], (2, 4, 16, 16)),
    ("conv_fwd", 3, 1, [480], [480], (2, 2, 8, 8)),
    ("conv_fwd", 1, 1, [60, 60], [60], (1, 4, 16, 16)),
    ("conv_fwd", 3, 2, [30], [60], (1, 8, 32, 32)),
    ("conv_fwd", 1, 2, [30], [60], (1, 8, 32, 32)),
    ("conv_fwd", 3, 1, [8], [8], (1, 3, 5, 5)),
    ("conv_dgrad", 3, 1, [30], [30], (1, 4, 16, 16)),
    ("conv_dgrad", 3, 1, [30], [30, 30], (1, 4, 16, 16)),
    ("conv_dgrad", 1, 1, [60], [60, 60], (1, 4, 16, 16)),
    ("conv_dgrad", 3, 2, [60], [30], (1, 8, 32, 32)),
    ("conv_dgrad", 1, 2, [60], [30], (1, 8, 32, 32)),
    ("convT_fwd", 3, 2, [60], [30], (1, 4, 16, 16)),
    ("convT_fwd", 3, 2, [480], [240], (1, 2, 4, 4)),
    ("convT_dgrad", 3, 2, [30], [60], (1, 4, 8, 8)),
]


@pytest.mark.parametrize("kind,ks,stride,cins,couts,dims", CONV_CASES)
def test_conv_gemm(kind, ks, stride, cins, couts, dims):
    torch.manual_seed(1)
    N, D, H, W = dims
    zero_last = False
    b = None
    if kind == "conv_fwd":
        xs = [bf(torch.randn(N, c, D, H, W, device=DEV)) for c in cins]
        w = bf(torch.randn(couts[0], sum(cins), ks, ks, ks, device=DEV) * 0.1)
        b = torch.randn(couts[0], device=DEV)
        ref = [F.conv3d(torch.cat(xs, 1), w, b, stride=stride, padding=ks // 2)]
        grid = (N, D // stride, H // stride, W // stride)
        inputs = [to_ndhwc(x) for x in xs]
    elif kind == "conv_dgrad":
        x = torch.randn(N, sum(couts), D, H, W, device=DEV, requires_grad=True)
        w = bf(torch.randn(cins[0], sum(couts), ks, ks, ks, device=DEV) * 0.1)
        y = F.conv3d(x, w, None, stride=stride, padding=ks // 2)
        dy = bf(torch.randn_like(y))
        y.backward(dy)
        ref, off = [], 0
        for c in couts:
            ref.append(x.grad[:, off:off + c]); off += c
        grid = (N, D // stride, H // stride, W // stride)
        inputs = [to_ndhwc(dy)]
    elif kind == "convT_fwd":
        x = bf(torch.randn(N, cins[0], D, H, W, device=DEV))
        w = bf(torch.randn(cins[0], couts[0], 3, 3, 3, device=DEV) * 0.1)
        b = torch.randn(couts[0], device=DEV)
        ref = [F.pad(F.conv_transpose3d(x, w, b, stride=2, padding=1), (0, 1, 0, 1, 0, 1))]
        grid = (N, D, H, W)
        inputs = [to_ndhwc(x)]
        zero_last = True
    else:   # convT_dgrad: dims = coarse grid; dy on the fine grid with zeroed pad planes
        x = torch.randn(N, couts[0], D, H, W, device=DEV, requires_grad=True)
        w = bf(torch.randn(couts[0], cins[0], 3, 3, 3, device=DEV) * 0.1)
        y = F.pad(F.conv_transpose3d(x, w, None, stride=2, padding=1), (0, 1, 0, 1, 0, 1))
        dy = bf(torch.randn_like(y))
        y.backward(dy)
        dyz = dy.clone()
        dyz[:, :, -1] = 0; dyz[:, :, :, -1] = 0; dyz[..., -1] = 0
        ref = [x.grad]
        grid = (N, D, H, W)
        inputs = [to_ndhwc(dyz)]
    pl = P.make_conv_plan(kind, ks, stride, cins, couts, grid[1])
    dp = ops.DeviceConvPlan(pl, DEV)
    outs = [torch.full((N, *r.shape[2:], P.pad_channels(r.shape[1])), float("nan"), device=DEV, dtype=torch.bfloat16)
            for r in ref]
    if kind == "conv_dgrad" and stride == 2 and ks == 1:
        for o in outs:
            o.zero_()
    adds = [torch.zeros_like(o) for o in outs]
    if not (kind == "conv_dgrad" and stride == 2 and ks == 1):     # that launch only touches even voxels: no addend
        for a, r in zip(adds, ref):
            a[..., :r.shape[1]] = torch.randn(*a.shape[:-1], r.shape[1], device=DEV)
    st = torch.zeros(N, outs[0].shape[-1], 2, device=DEV, dtype=torch.float64)
    ops.conv_gemm(dp, inputs, dp.packed_weight(w), outs, grid, bias=dp.packed_bias(b), addends=adds,
                  stats=st if len(outs) == 1 else None, zero_last=zero_last)
    torch.cuda.synchronize()
    ops.check_device_errors()
    for o, r, a in zip(outs, ref, adds):
        want = r + from_ndhwc(a, r.shape[1])
        if zero_last:
            want[:, :, -1] = 0; want[:, :, :, -1] = 0; want[..., -1] = 0
        assert rel(from_ndhwc(o, r.shape[1]), want) < TOL_BF16
        if o.shape[-1] > r.shape[1]:
            assert o[..., r.shape[1]:].float().abs().max().item() == 0.0       # padded channels stay exactly zero
    if len(outs) == 1:
        v = outs[0].float()
        n_el = v[0, ..., 0].numel()
        assert torch.allclose(st[..., 0] / n_el, v.mean(dim=(1, 2, 3)).double(), atol=2e-2, rtol=1e-2)
        assert torch.allclose(st[..., 1] / n_el, (v * v).mean(dim=(1, 2, 3)).double(), atol=2e-2, rtol=1e-2)


WGRAD_CASES = [
    ("conv", 3, 1, [30], 30, (2, 5, 20, 12)), ("conv", 3, 1, [30, 30], 30, (1, 4, 16, 16)),
    ("conv", 1, 1, [60, 60], 60, (1, 4, 16, 16)), ("conv", 3, 1, [60], 60, (1, 6, 16, 16)),
    ("conv", 3, 1, [120], 120, (1, 4, 16, 16)), ("conv", 3, 1, [240], 240, (2, 2, 16, 16)),
    ("conv", 3, 1, [480], 480, (2, 2, 8, 8)), ("conv", 3, 1, [240, 240], 240, (1, 2, 16, 16)),
    ("conv", 3, 2, [30], 60, (1, 4, 16, 16)), ("conv", 1, 2, [30], 60, (1, 4, 16, 16)),
    ("conv", 3, 1, [8], 8, (1, 3, 5, 5)),
    ("convT", 3, 2, [60], 30, (1, 4, 16, 16)), ("convT", 3, 2, [480], 240, (1, 2, 4, 4)),
]


@pytest.mark.parametrize("kind,ks,stride,cins,cout,dims", WGRAD_CASES)
def test_wgrad_gemm(kind, ks, stride, cins, cout, dims):
    torch.manual_seed(2)
    N, D, H, W = dims          # tile grid
    if kind == "conv":
        x = bf(torch.randn(N, sum(cins), D * stride, H * stride, W * stride, device=DEV))
        w = (torch.randn(cout, sum(cins), ks, ks, ks, device=DEV) * 0.1).requires_grad_(True)
        y = F.conv3d(x, w, None, stride=stride, padding=ks // 2)
        dy = bf(torch.randn_like(y))
        y.backward(dy)
        xs, off = [], 0
        for c in cins:
            xs.append(to_ndhwc(x[:, off:off + c])); off += c
        dyn = to_ndhwc(dy)
    else:
        x = bf(torch.randn(N, cins[0], D, H, W, device=DEV))
        w = (torch.randn(cins[0], cout, 3, 3, 3, device=DEV) * 0.1).requires_grad_(True)
        y = F.pad(F.conv_transpose3d(x, w, None, stride=2, padding=1), (0, 1, 0, 1, 0, 1))
        dy = bf(torch.randn_like(y))
        dy[:, :, -1] = 0; dy[:, :, :, -1] = 0; dy[..., -1] = 0
        y.backward(dy)
        xs, dyn = [to_ndhwc(x)], to_ndhwc(dy)
    pl = P.make_wgrad_plan(kind, ks, stride, cins, cout, dims, 148)
    dp = ops.DeviceWgradPlan(pl, DEV)
    dw = torch.zeros(pl.dw_numel + 1, device=DEV)
    ops.wgrad_gemm(dp, xs, dyn, dw, dims)
    torch.cuda.synchronize()
    ops.check_device_errors()
    got = dw.index_select(0, dp.gidx).view_as(w)
    assert rel(got, w.grad) < TOL_BF16


@pytest.mark.parametrize("cins,cout,dims", [([30], 30, (1, 4, 16, 16)), ([60, 60], 60, (1, 4, 16, 16)), ([240], 120, (1, 2, 16, 16))])
def test_fp16_storage_conv_and_wgrad(cins, cout, dims):
    """precision="fp16": fp16 x fp16 MMAs in the forward conv and in the weight gradient (tcgen05.mma raises an
    illegal-instruction fault for an fp16 A with a bf16 B operand -- tried in round 1 -- so gradients are fp16 too)."""
    torch.manual_seed(7)
    N, D, H, W = dims
    h = lambda t: t.to(torch.float16).float()
    x = h(torch.randn(N, sum(cins), D, H, W, device=DEV))
    w = (h(torch.randn(cout, sum(cins), 3, 3, 3, device=DEV) * 0.1)).requires_grad_(True)
    y = F.conv3d(x, w, None, padding=1)
    dy = h(torch.randn_like(y))
    y.backward(dy)
    xs, off = [], 0
    for c in cins:
        t = torch.zeros(N, D, H, W, P.pad_channels(c), device=DEV, dtype=torch.float16)
        t[..., :c] = x[:, off:off + c].permute(0, 2, 3, 4, 1).to(torch.float16)
        xs.append(t); off += c
    pl = P.make_conv_plan("conv_fwd", 3, 1, cins, [cout], D)
    dp = ops.DeviceConvPlan(pl, DEV)
    out = torch.full((N, D, H, W, P.pad_channels(cout)), float("nan"), device=DEV, dtype=torch.float16)
    ops.conv_gemm(dp, xs, dp.packed_weight(w, torch.float16), [out], dims)
    assert rel(from_ndhwc(out, cout), y.detach()) < 6e-4          # one fp16 output rounding
    wp = ops.DeviceWgradPlan(P.make_wgrad_plan("conv", 3, 1, cins, cout, dims, 148), DEV)
    dw = torch.zeros(wp.plan.dw_numel + 1, device=DEV)
    ops.wgrad_gemm(wp, xs, to_ndhwc(dy).to(torch.float16), dw, dims)
    torch.cuda.synchronize()
    ops.check_device_errors()
    assert rel(dw.index_select(0, wp.gidx).view_as(w), w.grad) < 5e-3


def test_instance_norm_fwd_bwd():
    torch.manual_seed(3)
    N, C, D, H, W = 2, 30, 6, 10, 12
    y = bf(torch.randn(N, C, D, H, W, device=DEV) * 2 + 0.5).requires_grad_(True)
    skip = bf(torch.randn(N, C, D, H, W, device=DEV)).requires_grad_(True)
    drop = torch.empty(N, C, device=DEV).bernoulli_(0.5).mul_(2.0)
    z = y * drop.view(N, C, 1, 1, 1)
    out = F.leaky_relu(F.instance_norm(z, eps=1e-5) + skip, 0.01)
    dout = bf(torch.randn_like(out))
    out.backward(dout)
    cp = P.pad_channels(C)
    yn, sn = to_ndhwc(y.detach()), to_ndhwc(skip.detach())
    stats = torch.zeros(N, cp, 2, device=DEV, dtype=torch.float64)
    yf = yn.float()
    stats[..., 0] = yf.sum(dim=(1, 2, 3)); stats[..., 1] = (yf * yf).sum(dim=(1, 2, 3))
    dropp = torch.zeros(N, cp, device=DEV); dropp[:, :C] = drop
    table = torch.empty(N, cp, 2, device=DEV)
    ops.in_finalize(stats, dropp, table, D * H * W)
    o = torch.empty_like(yn)
    ops.in_apply(yn, sn, o, table)
    assert rel(from_ndhwc(o, C), out.detach()) < 6e-3
    g = torch.empty_like(yn); dy = torch.empty_like(yn)
    sums = torch.zeros(N, cp, 2, device=DEV, dtype=torch.float64)
    ops.in_bwd_reduce(to_ndhwc(dout), None, o, yn, g, table, sums)
    ops.in_bwd_apply(g, yn, dy, table, sums)
    torch.cuda.synchronize()
    assert rel(from_ndhwc(g, C), skip.grad) < 1e-2
    assert rel(from_ndhwc(dy, C), y.grad) < 2e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C,dims", [(240, (16, 16, 16)), (480, (8, 8, 8)), (30, (5, 6, 7)), (8, (3, 3, 3))])
def test_instance_norm_bwd_one_launch(C, dims, dtype):
    """in_bwd_small (levels 3-4: both backward passes in one launch) against the reduce + apply pair on the same tensors,
    in its three uses: first norm of a block (g recomputed), second norm (residual, g written), encoder output (two
    upstream gradients).  g is bit-identical; the sums differ only by the fp32 partial-sum order, dy by an occasional
    16-bit rounding flip."""
    torch.manual_seed(11)
    N = 2
    cp = P.pad_channels(C)
    shape = (N, *dims, cp)
    V = dims[0] * dims[1] * dims[2]
    y = (torch.randn(shape, device=DEV) * 2 + 0.5).to(dtype)
    y[..., C:] = 0
    skip = torch.randn(shape, device=DEV).to(dtype)
    dout = torch.randn(shape, device=DEV).to(dtype)
    dout2 = torch.randn(shape, device=DEV).to(dtype)
    yf = y.float()
    stats = torch.zeros(N, cp, 2, device=DEV, dtype=torch.float64)
    stats[..., 0] = yf.sum(dim=(1, 2, 3)); stats[..., 1] = (yf * yf).sum(dim=(1, 2, 3))
    table = torch.empty(N, cp, 2, device=DEV)
    ops.in_finalize(stats, None, table, V)
    out = torch.empty_like(y)
    ops.in_apply(y, skip, out, table)
    for d2, o, want_g in ((None, None, False), (None, out, True), (dout2, out, True), (dout2, None, True)):
        g_ref = torch.empty_like(y) if want_g else None
        dy_ref = torch.empty_like(y)
        sums_ref = torch.zeros(N, cp, 2, device=DEV, dtype=torch.float64)
        ops.in_bwd_reduce(dout, d2, o, y, g_ref, table, sums_ref)
        if want_g:
            ops.in_bwd_apply(g_ref, y, dy_ref, table, sums_ref)
        else:
            ops.in_bwd_apply(dout, y, dy_ref, table, sums_ref, g_is_dout=True)
        g = torch.full_like(y, float("nan")) if want_g else None
        dy = torch.full_like(y, float("nan"))
        sums = torch.full((N, cp, 2), float("nan"), device=DEV, dtype=torch.float64)
        ops.in_bwd_small(dout, d2, o, y, g, dy, table, sums)
        torch.cuda.synchronize()
        ops.check_device_errors()
        if want_g:
            assert torch.equal(g, g_ref)
        scale = sums_ref.abs().amax(dim=(0, 1), keepdim=True)
        assert float(((sums - sums_ref).abs() / scale).max()) < 1e-5
        assert not torch.isnan(dy.float()).any()
        assert rel(dy.float(), dy_ref.float()) < 2e-4, (C, dims, d2 is not None, o is not None)


@pytest.mark.parametrize("cin", [1, 2, 3])
def test_stem_and_head(cin):
    torch.manual_seed(4)
    N, D, H, W, C, K = 2, 6, 10, 12, 30, 3
    x = torch.randn(N, cin, D, H, W, device=DEV)
    w = (torch.randn(C, cin, 3, 3, 3, device=DEV) * 0.3).requires_grad_(True)
    b = torch.randn(C, device=DEV).requires_grad_(True)
    y = F.conv3d(x, w, b, padding=1)
    cp = P.pad_channels(C)
    w0 = torch.zeros(cin, 27, cp, device=DEV); w0[:, :, :C] = w.detach().reshape(C, cin, 27).permute(1, 2, 0)
    b0 = torch.zeros(cp, device=DEV); b0[:C] = b.detach()
    out = torch.empty(N, D, H, W, cp, device=DEV, dtype=torch.bfloat16)
    ops.stem_fwd(x.contiguous(), w0, b0, out)
    assert rel(from_ndhwc(out, C), y.detach()) < 4e-3
    dy = bf(torch.randn_like(y))
    y.backward(dy)
    dw0 = torch.zeros(cin, 28, cp, device=DEV)
    ops.stem_wgrad(x.contiguous(), to_ndhwc(dy), dw0)
    assert rel(dw0[:, :27, :C].permute(2, 0, 1).reshape(w.shape), w.grad) < 1e-3
    assert rel(dw0[0, 27, :C], b.grad) < 1e-3
    # head
    a = bf(torch.randn(N, C, D, H, W, device=DEV)).requires_grad_(True)
    wf = (torch.randn(K, C, 1, 1, 1, device=DEV) * 0.3).requires_grad_(True)
    bfc = torch.randn(K, device=DEV).requires_grad_(True)
    lg = F.conv3d(a, wf, bfc)
    wfp = torch.zeros(K, cp, device=DEV); wfp[:, :C] = wf.detach().reshape(K, C)
    logits = torch.empty(N, K, D, H, W, device=DEV)
    ops.head_fwd(to_ndhwc(a.detach()), wfp, bfc.detach().contiguous(), logits)
    assert rel(logits, lg.detach()) < 1e-5
    dl = torch.randn_like(lg)
    lg.backward(dl)
    da = torch.empty(N, D, H, W, cp, device=DEV, dtype=torch.bfloat16)
    dwf = torch.zeros(K * cp + K, device=DEV)
    ops.head_bwd(dl.contiguous(), to_ndhwc(a.detach()), wfp, da, dwf)
    assert rel(from_ndhwc(da, C), a.grad) < 4e-3
    assert rel(dwf[:K * cp].view(K, cp)[:, :C], wf.grad.reshape(K, C)) < 1e-3
    assert rel(dwf[K * cp:], bfc.grad) < 1e-3


def test_losses_vs_oracle():
    from oracle import unet3d_oracle as O
    L = importlib.import_module("unet3d_b200.loss")
    torch.manual_seed(5)
    N, K, D, H, W = 2, 3, 6, 10, 12
    lg = torch.randn(N, K, D, H, W) * 2
    tg = torch.randint(0, K, (N, D, H, W))
    cases = {
        "dice": (L.DiceLoss(), lambda a: O.dice_loss(a, tg)),
        "dice_w": (L.DiceLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1), lambda a: O.dice_loss(a, tg, [1, 148, 191], 0.9, 0.1)),
        "focal": (L.FocalLoss(), lambda a: O.focal_loss(a, tg)),
        "ce": (L.FocalLoss(gamma=0), lambda a: O.focal_loss(a, tg, gamma=0)),
        "hybrid_w": (L.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1),
                     lambda a: O.hybrid_loss(a, tg, weight_v=[1, 148, 191], alpha=0.9, beta=0.1)),
        "metric": (L.Dice(weight_v=[0, 1, 0]), lambda a: O.dice_metric(a, tg, weight_v=[0, 1, 0])),
    }
    for name, (mod, fn) in cases.items():
        a = lg.clone().requires_grad_(True)
        want = fn(a); want.backward()
        g = lg.clone().to(DEV).requires_grad_(True)
        got = mod(g, tg.to(DEV)); (got * 1.5).backward()
        assert abs(got.item() - want.item()) < 2e-5, name
        assert torch.allclose(g.grad.cpu() / 1.5, a.grad, rtol=2e-3, atol=1e-8), name

"""Host plans (weight packing, tap tables, parity views) executed on CPU vs torch.nn.functional."""
import importlib
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.emulate import from_ndhwc, pack_weights, run_conv_plan, to_ndhwc

P = importlib.import_module("unet3d_b200.plan")


def _rand(*s):
    return torch.randn(*s, generator=torch.Generator().manual_seed(sum(s) + len(s)))


@pytest.mark.parametrize("cins,cout,ks", [([30], 30, 3), ([12], 20, 3), ([30, 30], 30, 3), ([60], 30, 1), ([8, 8], 8, 1),
                                          ([240], 40, 3)])
def test_conv_fwd_s1(cins, cout, ks):
    N, D, H, W = 1, 3, 5, 9
    xs = [_rand(N, c, D, H, W) for c in cins]
    w = _rand(cout, sum(cins), ks, ks, ks) * 0.1
    b = _rand(cout)
    plan = P.make_conv_plan("conv_fwd", ks, 1, cins, [cout], D)
    outs = run_conv_plan(plan, [to_ndhwc(x, P.pad_channels(x.shape[1])) for x in xs], pack_weights(plan, w), (N, D, H, W),
                         (D, H, W), bias_vec=torch.cat([b, torch.zeros(1)])[torch.from_numpy(P.bias_index(plan))])
    ref = F.conv3d(torch.cat(xs, 1), w, b, padding=ks // 2)
    assert torch.allclose(from_ndhwc(outs[0], cout), ref, atol=1e-4, rtol=1e-4)
    assert outs[0][..., cout:].abs().max() == 0


@pytest.mark.parametrize("cin,cout,ks", [(30, 60, 3), (30, 60, 1), (16, 16, 3)])
def test_conv_fwd_s2(cin, cout, ks):
    N, D, H, W = 2, 4, 6, 8
    x = _rand(N, cin, D, H, W)
    w = _rand(cout, cin, ks, ks, ks) * 0.1
    plan = P.make_conv_plan("conv_fwd", ks, 2, [cin], [cout], D // 2)
    outs = run_conv_plan(plan, [to_ndhwc(x, P.pad_channels(cin))], pack_weights(plan, w), (N, D // 2, H // 2, W // 2),
                         (D // 2, H // 2, W // 2))
    ref = F.conv3d(x, w, None, stride=2, padding=ks // 2)
    assert torch.allclose(from_ndhwc(outs[0], cout), ref, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("cins,cout,ks", [([30], 30, 3), ([30, 30], 30, 3), ([60, 60], 60, 1), ([8], 24, 3)])
def test_conv_dgrad_s1(cins, cout, ks):
    N, D, H, W = 1, 3, 4, 8
    x = _rand(N, sum(cins), D, H, W).requires_grad_(True)
    w = _rand(cout, sum(cins), ks, ks, ks) * 0.1
    dy = _rand(N, cout, D, H, W)
    F.conv3d(x, w, None, padding=ks // 2).backward(dy)
    plan = P.make_conv_plan("conv_dgrad", ks, 1, [cout], cins, D)
    outs = run_conv_plan(plan, [to_ndhwc(dy, P.pad_channels(cout))], pack_weights(plan, w), (N, D, H, W), (D, H, W))
    off = 0
    for o, c in zip(outs, cins):
        assert torch.allclose(from_ndhwc(o, c), x.grad[:, off:off + c], atol=1e-4, rtol=1e-4)
        off += c


@pytest.mark.parametrize("cin,cout,ks", [(30, 60, 3), (30, 60, 1)])
def test_conv_dgrad_s2(cin, cout, ks):
    N, D, H, W = 1, 4, 6, 8
    x = _rand(N, cin, D, H, W).requires_grad_(True)
    w = _rand(cout, cin, ks, ks, ks) * 0.1
    dy = _rand(N, cout, D // 2, H // 2, W // 2)
    F.conv3d(x, w, None, stride=2, padding=ks // 2).backward(dy)
    plan = P.make_conv_plan("conv_dgrad", ks, 2, [cout], [cin], D // 2)
    outs = run_conv_plan(plan, [to_ndhwc(dy, P.pad_channels(cout))], pack_weights(plan, w), (N, D // 2, H // 2, W // 2),
                         (D, H, W))
    assert torch.allclose(from_ndhwc(outs[0], cin), x.grad, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("cin,cout", [(60, 30), (16, 8)])
def test_convT_fwd_and_dgrad(cin, cout):
    N, D, H, W = 1, 2, 3, 4
    x = _rand(N, cin, D, H, W).requires_grad_(True)
    w = _rand(cin, cout, 3, 3, 3) * 0.1
    b = _rand(cout)
    ref = F.pad(F.conv_transpose3d(x, w, b, stride=2, padding=1), (0, 1, 0, 1, 0, 1))
    plan = P.make_conv_plan("convT_fwd", 3, 2, [cin], [cout], D)
    outs = run_conv_plan(plan, [to_ndhwc(x.detach(), P.pad_channels(cin))], pack_weights(plan, w), (N, D, H, W),
                         (2 * D, 2 * H, 2 * W),
                         bias_vec=torch.cat([b, torch.zeros(1)])[torch.from_numpy(P.bias_index(plan))], zero_last=True)
    assert torch.allclose(from_ndhwc(outs[0], cout), ref, atol=1e-4, rtol=1e-4)
    dy = _rand(*ref.shape)
    ref.backward(dy)
    dyz = dy.clone()
    dyz[:, :, -1] = 0; dyz[:, :, :, -1] = 0; dyz[..., -1] = 0      # in_bwd_apply(zero_last) does this on the GPU
    plan2 = P.make_conv_plan("convT_dgrad", 3, 2, [cout], [cin], D)
    outs2 = run_conv_plan(plan2, [to_ndhwc(dyz, P.pad_channels(cout))], pack_weights(plan2, w), (N, D, H, W), (D, H, W))
    assert torch.allclose(from_ndhwc(outs2[0], cin), x.grad, atol=1e-4, rtol=1e-4)


from tests.emulate import run_wgrad_plan


@pytest.mark.parametrize("kind,ks,stride,cins,cout,D", [p if len(p) == 6 else p + (2,) for p in [
    ("conv", 3, 1, [30], 30), ("conv", 3, 1, [30, 30], 30), ("conv", 1, 1, [60, 60], 60), ("conv", 3, 1, [120], 24),
    ("conv", 3, 1, [136], 8), ("conv", 3, 2, [30], 60), ("conv", 1, 2, [30], 60), ("conv", 3, 1, [8], 8),
    ("conv", 3, 1, [40], 20), ("convT", 3, 2, [60], 30), ("convT", 3, 2, [24], 8),
    ("conv", 3, 1, [30], 30, 5), ("conv", 3, 1, [30, 30], 30, 3), ("conv", 3, 1, [12], 30, 4), ("conv", 3, 1, [30], 16, 1)]])
def test_wgrad_plans(kind, ks, stride, cins, cout, D):
    N, H, W = 1, 3, 4
    g = torch.Generator().manual_seed(5)
    if kind == "conv":
        fd = (D * stride, H * stride, W * stride)
        x = torch.randn(N, sum(cins), *fd, generator=g)
        w = (torch.randn(cout, sum(cins), ks, ks, ks, generator=g) * 0.1).requires_grad_(True)
        y = F.conv3d(x, w, None, stride=stride, padding=ks // 2)
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        xs, off = [], 0
        for c in cins:
            xs.append(to_ndhwc(x[:, off:off + c], P.pad_channels(c))); off += c
        dyn = to_ndhwc(dy, P.pad_channels(cout))
    else:
        x = torch.randn(N, cins[0], D, H, W, generator=g)
        w = (torch.randn(cins[0], cout, 3, 3, 3, generator=g) * 0.1).requires_grad_(True)
        y = F.pad(F.conv_transpose3d(x, w, None, stride=2, padding=1), (0, 1, 0, 1, 0, 1))
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        dyz = dy.clone()
        dyz[:, :, -1] = 0; dyz[:, :, :, -1] = 0; dyz[..., -1] = 0
        xs = [to_ndhwc(x, P.pad_channels(cins[0]))]
        dyn = to_ndhwc(dyz, P.pad_channels(cout))
    plan = P.make_wgrad_plan(kind, ks, stride, cins, cout, (N, D, H, W))
    dw = run_wgrad_plan(plan, xs, dyn, (N, D, H, W))
    got = dw[torch.from_numpy(plan.gidx)].reshape(w.shape).float()
    assert torch.allclose(got, w.grad, atol=1e-3, rtol=1e-3), (got - w.grad).abs().max()
    # the analytic unpack description (unet3d_dw_unpack) addresses exactly the elements gidx does
    u = plan.unpack
    tap, row, col = np.meshgrid(np.arange(u["k3"]), np.arange(u["Kp"]), np.arange(u["Np"]), indexing="ij")
    ok = (u["rowmap"][row] >= 0) & (col < u["ncols"])
    dest = u["rowmap"][row].astype(np.int64) + tap + col * u["col_stride"]
    out = np.full(w.numel(), np.nan, dw.numpy().dtype)
    o0 = u["origin"]                     # pair-mode plans keep a scratch tap slot in front of the 27 real ones
    out[dest[ok]] = dw[o0:o0 + tap.size].numpy().reshape(tap.shape)[ok]
    assert ok.sum() == w.numel() and np.array_equal(out, dw[torch.from_numpy(plan.gidx)].numpy())


@pytest.mark.parametrize("cins,cout,stride", [([30], 60, 1), ([30, 30], 30, 1), ([30], 60, 2), ([120, 120], 120, 1)])
def test_block_input_gradient_in_one_launch(cins, cout, stride):
    """dx = dgrad_conv1(dy1) + dgrad_skip_conv(g2) as one plan with two A sources (skip_k1)."""
    N, D, H, W = 1, 4, 6, 8
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, sum(cins), D, H, W, generator=g).requires_grad_(True)
    w1 = torch.randn(cout, sum(cins), 3, 3, 3, generator=g) * 0.1
    ws = torch.randn(cout, sum(cins), 1, 1, 1, generator=g) * 0.1
    y1 = F.conv3d(x, w1, None, stride=stride, padding=1)
    s = F.conv3d(x, ws, None, stride=stride)
    dy1, g2 = torch.randn(y1.shape, generator=g), torch.randn(s.shape, generator=g)
    (y1 * dy1).sum().backward(retain_graph=True)
    (s * g2).sum().backward()
    grid = (N, D // stride, H // stride, W // stride)
    plan = P.make_conv_plan("conv_dgrad", 3, stride, [cout, cout], cins, grid[1], grid, skip_k1=True)
    wp = pack_weights(plan, torch.cat([w1.reshape(-1), ws.reshape(-1)]))
    cp = P.pad_channels(cout)
    outs = run_conv_plan(plan, [to_ndhwc(dy1, cp), to_ndhwc(g2, cp)], wp, grid, (D, H, W))
    off = 0
    for o, c in zip(outs, cins):
        assert torch.allclose(from_ndhwc(o, c), x.grad[:, off:off + c], atol=1e-4, rtol=1e-4)
        off += c


@pytest.mark.parametrize("cins,cout,D", [([30], 30, 2), ([30], 30, 5), ([30, 30], 30, 3), ([12], 30, 4), ([8], 16, 6)])
def test_wgrad_pair_mode_plans(cins, cout, D, monkeypatch):
    """U3D_WG_PAIR=1 (tuning switch, off by default: measured slower, profiles/r01_notes.md): two dy planes side by side
    in N, taps kd = -1 and 3 in scratch slots around the 27 real ones."""
    monkeypatch.setenv("U3D_WG_PAIR", "1")
    N, H, W = 1, 3, 4
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, sum(cins), D, H, W, generator=g)
    w = (torch.randn(cout, sum(cins), 3, 3, 3, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(N, cout, D, H, W, generator=g)
    F.conv3d(x, w, None, padding=1).backward(dy)
    xs, off = [], 0
    for c in cins:
        xs.append(to_ndhwc(x[:, off:off + c], P.pad_channels(c))); off += c
    plan = P.make_wgrad_plan("conv", 3, 1, cins, cout, (N, D, H, W))
    assert plan.unpack["origin"] > 0 and (int(plan.tab.reshape(plan.n_jobs, -1)[0][7]) >> 30) & 1
    dw = run_wgrad_plan(plan, xs, to_ndhwc(dy, P.pad_channels(cout)), (N, D, H, W))
    got = dw[torch.from_numpy(plan.gidx)].reshape(w.shape).float()
    assert torch.allclose(got, w.grad, atol=1e-3, rtol=1e-3)

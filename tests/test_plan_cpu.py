"""Host plans (weight packing, tap tables, parity views) executed on CPU vs torch.nn.functional."""
import importlib
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

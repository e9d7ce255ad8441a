"""Model variants next to ResUnet3D on the GPU against golden vectors made from the live reference
(tests/golden/make_golden_variants.py):

* nn.MaxPool3d(2, 2) kernel: values, argmax indices and the routed gradient BIT-EXACT (ties -> first element in
  PyTorch's scan order, NaN wins), both storage formats;
* plain ``Unet`` (ConvBlockStack + MaxPoolBlock, network.py:470-487 defaults) and ``ResAttrUnet3D`` (attention gates,
  network.py:72-101,353-371): logits rel-L2 <= 1e-2 (fp16 storage) / 3e-2 (bf16 storage), loss within 5e-3, per-tensor
  gradients rel-L2 <= 0.15 (fp16) / 0.5 (bf16) (same buckets as test_model_gpu.py for 4-8 conv blocks: 16-bit rounding flips LeakyReLU / max-pool
  selections of near-tied units), InstanceNorm-cancelled conv biases ~0.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import unet3d_b200  # noqa: E402
from unet3d_b200 import ops  # noqa: E402
from unet3d_b200 import plan as P  # noqa: E402

DEV = "cuda"


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def _to_ndhwc(x, dtype):
    n, c, d, h, w = x.shape
    out = torch.zeros(n, d, h, w, P.pad_channels(c), dtype=dtype, device=DEV)
    out[..., :c] = x.permute(0, 2, 3, 4, 1).to(DEV).to(dtype)
    return out


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_maxpool_bit_exact(golden_dir, dtype):
    z = np.load(os.path.join(golden_dir, "maxpool.npz"))
    x = torch.from_numpy(z["x"])
    if dtype == torch.float16:          # the fixture is bf16-representable; fp16 needs its own reference
        x = x.half().float()
        ref_out, ref_idx = torch.nn.functional.max_pool3d(x, 2, 2, return_indices=True)
        xr = x.clone().requires_grad_(True)
        gout = torch.from_numpy(z["gout"]).half().float()
        torch.nn.functional.max_pool3d(xr, 2, 2).backward(gout)
        ref_gin = xr.grad
    else:
        ref_out, ref_idx, gout, ref_gin = (torch.from_numpy(z[k]) for k in ("out", "idx", "gout", "gin"))
    c = x.shape[1]
    xd = _to_ndhwc(x, dtype)
    n, d, h, w, cp = xd.shape
    out = torch.empty(n, d // 2, h // 2, w // 2, cp, dtype=dtype, device=DEV)
    code = torch.empty(out.shape, dtype=torch.uint8, device=DEV)
    ops.maxpool_fwd(xd, out, code)
    got = out[..., :c].permute(0, 4, 1, 2, 3).float().cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(ref_out))
    assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(ref_out))
    assert torch.equal(ops.maxpool_flat_index(code, c).cpu(), ref_idx)
    assert (out[..., c:] == 0).all() and (code[..., c:] == 0).all()
    dx = torch.full_like(xd, 7.0)                       # every element must be overwritten
    ops.maxpool_bwd(_to_ndhwc(gout, dtype), code, dx)
    assert torch.equal(dx[..., :c].permute(0, 4, 1, 2, 3).float().cpu(), ref_gin)
    assert (dx[..., c:] == 0).all()


def _variant(golden_dir, fixture, build, loss_mod, precision):
    z = np.load(os.path.join(golden_dir, fixture))
    model = build()
    model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")})
    model = model.to(DEV).eval()
    model.precision = precision
    logits = model(torch.from_numpy(z["x"]).to(DEV))
    loss = loss_mod(logits, torch.from_numpy(z["y"]).to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    ops.check_device_errors()
    ref_logits = torch.from_numpy(z["logits"])
    tol, gtol = (1e-2, 0.15) if precision == "fp16" else (3e-2, 0.5)
    r = rel(logits.detach().cpu(), ref_logits)
    print(f"[{fixture} {precision}] logits rel-L2 {r:.3e}  loss {loss.item():.6f} vs {float(z['loss']):.6f}")
    assert r < tol
    assert abs(loss.item() - float(z["loss"])) < 5e-3 * max(1.0, abs(float(z["loss"])))
    unused = set(z["unused"].tolist())
    worst = 0.0
    for name, p in model.named_parameters():
        if name in unused:
            assert p.grad is None, name
            continue
        assert p.grad is not None, name
        ref = torch.from_numpy(z["grad/" + name])
        got = p.grad.detach().cpu()
        if name.endswith("bias") and any(s in name for s in ("conv1.", "conv2.", "conv_blocks")):
            assert got.abs().max().item() < 1e-5 + 10 * ref.abs().max().item(), name      # cancelled by the norm (S1)
            continue
        rr = rel(got, ref)
        worst = max(worst, rr)
        print(f"   grad rel-L2 {rr:.3e}  {name}")
        assert rr < gtol, (name, rr)
    return worst


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_plain_unet_maxpool_vs_reference(golden_dir, precision):
    _variant(golden_dir, "plain_unet_train.npz",
             lambda: unet3d_b200.Unet(1, 3, unet3d_b200.generate_paired_features2(2, 4)), unet3d_b200.DiceLoss(), precision)


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_attention_resunet_vs_reference(golden_dir, precision):
    _variant(golden_dir, "attr_resunet.npz",
             lambda: unet3d_b200.ResAttrUnet3D(num_pool=2, num_features=8, out_channels=3),
             unet3d_b200.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1), precision)


def test_attention_default_net_runs_train_mode():
    """Default-width ResAttrUnet3D (the cascade's coarse model, nb_post_iia.py:20) trains one step at 2 x 32^3."""
    torch.manual_seed(0)
    m = unet3d_b200.ResAttrUnet3D(out_channels=3).to(DEV).train()
    x = torch.randn(2, 1, 32, 32, 32, device=DEV)
    y = torch.randint(0, 3, (2, 32, 32, 32), device=DEV)
    loss = unet3d_b200.DiceLoss()(m(x), y)
    loss.backward()
    torch.cuda.synchronize()
    ops.check_device_errors()
    assert torch.isfinite(loss)
    for name, p in m.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), name
    assert m.net.up_blocks[0].att_gate.conv.weight.grad.abs().sum() > 0


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_batchnorm_variant_vs_reference(golden_dir, precision):
    """BatchNorm3d + attention net (network.py:38-69 wiring, dropout hooks off) against the live reference's vectors:
    training mode -- logits, loss, gradients (incl. gamma, beta and the conv biases BatchNorm does not cancel in eval
    mode), running-buffer updates -- then eval mode on the updated buffers."""
    from tests.test_host_cpu import _bn_net
    z = np.load(os.path.join(golden_dir, "bn_attr_resunet.npz"))
    model = _bn_net()
    model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")})
    model = model.to(DEV)
    model.precision = precision
    x, y = torch.from_numpy(z["x"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    loss_mod = unet3d_b200.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1)
    # gradient buckets: this 4-channel fixture is noisier than the 8-channel ones -- the InstanceNorm version of the
    # very same net shows the same errors (tests/tools/bn_vs_in_check.py on a B200: median / max per-tensor rel-L2
    # 0.093 / 0.147 fp16 and 0.30 / 0.49 bf16 with InstanceNorm, 0.101 / 0.223 and 0.28 / 0.53 with BatchNorm)
    tol, gtol = (1e-2, 0.3) if precision == "fp16" else (3e-2, 0.7)
    for mode in ("train", "eval"):
        model.train(mode == "train")
        model.zero_grad(set_to_none=True)
        logits = model(x)
        loss = loss_mod(logits, y)
        loss.backward()
        torch.cuda.synchronize()
        ops.check_device_errors()
        r = rel(logits.detach().cpu(), torch.from_numpy(z[mode + "_logits"]))
        print(f"[BN {precision} {mode}] logits rel-L2 {r:.3e} loss {loss.item():.6f} vs {float(z[mode + '_loss']):.6f}")
        assert r < tol
        assert abs(loss.item() - float(z[mode + "_loss"])) < 5e-3
        for name, p in model.named_parameters():
            key = f"{mode}_grad/{name}"
            if key not in z.files:
                assert p.grad is None, name
                continue
            ref = torch.from_numpy(z[key])
            got = p.grad.detach().cpu()
            if mode == "train" and name.endswith("bias") and ("conv1." in name or "conv2." in name or "up.0." in name):
                # cancelled by the batch statistics: sum(dy) = 0 up to the 16-bit rounding of dy (8k-64k summands)
                wname = name[:-4] + "weight"
                wmax = dict(model.named_parameters())[wname].grad.abs().max().item()
                assert got.abs().max().item() < 1e-4 + 10 * ref.abs().max().item() + 0.05 * wmax, name
                continue
            rr = rel(got, ref)
            print(f"   {mode} grad rel-L2 {rr:.3e}  {name}")
            assert rr < gtol, (mode, name, rr)
        if mode == "train":
            for k in z.files:
                if k.startswith("after/"):
                    got = model.state_dict()[k[6:]].float().cpu()
                    ref = torch.from_numpy(z[k]).float()
                    assert torch.allclose(got, ref, rtol=2e-2, atol=2e-3), k


def test_attr2_net_vs_reference(golden_dir):
    """ResAttrUnet3D2 (network.py:6-35): five poolings, widths 30/60/120/240/320/320 (a 320 -> 320 stride-2 pooling block
    with its skip conv, a 640 -> 320 decoder), attention gates on every level.  Weights by seed (the parameter order is
    the reference's), one 1 x 64^3 patch, fp16 storage.

    This randomly initialised net is ill-conditioned at 64^3 (InstanceNorm over the 8 voxels of its 2^3 bottom grid, six
    levels deep, gates on top): the oracle's own 16-bit storage model moves the logits by 4.4e-2 in fp16 and 2.5e-1 in
    bf16 (tests/test_oracle_golden.py pins the oracle to the live reference's fixture), and two 16-bit realisations of
    it differ from each other by as much (3.5e-2 measured between the CUDA path and that storage model).  So here the
    logits are held to the live reference at the bar the storage model predicts (6e-2), the loss to 5e-3 and the
    decoder-level gradients to the fixture; the tight statement for this net is block-wise:
    tests/test_block_parity_gpu.py::test_attr2_net_every_block (every forward stage 4e-3, every gradient 1e-2)."""
    from oracle import unet3d_oracle as O
    from oracle import bf16_model as Q
    z = np.load(os.path.join(golden_dir, "attr2_64.npz"))
    torch.manual_seed(int(z["weight_seed"]))
    model = unet3d_b200.ResAttrUnet3D2(in_channels=1, out_channels=3)
    assert np.allclose([float(p.detach().double().sum()) for _, p in model.named_parameters()], z["weight_sum"], rtol=1e-9)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).eval()
    model.precision = "fp16"
    from tests.golden.make_golden import blocky_labels
    xc = torch.randn(1, 1, 64, 64, 64, generator=torch.Generator().manual_seed(int(z["x_seed"])))
    x = xc.to(DEV)
    y = torch.from_numpy(blocky_labels((1, 64, 64, 64), int(z["label_seed"]))).to(DEV)
    logits = model(x)
    loss = unet3d_b200.DiceLoss()(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    ops.check_device_errors()
    got_full = logits.detach().cpu()
    with torch.no_grad():
        emu = Q.resunet3d_forward(sd, xc, dtype=torch.float16, attention=True, pf=O.ATTR2_FEATURES)
    r_emu = rel(got_full, emu)
    ref = torch.from_numpy(z["logits_sub"])
    got = got_full[:, :, ::2, ::2, ::2]
    r = rel(got, ref)
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"[ResAttrUnet3D2 fp16] logits rel-L2 {r_emu:.3e} vs the 16-bit-storage oracle, {r:.3e} vs the live reference "
          f"(storage model: {rel(emu[:, :, ::2, ::2, ::2], ref):.3e}), argmax agreement {agree:.5f}, "
          f"loss {loss.item():.6f} vs {float(z['loss']):.6f}")
    assert r_emu < 6e-2 and r < 6e-2 and agree >= 0.98
    assert abs(loss.item() - float(z["loss"])) < 5e-3
    unused = set(z["unused"].tolist())
    names = z["names"].tolist()
    params = dict(model.named_parameters())
    for k in names:
        assert (params[k].grad is None) == (k in unused), k
    for key, name in (("grad_fc_w", "net.fc.weight"), ("grad_att0_w", "net.up_blocks.0.att_gate.conv.weight"),
                      ("grad_att0_b", "net.up_blocks.0.att_gate.conv.bias")):
        rr = rel(params[name].grad.detach().cpu(), torch.from_numpy(z[key]))
        print(f"   grad rel-L2 {rr:.3e}  {name}")
        assert rr < 0.15, (name, rr)         # decoder-level-0 tensors behind the ill-conditioned forward (see above)

"""Teacher-forced block-level gradient parity.

A whole-net gradient comparison against the fp32 oracle cannot be tight on a 16-bit path: the rounding of the forward
activations flips the LeakyReLU mask of near-zero units and forty layers of that decorrelate the encoder gradients
(tests/test_model_gpu.py measured 0.3-0.9 rel-L2 on the default net).  That says nothing about whether the BACKWARD
WIRING is right, so here every block of the backward pass is pinned on its own:

  * the CUDA run records, per block, the tensors it consumed (block input(s), upstream gradient) and produced
    (input gradient(s), parameter gradients)  -- ``UNetEngine.capture``;
  * the oracle block (``oracle/bf16_model.py``: the reference's ResBlock / ConvTrans3D / AttBlock, network.py:298-416,
    with the CUDA path's 16-bit storage points) is run on EXACTLY those inputs, STAGE BY STAGE: every stored
    intermediate of the CUDA block (conv1 output, first activation, conv2 output, block output) must equal the oracle
    stage computed from the CUDA run's previous stage to rel-L2 <= 4e-3 (forward parity, one 16-bit rounding), and the
    value that flows on is the CUDA run's own tensor (``res_block_forced``);
  * autograd through that forced block gives the oracle's gradients for exactly the CUDA run's upstream gradient: every
    parameter gradient and every block-input gradient must agree to rel-L2 <= 1e-2 (bf16 and fp16 alike).
    Measured on a B200 (gpurun_out/r2_diag1.log): WITHOUT forcing the intermediates two 16-bit realisations of one block
    differ by rounding flips that flip ~0.2 % of the LeakyReLU masks = 1.5-3 % in every gradient of the block; that
    un-forced comparison and the one against the plain fp32 block are printed as diagnostics, not asserted.

The head (Conv3d k1, network.py:545-547) and the stem (network.py:541) are pinned the same way.  A whole-net check
stays as a sanity bound: cosine similarity of the full parameter gradient against the fp32 oracle.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import unet3d_b200  # noqa: E402
from unet3d_b200 import ops  # noqa: E402
from oracle import unet3d_oracle as O  # noqa: E402
from oracle import bf16_model as Q  # noqa: E402

DEV = "cuda"
TOL = 1e-2                       # gradients vs the stage-forced oracle block
FWD_TOL = 4e-3                   # every stored forward stage vs the oracle stage computed from the previous CUDA stage
                                 # (measured: <= 1.9e-3 bf16 -- the block output sits behind two roundings, skip conv + sum)
NORM_TOL = 5e-3                  # normalised values from the CUDA table (statistics of the fp32 accumulators) vs
                                 # instance_norm of the stored 16-bit tensor (measured 1.2e-3 on an 8-voxel grid, bf16)


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def ncdhw(t, c):
    """(N, D, H, W, Cp) 16-bit device tensor -> (N, c, D, H, W) fp32 CPU."""
    return t[..., :c].permute(0, 4, 1, 2, 3).float().cpu().contiguous()


def _captured_step(model, x, y, loss_fn, train):
    model.train(train)
    loss_fn(model(x), y).backward()          # builds the engine; the first backward of a shape runs layer by layer
    model.zero_grad(set_to_none=True)
    eng = model.net._engine
    eng.capture = []
    if train:
        torch.manual_seed(77)
    logits = model(x)
    loss = loss_fn(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    ops.check_device_errors()
    cap, scale = eng.capture, eng.capture_scale
    eng.capture = None
    return cap, (1.0 if scale is None else float(scale.item())), logits.detach()


def _check_blocks(model, cap, gscale, precision):
    dt = torch.float16 if precision == "fp16" else torch.bfloat16
    names = {m: n for n, m in model.named_modules()}
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    grads = {k: (None if p.grad is None else p.grad.detach().cpu()) for k, p in model.named_parameters()}
    worst = {"grad": (0.0, None), "fwd": (0.0, None), "free": (0.0, None), "fp32": (0.0, None)}
    n_checked = 0

    def note(slot, r, what):
        if r > worst[slot][0]:
            worst[slot] = (r, what)

    def compare(tag, what, got, want, free=None, f32=None):
        nonlocal n_checked
        r = rel(got, want)
        n_checked += 1
        note("grad", r, f"{tag} {what}")
        if free is not None:
            note("free", rel(got, free), f"{tag} {what}")
        if f32 is not None:
            note("fp32", rel(got, f32), f"{tag} {what}")
        assert r <= TOL, f"{tag} {what}: rel-L2 {r:.3e} vs the stage-forced oracle block"

    def run(fn, params, ins, dout, modes=("forced", "free", "fp32")):
        """autograd through an oracle block: stage-forced (asserted), un-forced 16-bit storage and plain fp32
        (diagnostics).  Returns per mode ([input grads], {param: grad}, stages)."""
        outs = []
        for mode in modes:
            p = {k: sd[k].clone().requires_grad_(True) for k in params}
            xs = [t.clone().requires_grad_(True) for t in ins]
            out, st = fn({**sd, **p}, xs, None if mode == "fp32" else dt, mode == "forced")
            out.backward(dout)
            outs.append(([t.grad for t in xs], {k: v.grad for k, v in p.items()}, st))
        return outs

    def fwd_check(tag, st, gpu):
        for k, v in gpu.items():
            if k.startswith("t"):                      # an InstanceNorm table: compare the normalised tensors
                src = {"t1": "y1", "t": "y"}[k]
                r = rel(Q._tab_norm(gpu[src], v), st["n" + k[1:]].detach())
                assert r <= NORM_TOL, f"{tag} normalised values from table {k}: rel-L2 {r:.3e}"
                continue
            r = rel(v, st[k].detach())
            note("fwd", r, f"{tag} {k}")
            assert r <= FWD_TOL, f"{tag} forward stage {k}: rel-L2 {r:.3e}"

    for e in cap:
        kind, tag = e["kind"], "/".join(str(v) for v in e["key"])
        if kind == "res":
            blk = e["blk"]
            pre = names[blk] + "."
            cin, cout, stride = blk.in_channels, blk.out_channels, blk.stride
            x_in = torch.cat([ncdhw(t, c) for t, c in zip(e["inputs"], e["in_C"])], 1)
            dout = ncdhw(e["dout"], cout)
            if e["dout2"] is not None:
                dout = dout + ncdhw(e["dout2"], cout)
            dout = dout / gscale
            mask = None if e["drop"] is None else e["drop"][:, :cout].reshape(-1, cout, 1, 1, 1).float().cpu()
            gpu = {k: ncdhw(v, cout) for k, v in e["fwd"].items()}
            gpu["t1"] = e["tab"]["t1"][:, :cout].float().cpu()
            params = [pre + "conv1.weight", pre + "conv2.weight"]
            if blk.uses_skip_conv:
                params += [pre + "skip_conv.weight", pre + "skip_conv.bias"]
            (gi, gp, st), (gi_u, gp_u, _), (gi_f, gp_f, _) = run(
                lambda s, xs, d, forced: Q.res_block_forced(s, pre, xs[0], cin, cout, stride, d, mask, gpu if forced else {}),
                params, [x_in], dout)
            fwd_check(tag, st, gpu)
            off = 0
            for t, c, i in zip(e["dins"], e["in_C"], range(9)):
                sl = slice(off, off + c)
                compare(tag, f"d(input {i})", ncdhw(t, c) / gscale, gi[0][:, sl], gi_u[0][:, sl], gi_f[0][:, sl])
                off += c
            for k in params:
                compare(tag, k[len(pre):], grads[k], gp[k], gp_u[k], gp_f[k])
            for k in (pre + "conv1.bias", pre + "conv2.bias"):          # cancelled by the norm
                assert grads[k].abs().max().item() == 0.0
        elif kind == "up":
            ct = e["ct"]
            pre = names[ct][:-len("up.0")]
            x_in = ncdhw(e["xin"], ct.in_channels)
            dout = ncdhw(e["dout"], ct.out_channels) / gscale
            gpu = {k: ncdhw(v, ct.out_channels) for k, v in e["fwd"].items()}
            gpu["t"] = e["tab"]["t"][:, :ct.out_channels].float().cpu()
            params = [pre + "up.0.weight", pre + "up.0.bias"]
            (gi, gp, st), (gi_u, gp_u, _), (gi_f, gp_f, _) = run(
                lambda s, xs, d, forced: Q.conv_trans_forced(s, pre, xs[0], d, gpu if forced else {}), params, [x_in], dout)
            fwd_check(tag, st, gpu)
            compare(tag, "d(input)", ncdhw(e["din"], ct.in_channels) / gscale, gi[0], gi_u[0], gi_f[0])
            for k in params:
                compare(tag, k[len(pre):], grads[k], gp[k], gp_u[k], gp_f[k])
        elif kind == "att":
            gate = e["gate"]
            pre = names[gate] + "."
            c = gate.conv.in_channels
            skip, au = ncdhw(e["skip"], c), ncdhw(e["au"], c)
            dout = ncdhw(e["dout"], c) / gscale
            gpu = {k: ncdhw(v, c) for k, v in e["fwd"].items()}
            params = [pre + "conv.weight", pre + "conv.bias"]
            (gi, gp, st), (gi_u, gp_u, _), (gi_f, gp_f, _) = run(
                lambda s, xs, d, forced: Q.att_block_forced(s, pre, xs[0], xs[1], d, gpu if forced else {}),
                params, [skip, au], dout)
            fwd_check(tag, st, gpu)
            d_up_in = ncdhw(e["d_up_in"], c) / gscale
            compare(tag, "d(skip)", ncdhw(e["dskip"], c) / gscale, gi[0], gi_u[0], gi_f[0])
            compare(tag, "d(gate)", ncdhw(e["dup"], c) / gscale, gi[1] + d_up_in, gi_u[1] + d_up_in, gi_f[1] + d_up_in)
            for k in params:
                compare(tag, k[len(pre):], grads[k], gp[k], gp_u[k], gp_f[k])
        elif kind == "head":
            fc = model.net.fc
            a = ncdhw(e["a"], fc.in_channels)
            dl = e["dlogits"].float().cpu()
            params = ["net.fc.weight", "net.fc.bias"]
            ((gi, gp, _),) = run(lambda s, xs, d, forced: (F.conv3d(xs[0], s["net.fc.weight"], s["net.fc.bias"]), {}),
                                 params, [a], dl, modes=("forced",))
            compare(tag, "d(input)", ncdhw(e["din"], fc.in_channels) / gscale, gi[0])
            for k in params:
                compare(tag, k, grads[k], gp[k])
        elif kind == "stem":
            conv = model.net.conv
            x = e["x"].float().cpu()
            dout = ncdhw(e["dout"], conv.out_channels) / gscale
            params = ["net.conv.weight", "net.conv.bias"]
            ((_, gp, _),) = run(lambda s, xs, d, forced: (F.conv3d(xs[0], s["net.conv.weight"], s["net.conv.bias"], padding=1), {}),
                                params, [x], dout, modes=("forced",))
            for k in params:
                compare(tag, k, grads[k], gp[k])
        else:
            raise AssertionError(kind)
    print(f"[{precision}] {len(cap)} blocks, {n_checked} gradients: worst vs stage-forced oracle block "
          f"{worst['grad'][0]:.3e} ({worst['grad'][1]}); worst forward stage {worst['fwd'][0]:.3e} ({worst['fwd'][1]}); "
          f"diagnostics -- un-forced 16-bit oracle block {worst['free'][0]:.3e} ({worst['free'][1]}), plain fp32 block "
          f"{worst['fp32'][0]:.3e} ({worst['fp32'][1]})")
    return n_checked


def _model_and_batch(ctor, shape, precision, seed=0):
    torch.manual_seed(seed)
    model = ctor().to(DEV)
    model.precision = precision
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1234)).to(DEV)
    y = torch.randint(0, 3, (shape[0], *shape[2:]), generator=torch.Generator().manual_seed(4321)).to(DEV)
    return model, x, y


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_default_net_every_block(precision):
    """ResUnet3D(4, 30, out=3) at 1 x 32^3: 19 residual blocks (stride 1 / 2, concat inputs, skip convs), 4 transposed
    convs, head and stem -- each pinned to its oracle block on the CUDA run's own tensors."""
    model, x, y = _model_and_batch(lambda: unet3d_b200.ResUnet3D(4, 30, 1, 3), (1, 1, 32, 32, 32), precision)
    cap, gscale, _ = _captured_step(model, x, y, unet3d_b200.DiceLoss(), train=False)
    assert sum(1 for e in cap if e["kind"] == "res") == 19 and sum(1 for e in cap if e["kind"] == "up") == 4
    assert _check_blocks(model, cap, gscale, precision) >= 19 * 3 + 4 * 3 + 5


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_train_mode_masks_and_batch(precision):
    """Dropout3d masks (train mode), batch 2, hybrid loss, non-cubic patch: masks folded into the norm scale."""
    model, x, y = _model_and_batch(lambda: unet3d_b200.ResUnet3D(2, 8, 1, 3), (2, 1, 24, 16, 32), precision, seed=3)
    loss = unet3d_b200.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1)
    cap, gscale, _ = _captured_step(model, x, y, loss, train=True)
    assert any(e["kind"] == "res" and e["drop"] is not None for e in cap)
    _check_blocks(model, cap, gscale, precision)


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_attention_net_every_block(precision):
    """ResAttrUnet3D (network.py:72-101): the gate's shared 1x1x1 conv receives three gradient contributions."""
    model, x, y = _model_and_batch(lambda: unet3d_b200.ResAttrUnet3D(2, 16, 1, 3), (1, 1, 16, 16, 16), precision, seed=5)
    cap, gscale, _ = _captured_step(model, x, y, unet3d_b200.DiceLoss(), train=False)
    assert sum(1 for e in cap if e["kind"] == "att") == 2
    _check_blocks(model, cap, gscale, precision)


def test_attr2_net_every_block():
    """ResAttrUnet3D2 (network.py:6-35; 30/60/120/240/320/320, five poolings, gates on every level) at 1 x 64^3, fp16
    storage: 26 residual blocks, 5 transposed convs, 5 attention gates.  The whole-net logits of this randomly
    initialised net are ill-conditioned (tests/test_variants_gpu.py::test_attr2_net_vs_reference); block by block the
    CUDA path is held to the same bars as every other net."""
    model, x, y = _model_and_batch(lambda: unet3d_b200.ResAttrUnet3D2(1, 3), (1, 1, 64, 64, 64), "fp16")
    cap, gscale, _ = _captured_step(model, x, y, unet3d_b200.DiceLoss(), train=False)
    assert sum(1 for e in cap if e["kind"] == "att") == 5 and sum(1 for e in cap if e["kind"] == "up") == 5
    _check_blocks(model, cap, gscale, "fp16")


def test_whole_net_gradient_cosine_fp16():
    """Sanity bound on the whole backward pass (default net, fp16 storage, 1 x 32^3) against the fp32 oracle: the
    cosine of the full parameter gradient (all tensors concatenated) and of every weight tensor on its own."""
    model, x, y = _model_and_batch(lambda: unet3d_b200.ResUnet3D(4, 30, 1, 3), (1, 1, 32, 32, 32), "fp16")
    model.eval()
    unet3d_b200.DiceLoss()(model(x), y).backward()
    torch.cuda.synchronize()
    sdr = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    O.dice_loss(O.resunet3d_forward(sdr, x.cpu(), 4, 30), y.cpu()).backward()
    dot = na = nb = 0.0
    low = []
    for k, p in model.named_parameters():
        if p.grad is None:
            assert sdr[k].grad is None
            continue
        if k.endswith("bias") and ("conv1" in k or "conv2" in k):
            continue                                  # cancelled biases: float noise in the reference (S1)
        a, b = p.grad.detach().cpu().double().reshape(-1), sdr[k].grad.double().reshape(-1)
        dot, na, nb = dot + float(a @ b), na + float(a @ a), nb + float(b @ b)
        c = float(a @ b) / max(float(a.norm() * b.norm()), 1e-300)
        if k.endswith("weight"):
            low.append((c, k))
    cos = dot / (na * nb) ** 0.5
    low.sort()
    print(f"whole-net gradient cosine {cos:.5f}; lowest per-tensor: {[(round(c, 4), k) for c, k in low[:4]]}")
    # measured on a B200: 0.979 whole-net, 0.954 for the worst tensor (pool_blocks.3.conv1.weight): the LeakyReLU mask
    # flips of ~40 stacked 16-bit layers; the wiring itself is pinned block by block above
    assert cos >= 0.97, cos
    assert low[0][0] >= 0.9, low[0]

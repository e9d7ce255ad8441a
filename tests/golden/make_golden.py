"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

/root/reference does not exist on the GPU box, so everything the GPU-side tests need is
committed here as small .npz files.  The reference has no tests/fixtures of its own
(SURVEY.md §4), so these vectors -- outputs of the live reference modules on seeded
synthetic inputs -- are what pins the oracle (oracle/unet3d_oracle.py).

Stubs/shims follow SURVEY.md §8c: apex/torchsummary/nibabel/transforms3d are stubbed,
``np.int`` is restored, nothing in /root/reference is modified.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("apex", "torchsummary", "nibabel", "transforms3d", "transforms3d.affines",
                 "skimage", "skimage.measure", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["apex"].amp = types.SimpleNamespace()
    sys.modules["torchsummary"].summary = lambda *a, **k: None
    sys.modules["transforms3d.affines"].compose = None
    sys.modules["transforms3d.affines"].decompose = None
    sys.modules["transforms3d"].affines = sys.modules["transforms3d.affines"]
    if not hasattr(np, "int"):
        np.int = int
    import network, loss  # noqa: E401
    try:
        import trainer
    except Exception as e:  # pragma: no cover - diagnostic only
        print("trainer import failed:", repr(e))
        trainer = None
    return network, loss, trainer


def grads_of(model):
    return {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in model.named_parameters()}


def blocky_labels(shape, seed):
    """bg / kidney / tumour nested ellipsoids with a seeded centre jitter."""
    g = np.random.RandomState(seed)
    n, d, h, w = shape
    zz, yy, xx = np.meshgrid(np.linspace(-1, 1, d), np.linspace(-1, 1, h), np.linspace(-1, 1, w), indexing="ij")
    out = np.zeros(shape, dtype=np.int64)
    for i in range(n):
        c = g.uniform(-0.2, 0.2, size=3)
        r = (zz - c[0]) ** 2 + (yy - c[1]) ** 2 / 0.8 + (xx - c[2]) ** 2 / 0.6
        out[i][r < 0.5] = 1
        out[i][r < 0.12] = 2
    return out


def main():
    network, loss_mod, trainer = import_reference()
    torch.set_num_threads(os.cpu_count())

    # ---------------------------------------------------------------- 1. small net, weights stored
    torch.manual_seed(7)
    small = network.ResUnet3D(num_pool=2, num_features=8, in_channels=1, out_channels=3)
    sd = {k: v.detach().clone() for k, v in small.state_dict().items()}
    x = torch.randn(2, 1, 16, 16, 16, generator=torch.Generator().manual_seed(11))
    y = torch.from_numpy(blocky_labels((2, 16, 16, 16), 5))
    small.eval()
    logits = small(x)
    crit = loss_mod.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1)
    l = crit(logits, y)
    l.backward()
    g = grads_of(small)
    out = {"x": x.numpy(), "y": y.numpy(), "logits": logits.detach().numpy(), "loss": np.float32(l.item())}
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    for k, v in g.items():
        if v is not None:
            out["grad/" + k] = v.numpy()
    out["unused"] = np.array([k for k, v in g.items() if v is None])
    np.savez_compressed(os.path.join(HERE, "small_resunet.npz"), **out)
    print("small net: loss", l.item(), "params", sum(p.numel() for p in small.parameters()))

    # train-mode (dropout) forward with recorded masks
    torch.manual_seed(123)
    small.train()
    small.zero_grad()
    logits_t = small(x)
    np.savez_compressed(os.path.join(HERE, "small_resunet_train.npz"), logits=logits_t.detach().numpy(),
                        seed=np.int64(123))

    # ---------------------------------------------------------------- 2. default net, seed-only weights
    torch.manual_seed(0)
    net = network.ResUnet3D(out_channels=3)
    net.eval()
    x = torch.randn(1, 1, 32, 32, 32, generator=torch.Generator().manual_seed(1234))
    y = torch.from_numpy(blocky_labels((1, 32, 32, 32), 9))
    logits = net(x)
    dl = loss_mod.DiceLoss()
    l = dl(logits, y)
    l.backward()
    g = grads_of(net)
    names = [k for k, _ in net.named_parameters()]
    gnorm = np.array([0.0 if g[k] is None else float(g[k].double().norm()) for k in names])
    gsum = np.array([0.0 if g[k] is None else float(g[k].double().sum()) for k in names])
    wsum = np.array([float(p.detach().double().sum()) for _, p in net.named_parameters()])
    np.savez_compressed(os.path.join(HERE, "default_resunet_32.npz"),
                        logits=logits.detach().numpy().astype(np.float32), loss=np.float32(l.item()),
                        names=np.array(names), grad_norm=gnorm, grad_sum=gsum, weight_sum=wsum,
                        grad_fc_w=g["net.fc.weight"].numpy(), grad_conv_w=g["net.conv.weight"].numpy(),
                        grad_first_res_conv1=g["net.encode_blocks.0.res_blocks.0.conv1.weight"].numpy(),
                        unused=np.array([k for k in names if g[k] is None]),
                        x_seed=np.int64(1234), label_seed=np.int64(9), weight_seed=np.int64(0),
                        state_keys=np.array(list(net.state_dict().keys())),
                        state_shapes=np.array([str(tuple(v.shape)) for v in net.state_dict().values()]))
    print("default net 32^3: loss", l.item(), "n params", len(names),
          "unused", sum(1 for k in names if g[k] is None))

    # ---------------------------------------------------------------- 3. losses on random logits
    gen = torch.Generator().manual_seed(77)
    lg = torch.randn(2, 3, 12, 10, 8, generator=gen) * 2.0
    tg = torch.randint(0, 3, (2, 12, 10, 8), generator=gen)
    vals = {}
    for name, mod in {
        "dice": loss_mod.DiceLoss(),
        "dice_w": loss_mod.DiceLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1),
        "focal": loss_mod.FocalLoss(),
        "focal_w": loss_mod.FocalLoss(gamma=2, weight_v=[1, 148, 191]),
        "ce": loss_mod.FocalLoss(gamma=0),
        "hybrid": loss_mod.HybirdLoss(),
        "hybrid_w": loss_mod.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1),
        "metric": loss_mod.Dice(weight_v=[0, 1, 0]),
    }.items():
        a = lg.clone().requires_grad_(True)
        v = mod(a, tg)
        v.backward()
        vals[name] = np.float32(v.item())
        vals[name + "_grad"] = a.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "losses.npz"), logits=lg.numpy(), target=tg.numpy(), **vals)
    print("losses:", {k: float(v) for k, v in vals.items() if not k.endswith("_grad")})

    # ---------------------------------------------------------------- 4. tile grids + predict_per_patch
    if trainer is not None:
        cases = [(512, 128, 2), (256, 128, 2), (512, 128, 4), (256, 128, 4), (128, 128, 2), (130, 128, 4),
                 (96, 96, 4), (100, 96, 4), (160, 64, 2), (144, 96, 3), (300, 144, 4), (80, 32, 1), (33, 32, 2)]
        centres = {}
        for ext, p, spp in cases:
            start, end = p // 2, ext - p // 2
            ns = np.ceil((end - start) / (p / spp))
            step = (end - start) / (ns + 1e-8)
            if step == 0:
                step = 9999999
            centres[f"{ext}_{p}_{spp}"] = np.arange(start, end + 1e-8, step, dtype=int)
        np.savez_compressed(os.path.join(HERE, "tile_centres.npz"), **centres)

        # toy model: fixed 3-class 3x3x3 conv; volume smaller than patch on one axis (pad path)
        torch.manual_seed(3)
        toy = torch.nn.Conv3d(1, 3, 3, padding=1)
        vol = np.random.RandomState(2).randn(40, 21, 36, 1).astype(np.float32)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lab = trainer.predict_per_patch(vol, toy, num_classes=3, patch_size=(16, 24, 16),
                                            step_per_patch=2, verbose=False, one_hot=False)
            prob = trainer.predict_per_patch(vol, toy, num_classes=3, patch_size=(16, 24, 16),
                                             step_per_patch=2, verbose=False, one_hot=True)
        np.savez_compressed(os.path.join(HERE, "predict_toy.npz"), vol=vol, labels=lab, probs=prob,
                            w=toy.weight.detach().numpy(), b=toy.bias.detach().numpy())
        print("predict toy:", lab.shape, prob.shape, "nan frac", float(np.isnan(prob).mean()))

    # ---------------------------------------------------------------- 5. plain Unet (max-pool) small
    torch.manual_seed(5)
    pf = network.generate_paired_features2(2, 4)
    plain = network.Unet(1, 2, pf)
    plain.eval()
    x = torch.randn(1, 1, 16, 16, 16, generator=torch.Generator().manual_seed(21))
    out = {"x": x.numpy(), "logits": plain(x).detach().numpy(), "pf": np.array(pf)}
    for k, v in plain.state_dict().items():
        out["sd/" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "plain_unet.npz"), **out)
    print("done")


if __name__ == "__main__":
    main()

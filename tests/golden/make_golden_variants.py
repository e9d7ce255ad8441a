"""Golden fixtures for the model variants next to ResUnet3D, from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_variants.py      # writes attr_resunet.npz, plain_unet_train.npz, maxpool.npz
    python tests/golden/make_golden_variants.py --bn # writes bn_attr_resunet.npz

Kept apart from make_golden.py so that the fixtures that file wrote stay byte-identical.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, grads_of, blocky_labels  # noqa: E402


def pack(model, x, y, logits, loss):
    out = {"x": x.numpy(), "y": y.numpy(), "logits": logits.detach().numpy(), "loss": np.float32(loss.item())}
    for k, v in model.state_dict().items():
        out["sd/" + k] = v.detach().numpy()
    g = grads_of(model)
    for k, v in g.items():
        if v is not None:
            out["grad/" + k] = v.numpy()
    out["unused"] = np.array([k for k, v in g.items() if v is None])
    out["param_order"] = np.array([k for k, _ in model.named_parameters()])
    return out


def main():
    network, loss_mod, _ = import_reference()

    # ---- ResAttrUnet3D (attention gates, network.py:72-101), small, eval mode, hybrid loss
    torch.manual_seed(17)
    net = network.ResAttrUnet3D(num_pool=2, num_features=8, in_channels=1, out_channels=3).eval()
    x = torch.randn(2, 1, 16, 16, 16, generator=torch.Generator().manual_seed(31))
    y = torch.from_numpy(blocky_labels((2, 16, 16, 16), 6))
    logits = net(x)
    l = loss_mod.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1)(logits, y)
    l.backward()
    np.savez_compressed(os.path.join(HERE, "attr_resunet.npz"), **pack(net, x, y, logits, l))
    print("ResAttrUnet3D small: loss", l.item())

    # ---- plain Unet (ConvBlockStack + MaxPoolBlock, network.py:470-487 defaults), eval mode, Dice loss, gradients
    torch.manual_seed(5)
    pf = network.generate_paired_features2(2, 4)
    plain = network.Unet(1, 3, pf).eval()
    x = torch.randn(2, 1, 16, 16, 16, generator=torch.Generator().manual_seed(21))
    y = torch.from_numpy(blocky_labels((2, 16, 16, 16), 8))
    logits = plain(x)
    l = loss_mod.DiceLoss()(logits, y)
    l.backward()
    out = pack(plain, x, y, logits, l)
    out["pf"] = np.array(pf)
    np.savez_compressed(os.path.join(HERE, "plain_unet_train.npz"), **out)
    print("plain Unet: loss", l.item())

    # ---- nn.MaxPool3d(2, 2) values + indices on bf16-representable inputs with ties and a NaN
    g = torch.Generator().manual_seed(3)
    xin = torch.randn(2, 5, 6, 8, 4, generator=g).bfloat16().float()
    xin[0, 0, 0:2, 0:2, 0:2] = 1.5                  # an all-tie window: the first element wins
    xin[1, 2, 2, 3, 1] = xin[1, 2, 3, 2, 0] = 7.0   # a two-way tie inside one window
    xin[0, 3, 4, 6, 2] = float("nan")               # NaN propagates and is selected
    vals, idx = torch.nn.functional.max_pool3d(xin, 2, 2, return_indices=True)
    gout = torch.randn(vals.shape, generator=g).bfloat16().float()
    xr = xin.clone().requires_grad_(True)
    torch.nn.MaxPool3d(kernel_size=2, stride=2)(xr).backward(gout)
    np.savez_compressed(os.path.join(HERE, "maxpool.npz"), x=xin.numpy(), out=vals.numpy(), idx=idx.numpy(),
                        gout=gout.numpy(), gin=xr.grad.numpy())
    print("done")


def bn_variant():
    """ResAttrBNUnet3D (BatchNorm3d + attention gates, network.py:38-69) small: (a) TRAINING mode without dropout
    (the reference's kwargs hooks: dropout_op=None) -- batch statistics, running-buffer updates, gamma/beta gradients --
    and (b) eval mode of the same net afterwards (running statistics, conv biases no longer cancelled)."""
    import torch.nn as nn
    network, loss_mod, _ = import_reference()
    torch.manual_seed(23)
    bn = {'norm_op': nn.BatchNorm3d}
    nd = {'norm_op': nn.BatchNorm3d, 'dropout_op': None}
    net = network.Unet(1, 3, network.generate_paired_features(2, 4), pool_block=network.ResBlock,
                       pool_kwargs={'stride': 2, **nd}, up_kwargs={'attention': True, **bn},
                       encode_block=network.ResBlockStack, encode_kwargs=nd,
                       encode_kwargs_fn=lambda level: {'num_stacks': max(level, 1)},
                       decode_block=network.ResBlock, decode_kwargs=nd)
    # make gamma / beta / running statistics non-trivial
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.BatchNorm3d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.6, 1.4)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.randn(2, 1, 16, 16, 16, generator=torch.Generator().manual_seed(41))
    y = torch.from_numpy(blocky_labels((2, 16, 16, 16), 4))
    crit = loss_mod.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1)
    net.train()
    logits = net(x)
    l = crit(logits, y)
    l.backward()
    out = {"x": x.numpy(), "y": y.numpy(), "train_logits": logits.detach().numpy(), "train_loss": np.float32(l.item())}
    for k, v in sd0.items():
        out["sd/" + k] = v.numpy()
    for k, v in net.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["after/" + k] = v.detach().numpy()
    g = grads_of(net)
    for k, v in g.items():
        if v is not None:
            out["train_grad/" + k] = v.numpy()
    out["unused"] = np.array([k for k, v in g.items() if v is None])
    out["param_order"] = np.array([k for k, _ in net.named_parameters()])
    net.eval()
    net.zero_grad()
    logits_e = net(x)
    le = crit(logits_e, y)
    le.backward()
    out["eval_logits"] = logits_e.detach().numpy()
    out["eval_loss"] = np.float32(le.item())
    for k, v in grads_of(net).items():
        if v is not None:
            out["eval_grad/" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "bn_attr_resunet.npz"), **out)
    print("BN variant: train loss", l.item(), "eval loss", le.item())


if __name__ == "__main__":
    if "--bn" in sys.argv:          # python tests/golden/make_golden_variants.py --bn   -> bn_attr_resunet.npz only
        bn_variant()
    else:
        main()

"""Golden vectors for ResAttrUnet3D2 (network.py:6-35: five poolings, widths 30/60/120/240/320/320, attention gates)
from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_attr2.py        # writes tests/golden/attr2_64.npz

The net has 69 M parameters, so -- like default_resunet_32.npz -- the weights are pinned by seed (default init under
torch.manual_seed(0); unet3d_b200 registers its parameters in the reference's order, tests/test_host_cpu.py) and the
file holds the reference's logits on one 1 x 64^3 patch (every second voxel per axis), its Dice loss, and the norm / sum
of every parameter gradient plus three full gradient tensors.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, grads_of, blocky_labels  # noqa: E402


def main():
    network, loss_mod, _ = import_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    net = network.ResAttrUnet3D2(in_channels=1, out_channels=3).eval()
    x = torch.randn(1, 1, 64, 64, 64, generator=torch.Generator().manual_seed(1234))
    y = torch.from_numpy(blocky_labels((1, 64, 64, 64), 9))
    logits = net(x)
    l = loss_mod.DiceLoss()(logits, y)
    l.backward()
    g = grads_of(net)
    names = [k for k, _ in net.named_parameters()]
    np.savez_compressed(
        os.path.join(HERE, "attr2_64.npz"),
        logits_sub=logits.detach().numpy()[:, :, ::2, ::2, ::2].astype(np.float32), loss=np.float32(l.item()),
        names=np.array(names),
        grad_norm=np.array([0.0 if g[k] is None else float(g[k].double().norm()) for k in names]),
        grad_sum=np.array([0.0 if g[k] is None else float(g[k].double().sum()) for k in names]),
        weight_sum=np.array([float(p.detach().double().sum()) for _, p in net.named_parameters()]),
        unused=np.array([k for k in names if g[k] is None]),
        grad_fc_w=g["net.fc.weight"].numpy(), grad_att0_w=g["net.up_blocks.0.att_gate.conv.weight"].numpy(),
        grad_att0_b=g["net.up_blocks.0.att_gate.conv.bias"].numpy(),
        grad_conv_w=g["net.conv.weight"].numpy(),
        x_seed=np.int64(1234), label_seed=np.int64(9), weight_seed=np.int64(0))
    print("ResAttrUnet3D2 64^3: loss", l.item(), "params", sum(p.numel() for p in net.parameters()),
          "unused", sum(1 for k in names if g[k] is None))


if __name__ == "__main__":
    main()

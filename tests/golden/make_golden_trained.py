"""Golden fixture on TRAINED weights, generated with the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_trained.py        # writes tests/golden/trained_resunet.npz (~3 MB)

VERDICT r01 item 1b: parity on a randomly initialised net measures the worst case for 16-bit storage (tiny class
margins, 42 norms deep).  What a user runs are trained weights, so this script trains the live reference
(`/root/reference/network.py` ResUnet3D(num_pool=3, num_features=8, out_channels=3) + `loss.py` HybirdLoss, Adam, train
mode with Dropout3d) for a few hundred CPU steps on the blocky nested-ellipsoid phantom (bg / kidney / tumour,
SURVEY.md 8d) and stores the trained state_dict, a held-out batch, the reference's fp32 logits on it and its loss /
per-class Dice.  tests/test_trained_parity_gpu.py then holds the CUDA path to the north-star bars on these weights.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, blocky_labels  # noqa: E402

NUM_POOL, NUM_FEATURES, PATCH, STEPS = 3, 8, 32, 400


def phantom_batch(n, seed):
    """Labels: nested ellipsoids with a jittered centre.  Image: class-dependent intensity (CT-like: background low,
    kidney mid, tumour slightly lower than kidney) + smooth bias field + white noise, then z-scored per sample."""
    lab = blocky_labels((n, PATCH, PATCH, PATCH), seed)
    g = np.random.RandomState(seed + 1000)
    means = np.array([-1.0, 0.8, 0.35], np.float32)
    img = means[lab] + 0.35 * g.standard_normal(lab.shape).astype(np.float32)
    zz = np.linspace(-1, 1, PATCH, dtype=np.float32)
    for i in range(n):
        a = g.uniform(-0.3, 0.3, 3).astype(np.float32)
        img[i] += a[0] * zz[:, None, None] + a[1] * zz[None, :, None] + a[2] * zz[None, None, :]
        img[i] = (img[i] - img[i].mean()) / (img[i].std() + 1e-8)
    return torch.from_numpy(img[:, None]), torch.from_numpy(lab)


def main():
    network, loss_mod, _ = import_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    net = network.ResUnet3D(num_pool=NUM_POOL, num_features=NUM_FEATURES, in_channels=1, out_channels=3)
    crit = loss_mod.HybirdLoss(weight_v=[1, 2, 4], alpha=0.5, beta=0.5)
    opt = torch.optim.Adam(net.parameters(), lr=2e-3)
    net.train()
    for step in range(STEPS):
        x, y = phantom_batch(2, step)
        opt.zero_grad()
        l = crit(net(x), y)
        l.backward()
        opt.step()
        if step % 50 == 0 or step == STEPS - 1:
            print(f"step {step}: loss {l.item():.4f}", flush=True)
    net.eval()
    x, y = phantom_batch(2, 10 ** 6)
    with torch.no_grad():
        logits = net(x)
        dice = [float(loss_mod.dice(torch.softmax(logits, 1)[:, c], (y == c).float())) for c in range(3)]
        loss = float(loss_mod.DiceLoss()(logits, y))
    pred = logits.argmax(1)
    print("held-out: DiceLoss", loss, "soft dice per class", dice, "label accuracy", float((pred == y).float().mean()))
    out = {"x": x.numpy(), "y": y.numpy().astype(np.uint8), "logits": logits.numpy().astype(np.float32),
           "dice_loss": np.float32(loss), "soft_dice": np.array(dice, np.float32),
           "num_pool": np.int64(NUM_POOL), "num_features": np.int64(NUM_FEATURES), "steps": np.int64(STEPS)}
    for k, v in net.state_dict().items():
        out["sd/" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "trained_resunet.npz"), **out)
    print("wrote trained_resunet.npz,", sum(p.numel() for p in net.parameters()), "parameters")


if __name__ == "__main__":
    main()

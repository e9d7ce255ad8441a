"""Whole-net parity on TRAINED weights at the north-star bars, in the benchmarked dtype.

tests/golden/trained_resunet.npz holds a ResUnet3D(num_pool=3, num_features=8, out=3) trained with the LIVE reference
(network.py + loss.py, tests/golden/make_golden_trained.py: 400 Adam steps on the nested-ellipsoid phantom), a held-out
batch and the reference's own fp32 logits on it.  The CUDA path must reproduce them to
    logits rel-L2 <= 1e-2,  argmax agreement >= 99.9 %,  per-class Dice within 1e-3
in bf16 (bench.py's dtype) AND in fp16 -- BASELINE.json's north_star tolerances, unmodified.
(On a randomly initialised net bf16 storage costs 2.2e-2 -- class margins of a random net are tiny; see
tests/test_model_gpu.py for that worst case and tests/test_block_parity_gpu.py for the backward wiring.)
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import unet3d_b200  # noqa: E402
from unet3d_b200 import ops  # noqa: E402
from oracle import unet3d_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trained_resunet.npz")


def _load():
    d = np.load(GOLD)
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd/")}
    return d, sd


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_trained_weights_meet_north_star_bars(precision):
    d, sd = _load()
    model = unet3d_b200.ResUnet3D(int(d["num_pool"]), int(d["num_features"]), 1, 3)
    model.load_state_dict(sd)
    model = model.to("cuda").eval()
    model.precision = precision
    x, y = torch.from_numpy(d["x"]), torch.from_numpy(d["y"]).long()
    ref = torch.from_numpy(d["logits"])
    with torch.no_grad():
        logits = model(x.cuda()).cpu()
    torch.cuda.synchronize()
    ops.check_device_errors()
    r = ((logits - ref).norm() / ref.norm()).item()
    agree = (logits.argmax(1) == ref.argmax(1)).float().mean().item()
    dd = (O.dice_per_class(logits, y) - O.dice_per_class(ref, y)).abs().max().item()
    dl = abs(unet3d_b200.DiceLoss()(logits.cuda(), y.cuda()).item() - float(d["dice_loss"]))
    print(f"[{precision}] trained weights: logits rel-L2 {r:.3e}, argmax agreement {agree:.6f}, Dice diff {dd:.2e}, "
          f"DiceLoss diff {dl:.2e}")
    assert r <= 1e-2, r
    assert agree >= 0.999, agree
    assert dd <= 1e-3, dd
    assert dl <= 1e-3, dl


def test_trained_weights_label_map_through_predict_per_patch():
    """The same weights through the sliding-window path: the label volume equals the argmax of the reference's logits on
    >= 99.9 % of the voxels (one window = the whole 32^3 patch, so blending is the identity)."""
    d, sd = _load()
    model = unet3d_b200.ResUnet3D(int(d["num_pool"]), int(d["num_features"]), 1, 3)
    model.load_state_dict(sd)
    model = model.to("cuda").eval()
    ref = torch.from_numpy(d["logits"]).argmax(1).numpy().astype(np.uint8)
    for i in range(2):
        vol = np.ascontiguousarray(np.moveaxis(d["x"][i], 0, -1))           # (X, Y, Z, 1)
        lab = unet3d_b200.predict_per_patch(vol, model, 3, (32, 32, 32), 2, verbose=False)
        assert lab.dtype == np.uint8 and lab.shape == ref[i].shape
        assert (lab == ref[i]).mean() >= 0.999

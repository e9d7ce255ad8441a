"""Golden vectors for the case-level resample / normalise path, from the UNMODIFIED reference.

    python tests/golden/make_golden_resample.py       # writes tests/golden/resample.npz  (build container only)

Imports /root/reference/transform.py (numpy + scipy only) and data.py's ``resample_normalize_case`` arithmetic
(data.py imports nibabel / transforms3d, absent here: the function body is exercised through ``transform.rescale``
plus the clip / z-score lines, and ``apply_scale`` is left out -- see oracle/resample_oracle.py).
Checks oracle/resample_oracle.py bit-for-bit against the live outputs before writing them.
"""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def blobs(shape, seed, classes):
    g = np.random.RandomState(seed)
    ax = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij")
    out = np.zeros(shape, dtype=np.uint8)
    for c in range(1, classes):
        ctr = g.uniform(-0.3, 0.3, size=3)
        r = sum((a - o) ** 2 for a, o in zip(ax, ctr))
        out[r < 0.6 / c] = c
    return out


def main():
    sys.path.insert(0, REF)
    import transform as T
    from oracle import resample_oracle as O

    g = np.random.RandomState(7)
    out = {}
    cases = [  # (input shape, zoom)
        ((20, 17, 9), (1.3, 0.77, 2.0)),
        ((16, 16, 8), (0.5, 0.5, 1.0)),
        ((9, 30, 11), (2.5, 1.0, 0.4)),
        ((7, 5, 1), (1.5, 1.5, 3.0)),
        ((12, 10, 6), (1 / 3, 3.05, 1.25)),      # out_len rounding: 4, 30.5 -> 30 (half to even), 7.5 -> 8
    ]
    for i, (shape, zoom) in enumerate(cases):
        img = (g.randn(*shape) * 300 + 50).astype(np.float32)
        ref = T.rescale(img[..., None], zoom, multi_class=True)
        mine = O.rescale(img[..., None], zoom, multi_class=True)
        assert ref.dtype == mine.dtype and ref.shape == mine.shape and np.array_equal(ref, mine), (i, "image")
        out[f"img{i}_in"], out[f"img{i}_zoom"], out[f"img{i}_out"] = img, np.array(zoom, dtype=np.float64), ref
        for classes in (2, 3, 4):
            lab = blobs(shape, 100 + i, classes)
            if lab.max() + 1 != classes:
                lab.flat[0] = classes - 1
            ref = T.rescale(lab, zoom, is_label=True)
            mine = O.rescale(lab, zoom, is_label=True)
            assert ref.dtype == mine.dtype and np.array_equal(ref, mine), (i, classes, "label")
            out[f"lab{i}_{classes}_in"], out[f"lab{i}_{classes}_out"] = lab, ref
    # resize (what predict_case applies to the prediction, trainer.py:128): labels and a probability volume
    lab = blobs((24, 20, 10), 5, 3)
    ref = T.resize(lab, (31, 33, 17), is_label=True)
    assert np.array_equal(ref, O.resize(lab, (31, 33, 17), is_label=True))
    out["resize_lab_in"], out["resize_lab_out"] = lab, ref
    prob = g.rand(12, 9, 7, 3).astype(np.float32)
    ref = T.resize(prob, (20, 11, 5))
    assert np.array_equal(ref, O.resize(prob, (20, 11, 5)))
    out["resize_prob_in"], out["resize_prob_out"] = prob, ref
    # resample + clip + z-score: data.py:258-275 restated with the reference's own rescale
    img = (g.randn(22, 18, 12, 1) * 400).astype(np.float32)
    stats = {"mean": 101.5, "std": 76.25, "pct_00_5": -79.0, "pct_99_5": 304.0}
    affine = np.diag([0.8, 0.8, 3.0, 1.0])
    target = (1.0, 1.2, 2.0)
    scale = np.array([0.8, 0.8, 3.0]) / np.array(target)
    image_arr = T.rescale(img, scale, multi_class=True)
    chan = T.split_dim(image_arr)[0]
    ref = np.stack([(np.clip(chan, stats["pct_00_5"], stats["pct_99_5"]) - stats["mean"]) / (stats["std"] + 1e-8)], axis=-1)
    mine = O.resample_normalize_case({"image": img, "affine": affine}, target, stats)
    assert ref.dtype == mine["image"].dtype == np.float32 and np.array_equal(ref, mine["image"])
    assert np.allclose(mine["affine"], np.diag([1.0, 1.2, 2.0, 1.0]))
    out["norm_in"], out["norm_out"] = img, ref
    out["norm_stats"] = np.array([stats["mean"], stats["std"], stats["pct_00_5"], stats["pct_99_5"]])
    out["norm_affine"], out["norm_target"] = affine, np.array(target)
    # the whole chain, live: trainer.predict_case (trainer.py:101-133) with a toy 3-class conv as the model.  data.py needs
    # transforms3d's compose / decompose (absent): the oracle's restatement is injected -- predict_case only reads the
    # spacing from the affine, the prediction does not depend on them.
    sys.path.insert(0, HERE)
    import make_golden
    import torch
    _, _, trainer = make_golden.import_reference()
    import data as ref_data
    ref_data.compose, ref_data.decompose = O.compose, O.decompose
    torch.manual_seed(3)
    toy = torch.nn.Conv3d(1, 3, 3, padding=1)
    with torch.no_grad():
        toy.weight.mul_(4.0)
    img = (np.random.RandomState(12).randn(30, 22, 14, 1) * 120 + 90).astype(np.float32)
    affine = np.diag([1.5, 1.0, 2.0, 1.0])
    target = (1.0, 1.16, 1.0)                         # resampled grid 45 x 19 x 28; patch 24 on y: odd pad / crop
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lab = trainer.predict_case({"image": img.copy(), "affine": affine.copy()}, toy, target, stats, num_classes=3,
                                   patch_size=(16, 24, 16), step_per_patch=2, verbose=False)["pred"]
        prob = trainer.predict_case({"image": img.copy(), "affine": affine.copy()}, toy, target, stats, num_classes=3,
                                    patch_size=(16, 24, 16), step_per_patch=2, verbose=False, one_hot=True)["pred"]
    assert lab.shape == img.shape[:3] and prob.shape == img.shape[:3] + (3,)
    print("predict_case:", lab.dtype, lab.shape, np.bincount(lab.ravel()), prob.dtype, "nan frac", float(np.isnan(prob).mean()))
    out["case_image"], out["case_affine"], out["case_target"] = img, affine, np.array(target)
    out["case_w"], out["case_b"] = toy.weight.detach().numpy(), toy.bias.detach().numpy()
    out["case_labels"], out["case_probs"] = lab, prob
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **out)
    print("wrote resample.npz:", len(out), "arrays,", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()

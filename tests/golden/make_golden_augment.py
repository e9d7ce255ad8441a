"""Golden vectors for the train-loader augmentation, from the UNMODIFIED reference transforms (build container only).

    python tests/golden/make_golden_augment.py        # writes tests/golden/augment.npz

The live classes of /root/reference/transform.py -- RandomRescaleCrop(0.1, crop, crop_mode='random'), RandomMirror,
RandomContrast / RandomBrightness / RandomGamma(0.1), composed as in nb_train_iib.py:27-36 -- are run on one synthetic
case under several numpy seeds; each seed is run twice, with and without the final gamma stage (same draws up to there), so
that the stages that must match bit for bit are separated from np.power.  A second configuration enforces label 2 in
the crop (the retry loop) on a centre crop.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, blocky_labels  # noqa: E402

CROP = (24, 24, 24)
SEEDS = [0, 1, 2, 3, 4, 5]


def make_case():
    g = np.random.RandomState(42)
    shape = (44, 40, 36)
    lab = blocky_labels((1, *shape), 3)[0].astype(np.uint8)
    img = (np.array([-1.0, 0.7, 0.2], np.float32)[lab] + 0.4 * g.standard_normal(shape).astype(np.float32))
    xx = np.linspace(-1, 1, shape[0], dtype=np.float32)
    img = (img + 0.3 * xx[:, None, None]).astype(np.float32)
    return img[..., None], lab


def main():
    import_reference()
    import transform as RT                      # /root/reference/transform.py
    if not hasattr(np, "int"):
        np.int = int
    image, label = make_case()
    out = {"image": image, "label": label, "crop": np.array(CROP), "seeds": np.array(SEEDS)}

    def run(seed, stages):
        np.random.seed(seed)
        case = {"image": image.copy(), "label": label.copy()}
        for t in stages:
            case = t(case)
        return case["image"], case["label"]

    for s in SEEDS:
        base = lambda: [RT.RandomRescaleCrop(0.1, CROP, crop_mode='random'), RT.RandomMirror((0.5, 0.5, 0.5)),
                        RT.RandomContrast(0.1), RT.RandomBrightness(0.1)]
        img_b, lab_b = run(s, base())
        img_g, lab_g = run(s, base() + [RT.RandomGamma(0.1)])
        assert np.array_equal(lab_b, lab_g)
        out[f"s{s}/pre_gamma"], out[f"s{s}/image"], out[f"s{s}/label"] = img_b, img_g, lab_g
        # the crop alone (no intensity stages): pins the zoom of image and label
        img_c, lab_c = run(s, [RT.RandomRescaleCrop(0.1, CROP, crop_mode='random')])
        out[f"s{s}/crop_image"], out[f"s{s}/crop_label"] = img_c, lab_c
    # enforced label in a random crop small enough to miss it sometimes
    for s in SEEDS[:3]:
        img_e, lab_e = run(s, [RT.RandomRescaleCrop(0.2, (12, 12, 12), crop_mode='random', enforce_label_indices=[2]),
                               RT.RandomMirror(0.5)])
        out[f"e{s}/image"], out[f"e{s}/label"] = img_e, lab_e
    # label re-coding (transform.py:323-384), as nb_train_* scripts use it
    out["combine_12"] = RT.combination_labels(label.copy(), [[1, 2]], 3)
    out["combine_01"] = RT.combination_labels(label.copy(), [0, 1], 3)
    out["onehot"] = RT.to_one_hot(label[:6, :5, :4].copy(), 3)
    out["onehot_t"] = RT.to_one_hot(label[:6, :5, :4].copy(), 3, to_tensor=True)
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **out)
    print("wrote augment.npz:", {k: v.shape for k, v in out.items() if k.startswith("s0") or k.startswith("e0")})


if __name__ == "__main__":
    main()

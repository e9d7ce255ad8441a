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
    # ---- cascade (trainer.py:164-245), live: two bright blobs + a speck; the coarse "model" is a smoothing conv with one
    # output channel (sigmoid), the detail model the 3-class toy conv above.  Also the regions the reference cuts out of
    # the coarse prediction (data.regions_crop_case: ndi.label / find_objects / remove_small_region) -- integer work.
    ref_data.apply_translate.__globals__["compose"], ref_data.apply_translate.__globals__["decompose"] = O.compose, O.decompose
    rs = np.random.RandomState(21)
    shape = (40, 36, 20)
    ax = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    img = rs.randn(*shape).astype(np.float32) * 20
    for ctr, rad in (((11, 12, 8), (7, 6, 4)), ((29, 24, 12), (6, 8, 5)), ((35, 5, 3), (1.2, 1.2, 1.2))):
        d = sum(((a - c) / r) ** 2 for a, c, r in zip(ax, ctr, rad))
        img[d < 1] += 300
    img = img[..., None]
    coarse = torch.nn.Conv3d(1, 1, 3, padding=1)
    with torch.no_grad():
        coarse.weight.fill_(1.0 / 27)
        coarse.bias.fill_(-1.0)
    detail = toy
    detail.out_channels = 3
    c_stats = {"mean": 50.0, "std": 100.0, "pct_00_5": -100.0, "pct_99_5": 400.0}
    caff = np.diag([1.0, 1.0, 2.0, 1.0])
    c_target, d_target = (1.6, 1.6, 2.5), (1.0, 1.0, 1.6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ccase = trainer.predict_case({"image": img.copy(), "affine": caff.copy(), "case_id": "c"}, coarse, c_target, c_stats, 1,
                                     (16, 16, 8), 2, verbose=False)
        regions = ref_data.regions_crop_case(ccase, 60, 3, "pred")
        # trainer.cascade_predict_case itself cannot run under numpy >= 1.23 (trainer.py:225-226 index with a LIST of
        # slices), so its body is followed by hand: live regions, live per-region predict_case, and the merge lines
        # (trainer.py:189-241) through the oracle's restatement, which differs only in tuple(...) around the slices
        preds = []
        for region in regions:
            r = trainer.predict_case(dict(region), detail, d_target, stats, 3, (16, 24, 16), 2, verbose=False, one_hot=True)
            preds.append((region["bbox"], r["pred"]))
        full = {"pred": O.merge_regions(img.shape[:3], 3, preds)}
    print("cascade: coarse voxels", int(ccase["pred"].sum()), "regions", [r["bbox"].tolist() for r in regions],
          "labels", np.bincount(full["pred"].ravel()))
    assert len(regions) == 2
    out["casc_image"], out["casc_affine"] = img, caff
    out["casc_coarse_pred"] = ccase["pred"]
    out["casc_bboxes"] = np.array([r["bbox"] for r in regions])
    out["casc_region0_image"] = regions[0]["image"]
    out["casc_region_affines"] = np.array([r["affine"] for r in regions])
    out["casc_coarse_w"], out["casc_coarse_b"] = coarse.weight.detach().numpy(), coarse.bias.detach().numpy()
    out["casc_c_stats"] = np.array([c_stats["mean"], c_stats["std"], c_stats["pct_00_5"], c_stats["pct_99_5"]])
    out["casc_c_target"], out["casc_d_target"] = np.array(c_target), np.array(d_target)
    out["casc_pred"] = full["pred"]
    # oracle restatement of the glue, on the reference's own coarse prediction
    mine = O.regions_crop_case({"image": img, "affine": caff, "pred": ccase["pred"], "case_id": "c"}, 60, 3, "pred")
    assert len(mine) == 2 and all(np.array_equal(a["bbox"], b["bbox"]) and np.array_equal(a["image"], b["image"])
                                  and np.allclose(a["affine"], b["affine"]) for a, b in zip(mine, regions))
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **out)
    print("wrote resample.npz:", len(out), "arrays,", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()

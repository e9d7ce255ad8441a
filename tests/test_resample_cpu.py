"""CPU: the resample oracle (oracle/resample_oracle.py) against the golden vectors made from the live reference
(tests/golden/make_golden_resample.py: transform.rescale / resize, the clip + z-score of resample_normalize_case,
trainer.predict_case), and the host-side shape / affine logic of unet3d_b200.transform."""
import os

import numpy as np
import pytest
import torch

import unet3d_b200
from oracle import resample_oracle as R
from oracle import unet3d_oracle as O


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "resample.npz"))


def test_oracle_zoom_bit_exact_with_reference(gold):
    for i in range(5):
        img, zoom = gold[f"img{i}_in"], gold[f"img{i}_zoom"]
        got = R.rescale(img[..., None], zoom, multi_class=True)
        assert got.dtype == np.float32 and np.array_equal(got, gold[f"img{i}_out"]), i
        for classes in (2, 3, 4):
            lab = gold[f"lab{i}_{classes}_in"]
            got = R.rescale(lab, zoom, is_label=True)
            assert got.dtype == np.uint8 and np.array_equal(got, gold[f"lab{i}_{classes}_out"]), (i, classes)
    assert np.array_equal(R.resize(gold["resize_lab_in"], (31, 33, 17), is_label=True), gold["resize_lab_out"])
    assert np.array_equal(R.resize(gold["resize_prob_in"], (20, 11, 5)), gold["resize_prob_out"])


def _stats(gold):
    m, s, lo, hi = (float(v) for v in gold["norm_stats"])
    return {"mean": m, "std": s, "pct_00_5": lo, "pct_99_5": hi}


def test_oracle_resample_normalize(gold):
    case = R.resample_normalize_case({"image": gold["norm_in"], "affine": gold["norm_affine"]},
                                     tuple(gold["norm_target"]), _stats(gold))
    assert case["image"].dtype == np.float32 and np.array_equal(case["image"], gold["norm_out"])
    assert np.allclose(R.get_spacing(case["affine"]), gold["norm_target"])


def test_oracle_predict_case_chain(gold):
    """resample_normalize_case -> predict_per_patch -> resize, restated, against the live trainer.predict_case."""
    w, b = torch.from_numpy(gold["case_w"]), torch.from_numpy(gold["case_b"])
    fn = lambda t: torch.nn.functional.conv3d(t, w, b, padding=1)
    case = {"image": gold["case_image"], "affine": gold["case_affine"]}
    rs = R.resample_normalize_case(case, tuple(gold["case_target"]), _stats(gold))
    lab = O.predict_per_patch(rs["image"], fn, 3, (16, 24, 16), 2)
    lab = R.resize(lab, case["image"].shape[:3], is_label=True)
    assert (lab != gold["case_labels"]).mean() < 2e-3
    prob = O.predict_per_patch(rs["image"], fn, 3, (16, 24, 16), 2, one_hot=True)
    prob = R.resize(prob, case["image"].shape[:3])
    assert np.array_equal(np.isnan(prob), np.isnan(gold["case_probs"]))
    assert np.allclose(np.nan_to_num(prob), np.nan_to_num(gold["case_probs"]), atol=1e-5)


def test_zoomed_shape_rounds_half_to_even():
    T = unet3d_b200.transform
    assert T.zoomed_shape((12, 10, 6), (1 / 3, 3.05, 1.25)) == (4, 30, 8)
    assert T.zoomed_shape((5, 7, 9), (0.5, 0.5, 0.5)) == (2, 4, 4)          # 2.5 -> 2, 3.5 -> 4, 4.5 -> 4
    rng = np.random.RandomState(0)
    for _ in range(100):
        shape = tuple(int(v) for v in rng.randint(1, 400, 3))
        zoom = tuple(rng.uniform(0.2, 3.0, 3))
        assert T.zoomed_shape(shape, zoom) == R.zoomed_shape(shape, zoom)


def test_spacing_and_apply_scale_match_oracle():
    rng = np.random.RandomState(1)
    for _ in range(20):
        q, _ = np.linalg.qr(rng.randn(3, 3))
        A = np.eye(4)
        A[:3, :3] = q @ np.diag(rng.uniform(0.5, 3.0, 3))
        A[:3, 3] = rng.randn(3) * 50
        s = rng.uniform(0.3, 2.5, 3)
        assert np.allclose(unet3d_b200.apply_scale(A, s), R.apply_scale(A, s), atol=1e-12)
        assert np.allclose(unet3d_b200.get_spacing(A), R.get_spacing(A))
        # scaling the zooms scales the column norms: spacing of the rescaled grid
        assert np.allclose(np.linalg.norm(unet3d_b200.apply_scale(A, s)[:3, :3], axis=0), np.linalg.norm(A[:3, :3], axis=0) * s)


def test_transform_has_no_cpu_path():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        unet3d_b200.rescale(np.zeros((4, 4, 4), np.float32), 2.0)
    with pytest.raises(NotImplementedError):
        unet3d_b200.rescale(np.zeros((4, 4, 4), np.float32), 2.0, order=3)


def test_oracle_cascade_glue_against_reference_golden(gold):
    """data.regions_crop_case restated (label / remove_small_region / find_objects / crop_pad_to_bbox / apply_translate)
    on the live reference's coarse prediction: boxes, crops and affines as committed by make_golden_resample.py."""
    case = {"image": gold["casc_image"], "affine": gold["casc_affine"], "pred": gold["casc_coarse_pred"], "case_id": "c"}
    regions = R.regions_crop_case(case, 60, 3, "pred")
    assert len(regions) == len(gold["casc_bboxes"]) == 2
    for r, bbox, aff in zip(regions, gold["casc_bboxes"], gold["casc_region_affines"]):
        assert np.array_equal(r["bbox"], bbox) and np.allclose(r["affine"], aff)
    assert np.array_equal(regions[0]["image"], gold["casc_region0_image"])
    # the speck below the threshold is a component of its own before filtering
    import scipy.ndimage as ndi
    lab, n = R.label_components(gold["casc_coarse_pred"] > 0)
    lab2, n2 = ndi.label(gold["casc_coarse_pred"] > 0)
    assert n == n2 and np.array_equal(lab, lab2)


def test_oracle_merge_regions_semantics():
    rng = np.random.RandomState(4)
    preds = [(np.array([[-2, 6], [1, 7], [0, 5]]), rng.rand(8, 6, 5, 3).astype(np.float32)),
             (np.array([[3, 12], [2, 9], [2, 8]]), rng.rand(9, 7, 6, 3).astype(np.float32))]
    preds[1][1][0, 0, 0] = np.nan
    out = R.merge_regions((10, 8, 6), 3, preds)
    assert out.shape == (10, 8, 6) and out.dtype == np.uint8
    assert out[9, 0, 0] == 0 and out[3, 2, 2] == 0                     # uncovered -> 0; NaN -> first class
    assert out[0, 1, 0] == preds[0][1][2, 0, 0].argmax()               # box starts at -2: region voxel 2 lands on voxel 0
    one = R.merge_regions((10, 8, 6), 1, [(b, p[..., :1]) for b, p in preds])
    assert set(np.unique(one)) <= {0, 1}


def test_crop_pad_to_bbox_and_translate_match_oracle():
    """Host-side index math of regions_crop_case (transform.py:422-437, data.py:68-71): numpy path of the product helper
    against the oracle's restatement, boxes partly and fully outside the volume included."""
    rng = np.random.RandomState(11)
    vol = rng.randn(9, 7, 5, 2).astype(np.float32)
    lab = rng.randint(0, 3, (9, 7, 5)).astype(np.uint8)
    for _ in range(50):
        lo = [int(rng.randint(-6, s + 2)) for s in lab.shape]
        bbox = np.array([[l, l + int(rng.randint(1, 9))] for l in lo])
        got = unet3d_b200.crop_pad_to_bbox(lab, bbox)
        assert got.dtype == lab.dtype and got.shape == tuple(b[1] - b[0] for b in bbox)
        inside = all(min(b[1], s) > max(b[0], 0) for b, s in zip(bbox, lab.shape))
        if inside:                                  # the reference (np.pad of an empty crop) is only defined for overlapping boxes
            assert np.array_equal(got, R.crop_pad_to_bbox(lab, bbox))
            bbox_c = np.concatenate([bbox, [[0, 2]]])
            assert np.array_equal(unet3d_b200.crop_pad_to_bbox(vol, bbox_c), R.crop_pad_to_bbox(vol, bbox_c))
        else:
            assert not got.any()
    A = np.diag([0.8, 1.2, 2.5, 1.0])
    A[:3, 3] = [4.0, -2.0, 7.5]
    assert np.allclose(unet3d_b200.apply_translate(A, [1.0, 2.0, -3.0]), R.apply_translate(A, [1.0, 2.0, -3.0]))
    t = unet3d_b200.transform.normalize_table({"mean": 101.5, "std": 76.25, "pct_00_5": -79.0, "pct_99_5": 304.0})
    assert t == [(np.float32(-79.0), np.float32(304.0), np.float32(101.5), np.float32(76.25 + 1e-8))]

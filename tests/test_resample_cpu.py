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

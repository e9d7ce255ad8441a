"""GPU parity of the case-level resampling kernels (csrc/resample.cu) through the C ABI: bit-exact against the golden
vectors made by the live reference (transform.rescale / resize = scipy.ndimage.zoom order 1) and against the oracle
on seeded shapes; the fused clip + z-score; predict_case end to end; full-size properties (identity, exact scaling)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import unet3d_b200  # noqa: E402
from unet3d_b200 import transform as T  # noqa: E402
from oracle import resample_oracle as R  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "resample.npz"))


def _stats(gold):
    m, s, lo, hi = (float(v) for v in gold["norm_stats"])
    return {"mean": m, "std": s, "pct_00_5": lo, "pct_99_5": hi}


def test_rescale_bit_exact_with_reference_golden(gold):
    for i in range(5):
        img, zoom = gold[f"img{i}_in"], gold[f"img{i}_zoom"]
        got = unet3d_b200.rescale(img[..., None], zoom, multi_class=True)
        assert got.dtype == np.float32 and np.array_equal(got, gold[f"img{i}_out"]), i
        assert np.array_equal(unet3d_b200.rescale(img, zoom), gold[f"img{i}_out"][..., 0]), i
        for classes in (2, 3, 4):            # 2: zoom as float32 + truncation; 3, 4: one-hot zoom + argmax
            got = unet3d_b200.rescale(gold[f"lab{i}_{classes}_in"], zoom, is_label=True)
            assert got.dtype == np.uint8 and np.array_equal(got, gold[f"lab{i}_{classes}_out"]), (i, classes)
    assert np.array_equal(unet3d_b200.resize(gold["resize_lab_in"], (31, 33, 17), is_label=True), gold["resize_lab_out"])
    assert np.array_equal(unet3d_b200.resize(gold["resize_prob_in"], (20, 11, 5)), gold["resize_prob_out"])


def test_resample_normalize_case_bit_exact(gold):
    case = unet3d_b200.resample_normalize_case({"image": gold["norm_in"], "affine": gold["norm_affine"]},
                                               tuple(gold["norm_target"]), _stats(gold))
    assert case["image"].dtype == np.float32 and np.array_equal(case["image"], gold["norm_out"])
    assert np.allclose(unet3d_b200.get_spacing(case["affine"]), gold["norm_target"])


@pytest.mark.parametrize("seed", range(6))
def test_rescale_vs_oracle_seeded(seed):
    rng = np.random.RandomState(seed)
    shape = tuple(int(v) for v in rng.randint(1, 40, 3))
    zoom = tuple(rng.uniform(0.3, 2.7, 3))
    if min(R.zoomed_shape(shape, zoom)) < 1:
        zoom = (1.0, 1.0, 1.0)
    img = (rng.randn(*shape, 2) * 1000).astype(np.float32)
    assert np.array_equal(unet3d_b200.rescale(img, zoom, multi_class=True), R.rescale(img, zoom, multi_class=True))
    for classes in (2, 3, 5, 9):
        lab = rng.randint(0, classes, shape).astype(np.uint8)            # salt-and-pepper: every corner differs
        lab.flat[0] = classes - 1
        assert np.array_equal(unet3d_b200.rescale(lab, zoom, is_label=True), R.rescale(lab, zoom, is_label=True)), classes
    i16 = (rng.randn(*shape) * 500).astype(np.int16)                      # non-float image: zoom as float32, cast back
    assert np.array_equal(unet3d_b200.rescale(i16, zoom), R.rescale(i16, zoom))


def test_strided_destination_and_nan():
    """The destination may be any view (predict_case writes the interior of the padded NCDHW model input); NaN inputs
    (uncovered border of the reference tile grid) propagate like in SciPy, zero weights included."""
    rng = np.random.RandomState(3)
    img = rng.randn(9, 8, 7, 2).astype(np.float32)
    img[2, 3, 4, 0] = np.nan
    zoom = (1.0, 1.5, 2.0)
    want = R.rescale(img, zoom, multi_class=True)
    x = torch.full((1, 2, 14, 15, 16), -7.0, device=DEV)
    view = x[0, :, 3:12, 2:14, 1:15].permute(1, 2, 3, 0)
    T.rescale_device(torch.from_numpy(img).to(DEV), zoom, multi_class=True, out=view)
    got = view.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(want).sum() > 1
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(want))
    x[0, :, 3:12, 2:14, 1:15] = -7.0
    assert bool((x == -7.0).all())                                       # nothing outside the view was written


class _ToyNet(torch.nn.Module):
    def __init__(self, w, b):
        super().__init__()
        self.w, self.b = torch.nn.Parameter(w), torch.nn.Parameter(b)

    def forward(self, x):
        return torch.nn.functional.conv3d(x, self.w, self.b, padding=1)


def test_predict_case_matches_live_reference_golden(gold):
    """trainer.predict_case of the live reference (toy conv model): labels and probabilities on the original grid,
    including the odd pad / crop offset and the NaN border."""
    torch.backends.cudnn.allow_tf32 = False
    net = _ToyNet(torch.from_numpy(gold["case_w"]), torch.from_numpy(gold["case_b"])).to(DEV)
    case = {"image": gold["case_image"].copy(), "affine": gold["case_affine"].copy()}
    out = unet3d_b200.predict_case(dict(case), net, tuple(gold["case_target"]), _stats(gold), num_classes=3,
                                   patch_size=(16, 24, 16), step_per_patch=2, verbose=False)
    assert out["pred"].dtype == np.uint8 and out["pred"].shape == gold["case_labels"].shape
    assert (out["pred"] != gold["case_labels"]).mean() < 2e-3
    assert np.array_equal(out["affine"], gold["case_affine"])
    out = unet3d_b200.predict_case(dict(case), net, tuple(gold["case_target"]), _stats(gold), num_classes=3,
                                   patch_size=(16, 24, 16), step_per_patch=2, verbose=False, one_hot=True)
    assert np.array_equal(np.isnan(out["pred"]), np.isnan(gold["case_probs"]))
    assert np.allclose(np.nan_to_num(out["pred"]), np.nan_to_num(gold["case_probs"]), atol=1e-5)


def test_predict_case_with_the_unet_equals_the_staged_pipeline():
    """predict_case (everything in HBM) == resample_normalize_case -> predict_per_patch -> resize, each through the host."""
    torch.manual_seed(0)
    model = unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3).to(DEV)
    rng = np.random.RandomState(5)
    case = {"image": (rng.randn(24, 20, 12, 1) * 150 + 40).astype(np.float32), "affine": np.diag([1.0, 1.4, 2.5, 1.0])}
    stats = {"mean": 40.0, "std": 150.0, "pct_00_5": -300.0, "pct_99_5": 350.0}
    got = unet3d_b200.predict_case(dict(case), model, (1.0, 1.0, 1.25), stats, 3, (16, 16, 16), 2, verbose=False,
                                   grid_mode="full_cover")["pred"]
    staged = unet3d_b200.resample_normalize_case(dict(case), (1.0, 1.0, 1.25), stats)
    lab = unet3d_b200.predict_per_patch(staged["image"], model, 3, (16, 16, 16), 2, verbose=False, grid_mode="full_cover")
    want = unet3d_b200.resize(lab, case["image"].shape[:3], is_label=True)
    assert got.shape == want.shape == case["image"].shape[:3]
    assert (got != want).mean() < 2e-3              # fp64 statistics atomics: run-to-run differences at exact ties only


def test_full_size_properties():
    """BASELINE cfg-4 volume (512 x 512 x 256): zoom 1 is the identity; scaling by 2 commutes exactly; upsampling labels by
    an integer-ratio grid and sampling them back returns the input (the coarse samples hit the original voxels)."""
    g = torch.Generator(device=DEV).manual_seed(0)
    vol = torch.randn(512, 512, 256, device=DEV, generator=g)
    same = T.rescale_device(vol, 1.0)
    assert torch.equal(same, vol)
    z = (0.8, 1.25, 1.5)
    a = T.rescale_device(vol, z)
    assert tuple(a.shape) == T.zoomed_shape(vol.shape, z)
    assert torch.equal(T.rescale_device(vol * 2, z), a * 2)
    lab = torch.randint(0, 3, (129, 129, 65), device=DEV, generator=g).to(torch.uint8)
    up = T.rescale_device(lab, (257 / 129, 257 / 129, 129 / 65), is_label=True, num_classes=3)       # step exactly 1/2
    assert tuple(up.shape) == (257, 257, 129)
    assert torch.equal(up[::2, ::2, ::2], lab)
    back = T.rescale_device(up, (129 / 257, 129 / 257, 65 / 129), is_label=True, num_classes=3)        # step exactly 2
    assert torch.equal(back, lab)

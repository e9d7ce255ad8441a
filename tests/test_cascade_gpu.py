"""GPU parity of the cascade glue (csrc/regions.cu) through the C ABI: connected components against SciPy (numbering,
sizes, boxes: bit-exact), regions_crop_case and the whole cascade against the golden vectors of the live reference,
the merge against the oracle."""
import os

import numpy as np
import pytest
import scipy.ndimage as ndi
import torch

pytestmark = pytest.mark.gpu

import unet3d_b200  # noqa: E402
from unet3d_b200 import ops, transform as T  # noqa: E402
from oracle import resample_oracle as R  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "resample.npz"))


def _check_components(mask_np):
    labels, roots, stats = ops.connected_components(torch.from_numpy(mask_np.astype(np.uint8)).to(DEV))
    want, n = ndi.label(mask_np)
    assert roots.numel() == n
    lab = labels.cpu().numpy().reshape(-1)
    rank = np.zeros(mask_np.size + 1, np.int32)
    rank[roots.cpu().numpy()] = np.arange(1, n + 1)
    got = np.where(lab >= 0, rank[lab], 0).reshape(mask_np.shape)
    assert np.array_equal(got, want)                                        # scipy's numbering, voxel for voxel
    st = stats.cpu().numpy()
    assert np.array_equal(st[:, 0], np.bincount(want.ravel(), minlength=n + 1)[1:])
    for i, sl in enumerate(ndi.find_objects(want)):
        assert [st[i, 1], st[i, 2] + 1, st[i, 3], st[i, 4] + 1, st[i, 5], st[i, 6] + 1] == \
            [sl[0].start, sl[0].stop, sl[1].start, sl[1].stop, sl[2].start, sl[2].stop]


@pytest.mark.parametrize("shape,density", [((9, 8, 7), 0.5), ((33, 20, 65), 0.3), ((40, 41, 37), 0.62), ((5, 1, 1), 0.9),
                                           ((1, 1, 70), 0.7), ((24, 24, 24), 0.0), ((24, 24, 24), 1.0)])
def test_connected_components_match_scipy(shape, density):
    rng = np.random.RandomState(sum(shape))
    _check_components(rng.rand(*shape) < density)


def test_connected_components_full_size_blobs():
    """cfg-4 sized mask (512 x 512 x 256): a few hundred ellipsoids, spirals through x / y, specks."""
    rng = np.random.RandomState(3)
    g = torch.Generator(device=DEV).manual_seed(0)
    vol = torch.rand(64, 64, 32, device=DEV, generator=g)
    vol = torch.nn.functional.interpolate(vol[None, None], size=(512, 512, 256), mode="trilinear")[0, 0]
    mask = (vol > 0.55).cpu().numpy()
    mask[rng.randint(0, 512, 500), rng.randint(0, 512, 500), rng.randint(0, 256, 500)] = True
    _check_components(mask)


def test_regions_crop_case_matches_reference_golden(gold):
    case = {"image": gold["casc_image"], "affine": gold["casc_affine"], "pred": gold["casc_coarse_pred"], "case_id": "c"}
    regions = unet3d_b200.regions_crop_case(case, 60, 3, "pred")
    assert len(regions) == 2 and [r["case_id"] for r in regions] == ["c_000", "c_001"]
    for r, bbox, aff in zip(regions, gold["casc_bboxes"], gold["casc_region_affines"]):
        assert np.array_equal(r["bbox"], bbox) and np.allclose(r["affine"], aff)
    assert np.array_equal(regions[0]["image"], gold["casc_region0_image"])
    # threshold 0 keeps the speck as a third region, in raster order
    assert len(unet3d_b200.regions_crop_case(case, 0, 3, "pred")) == ndi.label(gold["casc_coarse_pred"] > 0)[1]


@pytest.mark.parametrize("K", [1, 3])
def test_merge_matches_oracle(K):
    rng = np.random.RandomState(K)
    shape = (20, 17, 9)
    preds = []
    for _ in range(5):
        lo = [rng.randint(-4, s - 3) for s in shape]
        hi = [l + rng.randint(3, 12) for l in lo]
        p = rng.rand(*[h - l for l, h in zip(lo, hi)], K).astype(np.float32)
        p[rng.randint(p.shape[0]), rng.randint(p.shape[1]), :, :] = np.nan
        preds.append((np.array(list(zip(lo, hi))), p))
    want = R.merge_regions(shape, K, preds)
    result = torch.zeros((*shape, K), dtype=torch.float64, device=DEV)
    count = torch.zeros(shape, dtype=torch.int32, device=DEV)
    for bbox, p in preds:
        src0 = [max(-int(bbox[i][0]), 0) for i in range(3)]
        dst0 = [max(int(bbox[i][0]), 0) for i in range(3)]
        box = [min(int(bbox[i][1]), shape[i]) - dst0[i] for i in range(3)]
        ops.region_accumulate(torch.from_numpy(p).to(DEV), result, count, src0, dst0, box)
    got = ops.merge_finalize(result, count).cpu().numpy()
    if K == 1:       # NaN -> uint8 is platform-defined in numpy; compare where the mean is a number
        ok = ~np.isnan((result[..., 0] / count.clamp(min=1)).cpu().numpy())
        assert np.array_equal(got[ok], want[ok])
    else:
        assert np.array_equal(got, want)


class _ToyNet(torch.nn.Module):
    def __init__(self, w, b):
        super().__init__()
        self.w, self.b = torch.nn.Parameter(w), torch.nn.Parameter(b)
        self.out_channels = w.shape[0]

    def forward(self, x):
        return torch.nn.functional.conv3d(x, self.w, self.b, padding=1)


def test_cascade_predict_case_matches_reference_golden(gold):
    torch.backends.cudnn.allow_tf32 = False
    coarse = _ToyNet(torch.from_numpy(gold["casc_coarse_w"]), torch.from_numpy(gold["casc_coarse_b"])).to(DEV)
    detail = _ToyNet(torch.from_numpy(gold["case_w"]), torch.from_numpy(gold["case_b"])).to(DEV)
    cs = dict(zip(("mean", "std", "pct_00_5", "pct_99_5"), (float(v) for v in gold["casc_c_stats"])))
    ds = dict(zip(("mean", "std", "pct_00_5", "pct_99_5"), (float(v) for v in gold["norm_stats"])))
    case = {"image": gold["casc_image"].copy(), "affine": gold["casc_affine"].copy(), "case_id": "c"}
    # the coarse stage alone: the one-class (sigmoid) prediction on the original grid
    cp = unet3d_b200.predict_case(dict(case), coarse, tuple(gold["casc_c_target"]), cs, 1, (16, 16, 8), 2, verbose=False)
    assert (cp["pred"] != gold["casc_coarse_pred"]).mean() < 1e-3
    out = unet3d_b200.cascade_predict_case(dict(case), coarse, tuple(gold["casc_c_target"]), cs, (16, 16, 8), detail,
                                           tuple(gold["casc_d_target"]), ds, (16, 24, 16), num_classes=3, step_per_patch=2,
                                           region_threshold=60, crop_padding=3, verbose=False)
    assert out["pred"].dtype == np.uint8 and out["pred"].shape == gold["casc_pred"].shape
    assert (out["pred"] != gold["casc_pred"]).mean() < 2e-3
    assert out["pred"].max() == 2 and (out["pred"] > 0).sum() > 1000

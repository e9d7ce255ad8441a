"""CPU oracle for the case-level pre/post-processing around the window loop.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy float64 arithmetic, what the reference does either side of ``predict_per_patch``
(SURVEY.md 8f rank 2):

* ``transform.rescale`` / ``transform.resize`` (transform.py:32-100) -- wrappers around
  ``scipy.ndimage.zoom(order=1, mode='reflect')``; labels with >= 3 classes are zoomed as one float one-hot volume
  per class and arg-maxed, labels with < 3 classes are zoomed as float32 and truncated back to the label dtype;
* ``data.resample_normalize_case`` (data.py:223-284) -- rescale to the target spacing, clip to the 0.5 / 99.5
  percentiles, z-score;
* ``trainer.predict_case`` (trainer.py:101-133) -- the three steps chained.

The arithmetic lives in a third-party dependency that is not vendored under /root/reference: SciPy's
``ndimage.zoom`` (the reference pins no version; 1.18.1 is installed here).  Its published algorithm for
``order=1, grid_mode=False`` (``ni_interpolation.c: NI_ZoomShift``) is restated in ``zoom_linear``:

    out_len = round(in_len * zoom)                              (Python ``round``: half to even)
    step    = (in_len - 1) / (out_len - 1)   (1 when out_len == 1)     -- NOT the caller's zoom factor
    cc      = o * step                       (float64; mode 'reflect' leaves 0 <= cc < len untouched)
    start   = floor(cc);  x = cc - start;    weights (1 - x, x)
    idx     = start, start + 1, an index == len reflected to len - 1
    t       = sum over the 2^rank corners, last axis fastest, of ((v * w0) * w1) * w2      (float64)
    out     = float32(t)

Pinning: the reference ships no tests or vectors for this path, so **parity is unpinned by reference tests**; the
restatement is pinned bit-for-bit against the live ``transform.rescale / resize`` (imported from /root/reference,
which calls the installed SciPy) by ``tests/golden/make_golden_resample.py`` and against that script's committed
outputs (``tests/golden/resample.npz``).  ``apply_scale`` depends on ``transforms3d`` (absent here): its published
decompose / compose algorithm is restated and only self-checked (parity unpinned).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def zoomed_shape(shape: Sequence[int], zoom: Sequence[float]) -> tuple:
    """scipy.ndimage.zoom: ``tuple(int(round(ii * jj)) ...)`` (Python round, half to even)."""
    return tuple(int(round(float(i) * float(z))) for i, z in zip(shape, zoom))


def _axis_tables(n_in: int, n_out: int):
    step = (n_in - 1) / (n_out - 1) if n_out > 1 else 1.0
    cc = np.arange(n_out, dtype=np.float64) * np.float64(step)
    start = np.floor(cc)
    x = cc - start
    i0 = start.astype(np.int64)
    i1 = i0 + 1
    i1 = np.where(i1 >= n_in, 2 * n_in - i1 - 1, i1)     # 'reflect': d c b a | a b c d | d c b a
    i1 = np.clip(i1, 0, n_in - 1)                        # n_in == 1
    return i0, i1, 1.0 - x, x


def zoom_linear(vol: np.ndarray, zoom: Sequence[float]) -> np.ndarray:
    """``ndi.zoom(vol.astype(float32), zoom, order=1, mode='reflect')`` for a 3-D volume; float32 result."""
    v = np.asarray(vol, dtype=np.float32).astype(np.float64)
    if np.isscalar(zoom):
        zoom = [zoom] * v.ndim
    oshape = zoomed_shape(v.shape, zoom)
    tabs = [_axis_tables(v.shape[d], oshape[d]) for d in range(3)]
    t = np.zeros(oshape, dtype=np.float64)
    for a in (0, 1):
        ia, wa = tabs[0][a], tabs[0][2 + a]
        for b in (0, 1):
            ib, wb = tabs[1][b], tabs[1][2 + b]
            for c in (0, 1):
                ic, wc = tabs[2][c], tabs[2][2 + c]
                corner = v[np.ix_(ia, ib, ic)]
                t += ((corner * wa[:, None, None]) * wb[None, :, None]) * wc[None, None, :]
    return t.astype(np.float32)


def rescale(input: np.ndarray, scale, is_label: bool = False, multi_class: bool = False) -> np.ndarray:
    """transform.py:32-78 with the reference's defaults order=1, mode='reflect', cval=0."""
    dtype = input.dtype
    if is_label:
        num_classes = int(np.unique(input).max()) + 1
    if not is_label or num_classes < 3:
        if multi_class:                                            # (X, Y, Z, C): every channel on its own
            chans = [zoom_linear(input[..., c], scale) for c in range(input.shape[-1])]
            return np.stack(chans, axis=-1).astype(dtype)
        return zoom_linear(input, scale).astype(dtype)
    onehot = [zoom_linear((input == c).astype(dtype), scale) for c in range(num_classes)]   # to_one_hot(...).astype(dtype)
    return np.argmax(np.array(onehot), axis=0).astype(dtype)


def resize(input: np.ndarray, shape: Sequence[int], is_label: bool = False) -> np.ndarray:
    """transform.py:81-100."""
    orig = input.shape
    multi_class = len(shape) == len(orig) - 1
    scale = np.array(shape) / np.array(orig[:len(shape)])
    return rescale(input, scale, is_label=is_label, multi_class=multi_class)


def get_spacing(affine: np.ndarray):
    """data.py:55-59."""
    return tuple(float(np.linalg.norm(affine[i, :3])) for i in range(3))


def decompose(A: np.ndarray):
    """transforms3d.affines.decompose (published algorithm): A = T · R · diag(Z) · S."""
    A = np.asarray(A, dtype=np.float64)
    T = A[:3, 3].copy()
    RZS = A[:3, :3]
    ZS = np.linalg.cholesky(RZS.T @ RZS).T
    Z = np.diag(ZS).copy()
    shears = ZS / Z[:, None]
    S = shears[np.triu_indices(3, 1)]
    R = RZS @ np.linalg.inv(ZS)
    if np.linalg.det(R) < 0:
        Z[0] *= -1
        ZS[0] *= -1
        R = RZS @ np.linalg.inv(ZS)
    return T, R, Z, S


def compose(T, R, Z, S) -> np.ndarray:
    """transforms3d.affines.compose."""
    Smat = np.eye(3)
    Smat[np.triu_indices(3, 1)] = S
    A = np.eye(4)
    A[:3, :3] = R @ np.diag(Z) @ Smat
    A[:3, 3] = T
    return A


def apply_scale(affine: np.ndarray, scale) -> np.ndarray:
    """data.py:62-65."""
    T, R, Z, S = decompose(affine)
    return compose(T, R, Z * np.array(scale), S)


def resample_normalize_case(case: Dict, target_spacing, normalize_stats) -> Dict:
    """data.py:223-284."""
    case = dict(case)
    if not isinstance(normalize_stats, list):
        normalize_stats = [normalize_stats]
    scale = np.array(get_spacing(case["affine"])) / np.array(target_spacing)
    image = rescale(case["image"], scale, multi_class=True)
    chans: List[np.ndarray] = []
    for c, s in enumerate(normalize_stats):
        clipped = np.clip(image[..., c], s["pct_00_5"], s["pct_99_5"])
        chans.append((clipped - s["mean"]) / (s["std"] + 1e-8))
    case["image"] = np.stack(chans, axis=-1)
    if "label" in case:
        case["label"] = rescale(case["label"], scale, is_label=True)
    case["affine"] = apply_scale(case["affine"], 1 / scale)
    return case


# --------------------------------------------------------------------------------------
# cascade: regions of the coarse prediction, per-region detail prediction, merge
# --------------------------------------------------------------------------------------
def label_components(mask: np.ndarray):
    """``scipy.ndimage.label`` with the default structure (6-connectivity in 3-D), restated as a union-find over the
    raster: components are numbered 1.. in raster order of their first voxel.  Pure numpy/Python: small volumes only."""
    m = np.asarray(mask).astype(bool)
    X, Y, Z = m.shape
    parent = np.arange(m.size, dtype=np.int64)

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    flat = m.reshape(-1)
    for v in np.flatnonzero(flat):
        x, r = divmod(int(v), Y * Z)
        y, z = divmod(r, Z)
        for ok, n in ((x > 0, v - Y * Z), (y > 0, v - Z), (z > 0, v - 1)):
            if ok and flat[n]:
                a, b = find(int(v)), find(int(n))
                if a != b:
                    parent[max(a, b)] = min(a, b)
    out = np.zeros(m.size, dtype=np.int32)
    ids = {}
    for v in np.flatnonzero(flat):
        r = find(int(v))
        if r not in ids:
            ids[r] = len(ids) + 1          # roots are the raster-first voxels, visited in raster order
        out[v] = ids[r]
    return out.reshape(m.shape), len(ids)


def remove_small_region(mask: np.ndarray, threshold) -> np.ndarray:
    """transform.py:5-11."""
    labels, _ = label_components(mask)
    areas = np.bincount(labels.ravel())
    out = np.array(mask, copy=True)
    out[(areas < threshold)[labels]] = 0
    return out


def crop_pad_to_bbox(a: np.ndarray, bbox) -> np.ndarray:
    """transform.py:422-437 (constant zero padding)."""
    shape = a.shape
    sl = tuple(slice(max(0, int(bbox[d][0])), min(int(bbox[d][1]), shape[d])) for d in range(len(shape)))
    cropped = a[sl]
    pad = [[abs(min(0, int(bbox[d][0]))), abs(min(0, shape[d] - int(bbox[d][1])))] for d in range(len(shape))]
    if any(v > 0 for p in pad for v in p):
        cropped = np.pad(cropped, pad, "constant", constant_values=0)
    return cropped.astype(a.dtype)


def apply_translate(affine, offset):
    """data.py:68-71."""
    T, R, Z, S = decompose(affine)
    return compose(T + np.array(offset), R, Z, S)


def regions_crop_case(case: Dict, threshold=0, padding=20, based_on="label") -> List[Dict]:
    """data.py:464-492."""
    based = remove_small_region(case[based_on] > 0, threshold)
    labels, n = label_components(based)
    spacing = np.array(get_spacing(case["affine"]))
    pad = np.round(padding / spacing).astype(int)
    regions = []
    for i in range(1, n + 1):
        idx = np.nonzero(labels == i)
        bbox = np.array([[idx[d].min() - pad[d], idx[d].max() + 1 + pad[d]] for d in range(3)])
        bbox_c = np.concatenate([bbox, [[0, case["image"].shape[-1]]]])
        region = {"case_id": "%s_%03d" % (case.get("case_id", "case"), i - 1),
                  "affine": apply_translate(case["affine"], bbox[:, 0] * spacing), "bbox": bbox,
                  "image": crop_pad_to_bbox(case["image"], bbox_c)}
        if "label" in case:
            region["label"] = crop_pad_to_bbox(case["label"], bbox)
        regions.append(region)
    return regions


def merge_regions(orig_shape, num_classes: int, preds) -> np.ndarray:
    """trainer.py:189-241: ``preds`` = [(bbox, probabilities (rx, ry, rz, C))]; float64 running sums, mean where covered,
    softmax + argmax (or rounding for one class)."""
    result = np.zeros(list(orig_shape) + [num_classes])
    result_n = np.zeros_like(result)
    for bbox, pred in preds:
        shape = pred.shape[:3]
        rs, os_ = [], []
        for i in range(3):
            rs.append(slice(max(0 - int(bbox[i][0]), 0), shape[i] - max(int(bbox[i][1]) - orig_shape[i], 0)))
            os_.append(slice(max(int(bbox[i][0]), 0), min(int(bbox[i][1]), orig_shape[i])))
        result[tuple(os_)] += pred[tuple(rs)]
        result_n[tuple(os_)] += 1
    mask = result_n > 0
    result[mask] = result[mask] / result_n[mask]
    if num_classes == 1:
        result = np.around(np.squeeze(result, axis=-1))
    else:
        e = np.exp(result - result.max(axis=-1, keepdims=True))
        result = np.argmax(e / e.sum(axis=-1, keepdims=True), axis=-1)
    return result.astype(np.uint8)

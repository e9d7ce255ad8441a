"""Case-level resampling and normalisation either side of the window loop, on the GPU.

Drop-in for the reference functions that ``trainer.predict_case`` chains around ``predict_per_patch``:

  rescale(input, scale, order=1, mode='reflect', cval=0, is_label=False, multi_class=False)   transform.py:32-78
  resize(input, shape, order=1, mode='reflect', cval=0, is_label=False)                        transform.py:81-100
  resample_normalize_case(case, target_spacing, normalize_stats)                               data.py:223-284
  get_spacing(affine) / apply_scale(affine, scale)                                             data.py:55-65

numpy in, numpy out, like the reference; ``*_device`` variants keep the volumes in HBM (``predict_case`` in
``trainer.py`` uses those, so a case crosses PCIe once in each direction).  The interpolation is
``scipy.ndimage.zoom(order=1, mode='reflect')`` restated as a CUDA kernel (csrc/resample.cu) and is bit-exact with
SciPy; only the reference's defaults ``order=1, mode='reflect'`` are built -- anything else raises.  There is no CPU
path: the kernels need the library and a CUDA device.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import ops


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("unet3d_b200.transform runs on CUDA (sm_100a) only; there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def zoomed_shape(shape: Sequence[int], zoom: Sequence[float]) -> tuple:
    """Output shape of ``scipy.ndimage.zoom``: ``int(round(n * z))`` per axis (Python round: half to even)."""
    return tuple(int(round(float(n) * float(z))) for n, z in zip(shape, zoom))


def _zoom3(scale) -> tuple:
    z = np.asarray(scale, dtype=np.float64).reshape(-1)
    if z.size == 1:
        z = np.repeat(z, 3)
    if z.size != 3:
        raise ValueError(f"scale must have 1 or 3 entries, got {scale!r}")
    return tuple(float(v) for v in z)


def _check_defaults(order, mode, cval):
    if order != 1 or mode != 'reflect' or cval != 0:
        raise NotImplementedError("only the reference's defaults order=1, mode='reflect', cval=0 are built")


# ------------------------------------------------------------------------------------------------
# device-resident pieces
# ------------------------------------------------------------------------------------------------
def rescale_device(x: torch.Tensor, scale, is_label: bool = False, multi_class: bool = False,
                   num_classes: Optional[int] = None, out: Optional[torch.Tensor] = None,
                   norm=None) -> torch.Tensor:
    """transform.py:32-78 on a CUDA tensor: (X, Y, Z) or, with multi_class, (X, Y, Z, C), float32 or uint8.
    out: optional destination VIEW of the zoomed shape (e.g. the interior of a padded NCDHW model input).
    norm: per channel (pct_00_5, pct_99_5, mean, std + 1e-8): clip + z-score fused into the image zoom."""
    if not x.is_cuda:
        raise RuntimeError("rescale_device needs a CUDA tensor")
    if x.dtype not in (torch.float32, torch.uint8):
        raise NotImplementedError(f"dtype {x.dtype}: float32 images and uint8 labels are built")
    zoom = _zoom3(scale)
    src = x if multi_class else x.unsqueeze(-1)
    if src.dim() != 4:
        raise ValueError(f"expected a 3-D volume (4-D with multi_class), got {tuple(x.shape)}")
    oshape = zoomed_shape(src.shape[:3], zoom)
    if is_label:
        if x.dtype != torch.uint8 or multi_class:
            raise NotImplementedError("labels are 3-D uint8 volumes")
        if num_classes is None:
            num_classes = int(x.max().item()) + 1          # np.unique(input).max() + 1, transform.py:50
    if is_label and num_classes >= 3:
        dst = torch.empty(oshape, dtype=torch.uint8, device=x.device) if out is None else out
        if tuple(dst.shape) != oshape:
            raise ValueError(f"out has shape {tuple(dst.shape)}, the zoomed volume is {oshape}")
        ops.zoom_label(x, dst)
        return dst
    if out is None:
        dst4 = torch.empty((*oshape, src.shape[3]), dtype=x.dtype, device=x.device)
        res = dst4 if multi_class else dst4[..., 0]
    else:
        dst4 = out if multi_class else out.unsqueeze(-1)
        res = out
    if tuple(dst4.shape[:3]) != oshape:
        raise ValueError(f"out has shape {tuple(dst4.shape)}, the zoomed volume is {oshape}")
    ops.zoom_linear(src, dst4, norm)
    return res


def normalize_table(normalize_stats) -> list:
    """Per channel (lo, hi, mean, std + 1e-8) as the float32 values numpy uses in data.py:266-272."""
    if not isinstance(normalize_stats, list):
        normalize_stats = [normalize_stats]
    return [(np.float32(s['pct_00_5']), np.float32(s['pct_99_5']), np.float32(s['mean']), np.float32(s['std'] + 1e-8))
            for s in normalize_stats]


# ------------------------------------------------------------------------------------------------
# the reference's numpy-level API
# ------------------------------------------------------------------------------------------------
def rescale(input, scale, order=1, mode='reflect', cval=0, is_label=False, multi_class=False):
    _check_defaults(order, mode, cval)
    dev = _device()
    a = np.ascontiguousarray(input)
    dtype = a.dtype
    if is_label:
        if a.dtype != np.uint8:
            if a.min() < 0 or a.max() > 255:
                raise NotImplementedError("labels must fit uint8")
            a = a.astype(np.uint8)
        out = rescale_device(torch.from_numpy(a).to(dev), scale, is_label=True, num_classes=int(a.max()) + 1)
    else:
        out = rescale_device(torch.from_numpy(a.astype(np.float32, copy=False)).to(dev), scale, multi_class=multi_class)
    res = out.cpu().numpy()
    ops.check_device_errors()
    return res.astype(dtype)


def resize(input, shape, order=1, mode='reflect', cval=0, is_label=False):
    orig = input.shape
    multi_class = len(shape) == len(orig) - 1
    scale = np.array(shape) / np.array(orig[:len(shape)])
    return rescale(input, scale, order=order, mode=mode, cval=cval, is_label=is_label, multi_class=multi_class)


def get_spacing(affine):
    return tuple(np.linalg.norm(affine[i, :3]) for i in range(3))


def apply_scale(affine, scale):
    """data.py:62-65: decompose (translation, rotation, zooms, shears), scale the zooms, compose.  The published
    transforms3d algorithm, restated (the package is not a dependency here)."""
    A = np.asarray(affine, dtype=np.float64)
    RZS = A[:3, :3]
    ZS = np.linalg.cholesky(RZS.T @ RZS).T
    Z = np.diag(ZS).copy()
    shear = ZS / Z[:, None]
    R = RZS @ np.linalg.inv(ZS)
    if np.linalg.det(R) < 0:
        Z[0] *= -1
        ZS[0] *= -1
        R = RZS @ np.linalg.inv(ZS)
    S = np.eye(3)
    S[np.triu_indices(3, 1)] = shear[np.triu_indices(3, 1)]
    out = np.eye(4)
    out[:3, :3] = R @ np.diag(Z * np.array(scale)) @ S
    out[:3, 3] = A[:3, 3]
    return out


def resample_normalize_case(case: Dict, target_spacing, normalize_stats) -> Dict:
    """data.py:223-284: image zoomed to the target spacing, clipped and z-scored (one fused kernel); label zoomed."""
    dev = _device()
    case = case.copy()
    scale = np.array(get_spacing(case['affine'])) / np.array(target_spacing)
    image = np.ascontiguousarray(case['image'], dtype=np.float32)
    table = normalize_table(normalize_stats)
    if len(table) != image.shape[-1]:
        raise ValueError("one normalize_stats entry per image channel")
    case['image'] = rescale_device(torch.from_numpy(image).to(dev), scale, multi_class=True, norm=table).cpu().numpy()
    if 'label' in case:
        case['label'] = rescale(case['label'], scale, is_label=True)
    case['affine'] = apply_scale(case['affine'], 1 / scale)
    ops.check_device_errors()
    return case


# ------------------------------------------------------------------------------------------------
# cascade: regions of a coarse prediction (data.py:464-492, transform.py:5-11, 422-437)
# ------------------------------------------------------------------------------------------------
def apply_translate(affine, offset):
    """data.py:68-71 (translation + offset; rotation, zooms and shears unchanged)."""
    out = np.array(affine, dtype=np.float64, copy=True)
    out[:3, 3] = out[:3, 3] + np.array(offset)
    return out


def crop_pad_to_bbox(input, bbox, pad_mode='constant', pad_cval=0):
    """transform.py:422-437 for numpy arrays and CUDA tensors alike: crop to the box, zero-fill what lies outside."""
    if pad_mode != 'constant':
        raise NotImplementedError("only constant padding is built")
    shape = tuple(input.shape)
    src = tuple(slice(max(0, int(b[0])), min(int(b[1]), shape[d])) for d, b in enumerate(bbox))
    dst = tuple(slice(max(0, -int(b[0])), max(0, -int(b[0])) + (s.stop - s.start)) for b, s in zip(bbox, src))
    size = [int(b[1]) - int(b[0]) for b in bbox]
    if isinstance(input, torch.Tensor):
        out = torch.full(size, pad_cval, dtype=input.dtype, device=input.device)
    else:
        out = np.full(size, pad_cval, dtype=input.dtype)
    if all(s.stop > s.start for s in src):
        out[dst] = input[src]
    return out


def remove_small_region(input, threshold):
    """transform.py:5-11: zero every 6-connected component of the non-zero voxels that has fewer than ``threshold`` voxels
    (ndi.label + np.bincount + mask there; one ccl_label + ccl_stats pass here).  In place like the reference, for a
    numpy array or a CUDA tensor (X, Y, Z); returns its argument."""
    is_np = not isinstance(input, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(input)).to(_device()) if is_np else input
    if t.dim() != 3:
        raise ValueError("remove_small_region takes a (X, Y, Z) label volume")
    work = t if t.is_contiguous() else t.contiguous()
    labels, roots, stats = ops.connected_components((work != 0).to(torch.uint8))
    if roots.numel():
        flat = labels.view(-1)
        comp = torch.searchsorted(roots, flat.clamp_min(0)).clamp_max(roots.numel() - 1)    # roots are sorted flat indices
        small = (flat >= 0) & (stats[:, 0][comp] < threshold)
        work.view(-1)[small] = 0
    ops.check_device_errors()
    if is_np:
        input[...] = work.cpu().numpy()
        return input
    if work is not t:
        t.copy_(work)
    return t


def component_regions(mask: torch.Tensor, threshold=0):
    """Connected components (6-connectivity) of a boolean / uint8 CUDA volume with at least ``threshold`` voxels, in
    scipy.ndimage.label's order: [(voxels, ((x0, x1), (y0, y1), (z0, z1)))] with half-open boxes (find_objects).
    = remove_small_region + ndi.label + ndi.find_objects (transform.py:5-11, data.py:470-472) in one labelling."""
    m = (mask > 0).to(torch.uint8).contiguous()
    _, _, stats = ops.connected_components(m)
    rows = stats.cpu().numpy()
    ops.check_device_errors()
    return [(int(r[0]), ((int(r[1]), int(r[2]) + 1), (int(r[3]), int(r[4]) + 1), (int(r[5]), int(r[6]) + 1)))
            for r in rows if not r[0] < threshold]


def regions_crop_case(case: Dict, threshold=0, padding=20, based_on='label'):
    """data.py:464-492.  ``case[based_on]`` and ``case['image']`` may be numpy arrays or CUDA tensors; the crops keep the
    type of what they are cut from."""
    dev = _device()
    based = case[based_on]
    if not isinstance(based, torch.Tensor):
        based = torch.from_numpy(np.ascontiguousarray(based)).to(dev)
    comps = component_regions(based, threshold)
    spacing = np.array(get_spacing(case['affine']))
    pad = np.round(padding / spacing).astype(int)
    regions = []
    for i, (_, box) in enumerate(comps):
        bbox = np.array([[box[d][0] - pad[d], box[d][1] + pad[d]] for d in range(3)])
        bbox_c = np.concatenate([bbox, [[0, case['image'].shape[-1]]]])
        region = {'case_id': '%s_%03d' % (case.get('case_id', 'case'), i),
                  'affine': apply_translate(case['affine'], bbox[:, 0] * spacing),
                  'bbox': bbox,
                  'image': crop_pad_to_bbox(case['image'], bbox_c)}
        if 'label' in case:
            region['label'] = crop_pad_to_bbox(case['label'], bbox)
        regions.append(region)
    return regions

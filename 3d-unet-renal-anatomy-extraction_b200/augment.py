"""Train-loader augmentation on the GPU: drop-in counterparts of the reference's transform classes

    RandomRescaleCrop(scale, crop_size=128, crop_mode='center', crop_margin=0, enforce_label_indices=[], ...)  transform.py:573-652
    Crop / RandomCrop / CenterCrop(crop_size, ...)                                                              transform.py:440-571
    RandomMirror(p_per_axis)                                                                                    transform.py:279-301
    RandomContrast(factor_range) / RandomBrightness(factor_range) / RandomGamma(gamma_range)                    transform.py:176-259
    ToTensor()                                                                                                  transform.py:156-163

with the same constructor arguments, the same ``transform(case) -> case`` protocol on ``{'image': (X, Y, Z, C) float32,
'label': (X, Y, Z) uint8}`` dictionaries, and -- that is the point -- the SAME random decisions: every draw comes from
numpy's global RNG in the reference's order (``np.random.uniform`` / ``np.random.randint``), so that under
``np.random.seed(s)`` a pipeline of these classes returns what the reference's pipeline returns, bit for bit (gamma:
to one float32 ulp -- numpy's own float32 power is a CPU-dependent SIMD routine, see csrc/augment.cu).

At B200 step times (19 ms for two 128^3 patches) the reference's loader -- two worker processes running scipy's zoom on
the CPU, trainer.py:422 -- cannot feed one GPU.  Here the cheap, data-dependent part stays on the host (bounding box,
``np.unique`` of the cropped label, the crop itself: views and one small copy), the cropped block is uploaded once, and
everything that touches every voxel runs on the device: the resize of image and label (csrc/resample.cu: zoom_linear /
zoom_label, bit-exact with scipy.ndimage.zoom), mirror, contrast, brightness, gamma (csrc/augment.cu).  Cases that are
already CUDA tensors are cropped on the device.  The result stays in HBM: ``case['image']`` / ``case['label']`` are CUDA
tensors ready for ``DevicePrefetcher`` / the training step (labels stay uint8; ``Trainer`` casts them).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from . import ops
from . import transform as T


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("unet3d_b200.augment runs on CUDA (sm_100a) only; there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _to_device(a, device, dtype) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype={torch.float32: np.float32, torch.uint8: np.uint8}[dtype])).to(device)


# ------------------------------------------------------------------------------------------------
# numpy's float32 pairwise summation: leaf boundaries for n elements (cached per n and device)
# ------------------------------------------------------------------------------------------------
_LEAVES: Dict[tuple, tuple] = {}


def pairwise_leaves(n: int) -> np.ndarray:
    """Boundaries [0, ..., n] of the <= 128-element blocks numpy's pairwise sum of n contiguous floats ends up adding with
    its 8-accumulator loop: sum(n) = sum(n2) + sum(n - n2) with n2 = (n // 2) rounded down to a multiple of 8, until n <= 128
    (numpy/_core/src/umath/loops_utils.h.src: pairwise_sum).  Iterative (no Python recursion): a list of sizes is split
    level by level."""
    sizes = np.array([n], dtype=np.int64)
    while (sizes > 128).any():
        big = sizes > 128
        n2 = sizes // 2
        n2 -= n2 % 8
        out = np.empty(sizes.size + int(big.sum()), dtype=np.int64)
        pos = np.arange(sizes.size) + np.concatenate([[0], np.cumsum(big)[:-1]])
        out[pos] = np.where(big, n2, sizes)
        out[pos[big] + 1] = (sizes - n2)[big]
        sizes = out
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def _leaf_table(n: int, device):
    key = (int(n), str(device))
    if key not in _LEAVES:
        off = pairwise_leaves(int(n))
        _LEAVES[key] = (torch.from_numpy(off).to(device), torch.empty(off.size - 1, dtype=torch.float32, device=device))
    return _LEAVES[key]


def _stats(x: torch.Tensor, want_mean: bool) -> torch.Tensor:
    """float32[4] on the device: {min, max} (encoded), mean, sum of the contiguous float32 tensor x -- numpy's values."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    stats = torch.empty(4, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        if want_mean:
            off, scratch = _leaf_table(x.numel(), x.device)
            _lib.check(_lib.lib().unet3d_aug_stats(x.data_ptr(), x.numel(), off.data_ptr(), off.numel() - 1, scratch.data_ptr(),
                                                   stats.data_ptr(), ops._stream()), "unet3d_aug_stats")
        else:
            _lib.check(_lib.lib().unet3d_aug_stats(x.data_ptr(), x.numel(), None, 0, None, stats.data_ptr(), ops._stream()),
                       "unet3d_aug_stats")
    ops._count(4 if want_mean else 2)
    return stats


def mean_f32(x: torch.Tensor) -> torch.Tensor:
    """``x.mean()`` exactly as numpy computes it for a contiguous float32 array (0-dim device tensor)."""
    return _stats(x, True)[2]


# ------------------------------------------------------------------------------------------------
# device functions (transform.py:176-193, 279-301)
# ------------------------------------------------------------------------------------------------
def flip(x: torch.Tensor, axes) -> torch.Tensor:
    """np.flip over the given spatial axes (0..2) of a (X, Y, Z[, C]) float32 / uint8 CUDA tensor; returns a new tensor."""
    f = [int(a in axes) for a in range(3)]
    if not any(f):
        return x
    x = x.contiguous()
    X, Y, Z = (int(v) for v in x.shape[:3])
    C = int(x.shape[3]) if x.dim() == 4 else 1
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().unet3d_aug_flip(x.data_ptr(), out.data_ptr(), x.element_size(), X, Y, Z, C, f[0], f[1], f[2],
                                              ops._stream()), "unet3d_aug_flip")
    ops._count()
    return out


def adjust_contrast(x: torch.Tensor, factor: float) -> torch.Tensor:
    """transform.py:176-179: (input - input.mean()) * factor + input.mean(), float32 like numpy."""
    x = x.contiguous()
    st = _stats(x, True)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().unet3d_aug_affine(x.data_ptr(), out.data_ptr(), x.numel(), st.data_ptr(), 0,
                                                float(np.float32(factor)), ops._stream()), "unet3d_aug_affine")
    ops._count()
    return out


def adjust_brightness(x: torch.Tensor, factor: float) -> torch.Tensor:
    """transform.py:182-185: (input - input.min()) * factor + input.min()."""
    x = x.contiguous()
    st = _stats(x, False)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().unet3d_aug_affine(x.data_ptr(), out.data_ptr(), x.numel(), st.data_ptr(), 1,
                                                float(np.float32(factor)), ops._stream()), "unet3d_aug_affine")
    ops._count()
    return out


def adjust_gamma(x: torch.Tensor, gamma: float, epsilon: float = 1e-7) -> torch.Tensor:
    """transform.py:188-193: power((input - min) / (max - min + eps), gamma) * (max - min + eps) + min."""
    x = x.contiguous()
    st = _stats(x, False)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().unet3d_aug_gamma(x.data_ptr(), out.data_ptr(), x.numel(), st.data_ptr(), float(np.float32(gamma)),
                                               float(np.float32(epsilon)), ops._stream()), "unet3d_aug_gamma")
    ops._count()
    return out


# ------------------------------------------------------------------------------------------------
# host-side index math (transform.py:403-437), shared with the crop classes
# ------------------------------------------------------------------------------------------------
def gen_bbox_for_crop(crop_size, orig_shape, crop_margin, crop_mode):
    """transform.py:403-420, including its use of np.random.randint (one draw per axis that has room)."""
    assert crop_mode == "center" or crop_mode == "random", "crop mode must be either center or random"
    bbox = []
    for i in range(len(orig_shape)):
        if i < len(crop_size):
            if crop_mode == 'random' and orig_shape[i] - crop_size[i] - crop_margin[i] > crop_margin[i]:
                lower = np.random.randint(crop_margin[i], orig_shape[i] - crop_size[i] - crop_margin[i])
            else:
                lower = (orig_shape[i] - crop_size[i]) // 2
            bbox.append([lower, lower + crop_size[i]])
        else:
            bbox.append([0, orig_shape[i]])
    return bbox


def _label_values(cropped_label) -> np.ndarray:
    if isinstance(cropped_label, torch.Tensor):
        return torch.unique(cropped_label).cpu().numpy()
    return np.unique(cropped_label)


def _range(v, what):
    if isinstance(v, float):
        assert 0 <= v <= 1, "If range is a single number, it must be non negative"
        return [1 - v, 1 + v]
    return v


# ------------------------------------------------------------------------------------------------
# transform classes
# ------------------------------------------------------------------------------------------------
class Crop(object):
    """transform.py:440-511.  Host arrays are cropped on the host and uploaded; CUDA tensors are cropped on the device."""

    def __init__(self, crop_size=128, crop_mode='center', crop_margin=0, enforce_label_indices=[], image_pad_mode='constant',
                 image_pad_cval=0, label_pad_mode='constant', label_pad_cval=0, device=None):
        self.crop_size, self.crop_mode, self.crop_margin = crop_size, crop_mode, crop_margin
        self.enforce_label_indices = [enforce_label_indices] if isinstance(enforce_label_indices, int) else enforce_label_indices
        self.image_pad_mode, self.image_pad_cval = image_pad_mode, image_pad_cval
        self.label_pad_mode, self.label_pad_cval = label_pad_mode, label_pad_cval
        self.device = device

    def _expand(self, dim):
        if not isinstance(self.crop_size, (np.ndarray, tuple, list)):
            self.crop_size = [self.crop_size] * dim
        if not isinstance(self.crop_margin, (np.ndarray, tuple, list)):
            self.crop_margin = [self.crop_margin] * dim

    def _crop(self, image, label, size):
        """The reference's retry loop: draw boxes until the cropped label holds every enforced index."""
        while True:
            bbox = gen_bbox_for_crop(size, image.shape, self.crop_margin, self.crop_mode)
            cropped_label = T.crop_pad_to_bbox(label, bbox[:-1], self.label_pad_mode, self.label_pad_cval)
            present = _label_values(cropped_label) if len(self.enforce_label_indices) else ()
            if all(i in present for i in self.enforce_label_indices):
                break
        cropped_image = T.crop_pad_to_bbox(image, bbox, self.image_pad_mode, self.image_pad_cval)
        return cropped_image, cropped_label

    def __call__(self, case):
        image, label = case['image'], case['label']
        self._expand(len(image.shape) - 1)
        dev = _dev(self.device)
        ci, cl = self._crop(image, label, list(self.crop_size))
        case['image'] = _to_device(ci, dev, torch.float32)
        case['label'] = _to_device(cl, dev, torch.uint8)
        return case


class RandomCrop(Crop):
    """transform.py:514-544."""

    def __init__(self, crop_size=128, crop_margin=0, enforce_label_indices=[], image_pad_mode='constant', image_pad_cval=0,
                 label_pad_mode='constant', label_pad_cval=0, device=None):
        super().__init__(crop_size, crop_mode='random', crop_margin=crop_margin, enforce_label_indices=enforce_label_indices,
                         image_pad_mode=image_pad_mode, image_pad_cval=image_pad_cval, label_pad_mode=label_pad_mode,
                         label_pad_cval=label_pad_cval, device=device)


class CenterCrop(Crop):
    """transform.py:547-570."""

    def __init__(self, crop_size=128, image_pad_mode='constant', image_pad_cval=0, label_pad_mode='constant',
                 label_pad_cval=0, device=None):
        super().__init__(crop_size, crop_mode='center', image_pad_mode=image_pad_mode, image_pad_cval=image_pad_cval,
                         label_pad_mode=label_pad_mode, label_pad_cval=label_pad_cval, device=device)


class RandomRescaleCrop(Crop):
    """transform.py:573-652: draw a scale, crop round(crop_size / scale) voxels, resize the crop to crop_size (image:
    scipy zoom order 1; label: per-class one-hot zoom + argmax for >= 3 classes).  The zoom runs on the device."""

    def __init__(self, scale, crop_size=128, crop_mode='center', crop_margin=0, enforce_label_indices=[],
                 image_pad_mode='constant', image_pad_cval=0, label_pad_mode='constant', label_pad_cval=0, device=None):
        super().__init__(crop_size, crop_mode=crop_mode, crop_margin=crop_margin, enforce_label_indices=enforce_label_indices,
                         image_pad_mode=image_pad_mode, image_pad_cval=image_pad_cval, label_pad_mode=label_pad_mode,
                         label_pad_cval=label_pad_cval, device=device)
        self.scale = _range(scale, "scale")

    def __call__(self, case):
        image, label = case['image'], case['label']
        self._expand(len(image.shape) - 1)
        dev = _dev(self.device)
        scale = np.random.uniform(self.scale[0], self.scale[1])
        before = np.round(np.array(self.crop_size) / scale).astype(int)
        ci, cl = self._crop(image, label, before)
        ci, cl = _to_device(ci, dev, torch.float32), _to_device(cl, dev, torch.uint8)
        zoom = np.array(self.crop_size) / np.array(ci.shape[:3])                   # resize(): shape / orig_shape
        case['image'] = T.rescale_device(ci, zoom, multi_class=True)
        case['label'] = T.rescale_device(cl, zoom, is_label=True)                  # num_classes = max + 1 of the crop
        return case


class RandomMirror(object):
    """transform.py:279-301: one np.random.uniform() per axis; the selected axes are flipped in ONE device pass."""

    def __init__(self, p_per_axis):
        self.p_per_axis = p_per_axis

    def __call__(self, case):
        dim = len(case['image'].shape) - 1
        if not isinstance(self.p_per_axis, (np.ndarray, tuple, list)):
            self.p_per_axis = [self.p_per_axis] * dim
        axes = [i for i, p in enumerate(self.p_per_axis) if np.random.uniform() < p]
        dev = case['image'].device if isinstance(case['image'], torch.Tensor) and case['image'].is_cuda else _dev()
        img, lab = _to_device(case['image'], dev, torch.float32), _to_device(case['label'], dev, torch.uint8)
        case['image'], case['label'] = flip(img, axes), flip(lab, axes)
        return case


class _Intensity(object):
    fn = None

    def __init__(self, value_range):
        self.range = _range(value_range, "range")

    def __call__(self, case):
        v = np.random.uniform(self.range[0], self.range[1])
        img = case['image']
        dev = img.device if isinstance(img, torch.Tensor) and img.is_cuda else _dev()
        case['image'] = type(self).fn(_to_device(img, dev, torch.float32), v)
        return case


class RandomContrast(_Intensity):
    """transform.py:196-215."""
    fn = staticmethod(adjust_contrast)

    @property
    def factor_range(self):
        return self.range


class RandomBrightness(_Intensity):
    """transform.py:218-237."""
    fn = staticmethod(adjust_brightness)

    @property
    def factor_range(self):
        return self.range


class RandomGamma(_Intensity):
    """transform.py:240-259."""
    fn = staticmethod(adjust_gamma)

    @property
    def gamma_range(self):
        return self.range


def combination_lut(combinations, num_classes: int) -> np.ndarray:
    """Old label -> new label table of transform.combination_labels (transform.py:323-363): the listed combinations keep
    the order in which their first member appears when walking 0 .. num_classes - 1, every unlisted class becomes a
    combination of its own at its place in that walk; the new label is the index of the (first) combination that holds
    the old one (the reference takes the argmax over the OR-ed one-hot planes)."""
    combos = [list(c) for c in combinations] if len(np.array(combinations, dtype=object).shape) != 1 or \
        isinstance(combinations[0], (list, tuple, np.ndarray)) else [list(combinations)]
    full, used = [], []
    for c in range(num_classes):
        rows = [i for i, combo in enumerate(combos) if c in combo]
        if rows:
            for i in rows:
                if i not in used:
                    full.append(combos[i])
                    used.append(i)
        else:
            full.append([c])
    lut = np.zeros(num_classes, dtype=np.uint8)
    for c in range(num_classes):
        lut[c] = next(i for i, combo in enumerate(full) if c in combo)
    return lut


class CombineLabels(object):
    """transform.py:366-384: merge label indices (a lookup table on numpy arrays and CUDA tensors alike)."""

    def __init__(self, combinations, num_classes):
        self.combinations, self.num_classes = combinations, num_classes
        self.lut = combination_lut(combinations, num_classes)

    def __call__(self, case):
        lab = case['label']
        if isinstance(lab, torch.Tensor):
            case['label'] = torch.from_numpy(self.lut).to(lab.device)[lab.long()].to(lab.dtype)
        else:
            case['label'] = self.lut[lab].astype(lab.dtype)
        return case


class ToOnehot(object):
    """transform.py:304-320: label (d1, .., dn) -> one-hot (d1, .., dn, class), or (class, d1, .., dn) with to_tensor."""

    def __init__(self, num_classes, to_tensor=False):
        self.num_classes, self.to_tensor = num_classes, to_tensor

    def __call__(self, case):
        lab = case['label']
        if isinstance(lab, torch.Tensor):
            oh = (lab.long().unsqueeze(-1) == torch.arange(self.num_classes, device=lab.device)).to(lab.dtype)
            case['label'] = oh.permute(lab.dim(), *range(lab.dim())).contiguous() if self.to_tensor else oh
        else:
            oh = (lab[..., None] == np.arange(self.num_classes)).astype(lab.dtype)
            case['label'] = np.ascontiguousarray(np.moveaxis(oh, -1, 0)) if self.to_tensor else oh
        return case


def to_tensor(input):
    """transform.py:144-147: (d1, .., dn, class) -> (class, d1, .., dn) as a view (numpy or torch)."""
    n = input.dim() if isinstance(input, torch.Tensor) else input.ndim
    order = (n - 1, *range(n - 1))
    return input.permute(*order) if isinstance(input, torch.Tensor) else input.transpose(order)


def to_numpy(input):
    """transform.py:150-153: (class, d1, .., dn) -> (d1, .., dn, class) as a view (numpy or torch)."""
    n = input.dim() if isinstance(input, torch.Tensor) else input.ndim
    order = (*range(1, n), 0)
    return input.permute(*order) if isinstance(input, torch.Tensor) else input.transpose(order)


def to_one_hot(input, num_classes, to_tensor=False):
    """transform.py:262-276: label (d1, .., dn) -> one-hot (d1, .., dn, class) or (class, d1, .., dn), in the label's dtype."""
    return ToOnehot(num_classes, to_tensor)({'label': input})['label']


class ToNumpy(object):
    """transform.py:166-173: (C, X, Y, Z) -> (X, Y, Z, C)."""

    def __call__(self, case):
        img = to_numpy(case['image'])
        case['image'] = img.contiguous() if isinstance(img, torch.Tensor) else np.ascontiguousarray(img)
        return case


class RemoveSmallRegion(object):
    """transform.py:14-20."""

    def __init__(self, threshold):
        self.threshold = threshold

    def __call__(self, case):
        case['label'] = T.remove_small_region(case['label'], self.threshold)
        return case


class Resize(object):
    """transform.py:103-118: image (X, Y, Z, C) and label (X, Y, Z) resized to ``shape`` (linear; labels through the
    per-class zoom + argmax of transform.py:66-74)."""

    def __init__(self, shape):
        self.shape = shape

    def __call__(self, case):
        case['image'] = T.resize(case['image'], self.shape)
        case['label'] = T.resize(case['label'], self.shape, is_label=True)
        return case


class RandomRescale(object):
    """transform.py:121-141: one np.random.uniform draw, the same isotropic factor for image and label.  The reference
    passes the scalar factor to ndi.zoom for the 4-D image too, i.e. it also "zooms" the channel axis; for the
    single-channel images of every reference script that axis stays 1 and its values are unchanged, which is what is
    built here (spatial axes only); more channels are refused rather than silently resampled differently."""

    def __init__(self, scale):
        if isinstance(scale, float):
            assert 0 <= scale <= 1, "If range is a single number, it must be non negative"
            self.scale = [1 - scale, 1 + scale]
        else:
            self.scale = scale

    def __call__(self, case):
        scale = np.random.uniform(self.scale[0], self.scale[1])
        if case['image'].shape[-1] != 1:
            raise NotImplementedError("RandomRescale: single-channel images only (see the docstring)")
        case['image'] = T.rescale(case['image'], (scale, scale, scale), multi_class=True)
        case['label'] = T.rescale(case['label'], scale, is_label=True)
        return case


class ToTensor(object):
    """transform.py:156-163: (X, Y, Z, C) -> (C, X, Y, Z) (a contiguous device tensor here)."""

    def __call__(self, case):
        img = case['image']
        if isinstance(img, torch.Tensor):
            case['image'] = img.permute(3, 0, 1, 2).contiguous()
        else:
            case['image'] = np.ascontiguousarray(np.moveaxis(img, -1, 0))
        return case


class Compose(object):
    """torchvision.transforms.Compose as the reference's scripts use it (nb_train_iib.py:27-36)."""

    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __call__(self, case):
        for t in self.transforms:
            case = t(case)
        return case

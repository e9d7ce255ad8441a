"""CPU oracle for the train-loader augmentation.  TEST INFRASTRUCTURE ONLY (see unet3d_oracle.py).

A numpy restatement of what the reference's transform pipeline computes -- RandomRescaleCrop (transform.py:573-652, on
top of gen_bbox_for_crop :403-420, crop_pad_to_bbox :423-437 and resize :77-100), RandomMirror (:279-301) and the three
intensity transforms (:176-259) -- as ONE function that consumes numpy's global RNG in the reference's order.  Pinned by
tests/golden/augment.npz (tests/golden/make_golden_augment.py runs the live reference classes under fixed seeds).
The zoom is oracle/resample_oracle.py's restatement of scipy.ndimage.zoom (already pinned bit for bit).
"""
from __future__ import annotations

import numpy as np

from . import resample_oracle as R
from . import unet3d_oracle as O


def _interval(v):
    """transform.py:594-598 / 205-209: a float f means [1 - f, 1 + f]."""
    return [1 - v, 1 + v] if isinstance(v, float) else list(v)


def random_box(size, shape, margin, mode):
    """transform.py:403-420: per axis a uniform integer start when the axis has room beyond twice the margin (random
    mode), else the centred start floor((n - size) / 2); trailing axes are kept whole."""
    box = []
    for d, n in enumerate(shape):
        if d >= len(size):
            box.append([0, n])
            continue
        room = n - size[d] - margin[d]
        lo = np.random.randint(margin[d], room) if (mode == 'random' and room > margin[d]) else (n - size[d]) // 2
        box.append([lo, lo + size[d]])
    return box


def rescale_crop(image, label, scale_range, crop_size, mode='center', margin=(0, 0, 0), enforce=()):
    """transform.py:612-650: scale ~ U(range); crop round(crop_size / scale) voxels (boxes are redrawn until the label
    crop holds every enforced index), then resize image (order-1 zoom per channel) and label (one-hot zoom + argmax when
    the crop has >= 3 classes) to crop_size."""
    lo, hi = _interval(scale_range)
    scale = np.random.uniform(lo, hi)
    before = np.round(np.array(crop_size) / scale).astype(int)
    while True:
        box = random_box(before, image.shape, margin, mode)
        lab = O.crop_pad_to_bbox(label, box[:-1])
        if all(i in np.unique(lab) for i in enforce):
            break
    img = O.crop_pad_to_bbox(image, box)
    return R.resize(img, crop_size), R.resize(lab, crop_size, is_label=True)


def mirror(image, label, p_per_axis):
    """transform.py:289-301: one uniform draw per spatial axis; flip image and label where it falls below p."""
    for axis, p in enumerate(p_per_axis):
        if np.random.uniform() < p:
            image, label = np.flip(image, axis).copy(), np.flip(label, axis).copy()
    return image, label


def contrast(image, factor_range):
    """transform.py:176-179, 211-215: stretch about the (float32, numpy pairwise) mean."""
    f = np.random.uniform(*_interval(factor_range))
    m = image.mean()
    return ((image - m) * f + m).astype(image.dtype)


def brightness(image, factor_range):
    """transform.py:182-185, 233-237: stretch about the minimum."""
    f = np.random.uniform(*_interval(factor_range))
    m = image.min()
    return ((image - m) * f + m).astype(image.dtype)


def gamma(image, gamma_range, epsilon=1e-7):
    """transform.py:188-193, 255-259."""
    g = np.random.uniform(*_interval(gamma_range))
    lo, hi = image.min(), image.max()
    span = hi - lo + epsilon
    return (np.power((image - lo) / span, g) * span + lo).astype(image.dtype)


def train_pipeline(image, label, crop_size, scale=0.1, mode='random', p_mirror=(0.5, 0.5, 0.5), c=0.1, b=0.1, g=0.1,
                   enforce=(), with_gamma=True):
    """nb_train_iib.py:27-36 without the final ToTensor: returns (image (X, Y, Z, C) float32, label (X, Y, Z))."""
    image, label = rescale_crop(image, label, scale, list(crop_size), mode, [0] * len(crop_size), enforce)
    image, label = mirror(image, label, p_mirror)
    image = contrast(image, c)
    image = brightness(image, b)
    if with_gamma:
        image = gamma(image, g)
    return image, label

"""Augmentation oracle (oracle/augment_oracle.py) against the live reference's golden vectors
(tests/golden/make_golden_augment.py): bit for bit, gamma included (both are numpy on the CPU).  CPU only."""
import os

import numpy as np

from oracle import augment_oracle as A
import unet3d_b200


def _z(golden_dir):
    return np.load(os.path.join(golden_dir, "augment.npz"))


def test_oracle_pipeline_matches_reference(golden_dir):
    z = _z(golden_dir)
    crop = tuple(int(v) for v in z["crop"])
    for s in z["seeds"].tolist():
        np.random.seed(s)
        img, lab = A.train_pipeline(z["image"].copy(), z["label"].copy(), crop)
        assert np.array_equal(lab, z[f"s{s}/label"])
        assert np.array_equal(img.view(np.uint32), z[f"s{s}/image"].view(np.uint32)), s
        np.random.seed(s)
        img, _ = A.train_pipeline(z["image"].copy(), z["label"].copy(), crop, with_gamma=False)
        assert np.array_equal(img.view(np.uint32), z[f"s{s}/pre_gamma"].view(np.uint32)), s
        np.random.seed(s)
        ci, cl = A.rescale_crop(z["image"].copy(), z["label"].copy(), 0.1, list(crop), 'random')
        assert np.array_equal(ci.view(np.uint32), z[f"s{s}/crop_image"].view(np.uint32)) and np.array_equal(cl, z[f"s{s}/crop_label"])


def test_oracle_enforced_label_crop(golden_dir):
    z = _z(golden_dir)
    for s in z["seeds"].tolist()[:3]:
        np.random.seed(s)
        img, lab = A.rescale_crop(z["image"].copy(), z["label"].copy(), 0.2, [12, 12, 12], 'random', enforce=(2,))
        img, lab = A.mirror(img, lab, [0.5] * 3)
        assert 2 in lab
        assert np.array_equal(lab, z[f"e{s}/label"]) and np.array_equal(img.view(np.uint32), z[f"e{s}/image"].view(np.uint32))


def test_pairwise_leaf_table_reproduces_numpy_sum():
    """The leaf boundaries the CUDA mean kernel walks (augment.pairwise_leaves) + numpy's 8-accumulator leaf rule + the
    uneven binary tree reproduce np.sum / np.mean of float32 arrays bit for bit."""
    f32 = np.float32
    rng = np.random.RandomState(0)

    def leaf(a):
        if a.size < 8:
            r = f32(0.)
            for v in a:
                r = f32(r + v)
            return r
        r = a[:8].copy()
        m = a.size - a.size % 8
        for row in a[8:m].reshape(-1, 8):
            r = (r + row).astype(f32)
        res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
        for v in a[m:]:
            res = f32(res + v)
        return res

    def tree(sums, n):           # the recursion of csrc/augment.cu: aug_tree_kernel
        it = iter(sums)

        def go(m):
            if m <= 128:
                return next(it)
            m2 = m // 2
            m2 -= m2 % 8
            left = go(m2)
            return f32(left + go(m - m2))
        return go(n)

    for n in (5, 8, 127, 128, 129, 1000, 4097, 24 ** 3, 65539):
        a = (rng.randn(n) * 3 + 0.5).astype(f32)
        off = unet3d_b200.augment.pairwise_leaves(n)
        assert off[0] == 0 and off[-1] == n and (np.diff(off) <= 128).all() and (np.diff(off) > 0).all()
        total = tree([leaf(a[off[i]:off[i + 1]]) for i in range(off.size - 1)], n)
        assert total == a.sum()
        assert f32(total / f32(n)) == a.mean()


def test_label_recoding_matches_reference(golden_dir):
    """CombineLabels / ToOnehot (transform.py:304-384) on numpy labels against the live reference's outputs."""
    z = _z(golden_dir)
    G = unet3d_b200.augment
    lab = z["label"]
    assert np.array_equal(G.CombineLabels([[1, 2]], 3)({"label": lab.copy()})["label"], z["combine_12"])
    assert np.array_equal(G.CombineLabels([0, 1], 3)({"label": lab.copy()})["label"], z["combine_01"])
    small = lab[:6, :5, :4].copy()
    a = G.ToOnehot(3)({"label": small.copy()})["label"]
    b = G.ToOnehot(3, to_tensor=True)({"label": small.copy()})["label"]
    assert a.dtype == z["onehot"].dtype and np.array_equal(a, z["onehot"]) and np.array_equal(b, z["onehot_t"])


def test_layout_helpers_match_reference_semantics():
    """to_tensor / to_numpy / to_one_hot (transform.py:144-153, 262-276) on numpy and torch inputs."""
    import torch
    from unet3d_b200 import augment as G
    rng = np.random.RandomState(0)
    a = rng.randn(3, 4, 5, 2).astype(np.float32)
    dims = np.arange(a.ndim)
    assert np.array_equal(G.to_tensor(a), a.transpose(np.concatenate((dims[-1:], dims[:-1]))))
    assert np.array_equal(G.to_numpy(G.to_tensor(a)), a)
    assert torch.equal(G.to_numpy(G.to_tensor(torch.from_numpy(a))), torch.from_numpy(a))
    lab = rng.randint(0, 4, (3, 4, 5)).astype(np.uint8)
    want = np.eye(4)[lab]
    assert np.array_equal(G.to_one_hot(lab, 4), want.astype(np.uint8))
    assert np.array_equal(G.to_one_hot(lab, 4, to_tensor=True), np.moveaxis(want, -1, 0).astype(np.uint8))
    case = G.ToNumpy()({"image": np.ascontiguousarray(np.moveaxis(a, -1, 0))})
    assert np.array_equal(case["image"], a) and case["image"].flags["C_CONTIGUOUS"]

"""CPU oracle for the 3D U-Net hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (fp32, torch-CPU functional ops + numpy index math) of
the algorithm the reference runs on its hot path.  It is the checker, never the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
/ ``--impl reference`` legs may import it.  The product package must fail loudly if
its CUDA extension is missing; it never routes through this module.

Pinning: the reference ships **no tests, fixtures or golden vectors** (SURVEY.md §4), so
parity is *unpinned by reference tests*.  Instead the restatement is pinned against the
live reference modules imported from /root/reference in the build container
(``tests/golden/make_golden.py``) and against the committed outputs of that script
(``tests/golden/*.npz``), which travel to the GPU box.

Everything takes a plain ``state_dict`` (reference key names, PyTorch weight layouts)
so that the same weights can be fed to the reference, this oracle and the CUDA path.

Reference citations are ``file:line`` into /root/reference.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LRELU_SLOPE = 0.01   # nn.LeakyReLU default, network.py:165,390 (nonlin_kwargs has no slope)
IN_EPS = 1e-5        # nn.InstanceNorm3d default eps, affine=False (network.py:163,388)


# --------------------------------------------------------------------------------------
# network.py
# --------------------------------------------------------------------------------------
def paired_features(num_pool: int, num_features: int) -> List[List[int]]:
    """network.py:135-141 (generate_paired_features)."""
    down = [[num_features * 2 ** i] * 2 for i in range(num_pool)]
    bottom = [[num_features * 2 ** num_pool] * 2]
    up = [[num_features * 2 ** i] * 2 for i in range(num_pool - 1, -1, -1)]
    return down + bottom + up


def paired_features2(num_pool: int, num_features: int) -> List[List[int]]:
    """network.py:144-150 (generate_paired_features2): encoder pairs widen f -> 2f."""
    down = [[num_features * 2 ** i, num_features * 2 ** (i + 1)] for i in range(num_pool)]
    bottom = [[num_features * 2 ** num_pool] * 2]
    up = [[num_features * 2 ** i] * 2 for i in range(num_pool - 1, -1, -1)]
    return down + bottom + up


# network.py:18-22: down + bottom + up widths of ResAttrUnet3D2
ATTR2_FEATURES = ([[30, 30], [60, 60], [120, 120], [240, 240], [320, 320]] + [[320, 320]] +
                  [[320, 320], [240, 240], [120, 120], [60, 60], [30, 30]])


class DropoutMasks:
    """Source of Dropout3d channel masks (network.py:159-160,382-383,412-413).

    ``None`` masks (eval mode) make every dropout the identity.  In train mode the
    mask for a tensor of shape (N,C,...) is ``bernoulli(1-p)/(1-p)`` of shape
    (N,C,1,1,1) -- the same draw ``F.dropout3d`` makes.
    A recorded list can be replayed so two implementations see identical masks.
    """

    def __init__(self, train: bool, p: float = 0.5, generator: Optional[torch.Generator] = None,
                 replay: Optional[List[Tensor]] = None):
        self.train, self.p, self.generator = train, p, generator
        self.replay = list(replay) if replay is not None else None
        self.record: List[Tensor] = []
        self._i = 0

    def next(self, n: int, c: int) -> Optional[Tensor]:
        if not self.train:
            return None
        if self.replay is not None:
            m = self.replay[self._i]
            self._i += 1
        else:
            m = torch.empty(n, c, 1, 1, 1).bernoulli_(1 - self.p, generator=self.generator).div_(1 - self.p)
        self.record.append(m)
        return m


def _inorm(x: Tensor) -> Tensor:
    # InstanceNorm3d(affine=False, track_running_stats=False): biased variance per (n,c)
    return F.instance_norm(x, eps=IN_EPS)


BN_TRAIN = False        # set by resunet3d_forward(bn_train=...) for the duration of one forward


def _norm(sd: Dict[str, Tensor], pre: str, x: Tensor) -> Tensor:
    """The block's norm module: InstanceNorm3d (no state) unless the state dict holds BatchNorm3d tensors under
    `pre` (network.py:38-69 variant: affine, running statistics, momentum 0.1, eps 1e-5).  In training mode the
    running buffers IN `sd` are updated in place, once per application, exactly like nn.BatchNorm3d."""
    if pre + "weight" not in sd:
        return _inorm(x)
    if BN_TRAIN:
        sd[pre + "num_batches_tracked"] += 1
    return F.batch_norm(x, sd[pre + "running_mean"], sd[pre + "running_var"], sd[pre + "weight"], sd[pre + "bias"],
                        training=BN_TRAIN, momentum=0.1, eps=1e-5)


def _lrelu(x: Tensor) -> Tensor:
    return F.leaky_relu(x, LRELU_SLOPE)


def res_block(sd: Dict[str, Tensor], pre: str, x: Tensor, cin: int, cout: int, stride: int,
              masks: DropoutMasks) -> Tensor:
    """network.py:405-416 (ResBlock.forward)."""
    if cin != cout or stride != 1:
        skip = F.conv3d(x, sd[pre + "skip_conv.weight"], sd[pre + "skip_conv.bias"], stride=stride)
    else:
        skip = x
    y = F.conv3d(x, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], stride=stride, padding=1)
    m = masks.next(y.shape[0], y.shape[1])
    if m is not None:
        y = y * m
    y = _lrelu(_norm(sd, pre + "norm.", y))
    y = F.conv3d(y, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return _lrelu(_norm(sd, pre + "norm.", y) + skip)        # the SAME norm module twice (network.py:401-416)


def res_block_stack(sd, pre, x, cin, cout, num_stacks, masks) -> Tensor:
    """network.py:419-449 (ResBlockStack): first block cin->cout, rest cout->cout, stride 1."""
    for j in range(num_stacks):
        x = res_block(sd, f"{pre}res_blocks.{j}.", x, cin if j == 0 else cout, cout, 1, masks)
    return x


def conv_block(sd, pre, x, masks) -> Tensor:
    """network.py:153-182 (ConvBlock.forward): conv -> dropout -> IN -> LeakyReLU."""
    y = F.conv3d(x, sd[pre + "conv.weight"], sd[pre + "conv.bias"], padding=1)
    m = masks.next(y.shape[0], y.shape[1])
    if m is not None:
        y = y * m
    return _lrelu(_norm(sd, pre + "norm.", y))


def conv_block_stack(sd, pre, x, num_stacks, masks) -> Tensor:
    """network.py:185-214."""
    for j in range(num_stacks):
        x = conv_block(sd, f"{pre}conv_blocks.{j}.", x, masks)
    return x


def conv_trans3d(sd, pre, x) -> Tensor:
    """network.py:298-320: ConvTranspose3d(k3,s2,p1) -> ConstantPad3d((0,1)*3, 0) -> IN -> LeakyReLU.
    The pad plane is zero *before* the norm."""
    y = F.conv_transpose3d(x, sd[pre + "up.0.weight"], sd[pre + "up.0.bias"], stride=2, padding=1)
    y = F.pad(y, (0, 1, 0, 1, 0, 1), value=0.0)
    return _lrelu(_norm(sd, pre + "up.2.", y))


def att_block(sd, pre, x, gate) -> Tensor:
    """network.py:353-371 (AttBlock): one shared 1x1x1 conv used three times."""
    w, b = sd[pre + "conv.weight"], sd[pre + "conv.bias"]
    x = F.conv3d(x, w, b)
    g = F.conv3d(gate, w, b)
    f = _lrelu(x + g)
    rate = torch.sigmoid(F.conv3d(f, w, b))
    return x * rate


def up_concat(sd, pre, x, skip, attention: bool = False) -> Tensor:
    """network.py:323-350: cat order is [upsampled, skip]."""
    x = conv_trans3d(sd, pre + "conv_trans.", x)
    if attention:
        skip = att_block(sd, pre + "att_gate.", skip, x)
    return torch.cat((x, skip), dim=1)


def resunet3d_forward(sd: Dict[str, Tensor], x: Tensor, num_pool: int = 4, num_features: int = 30,
                      masks: Optional[DropoutMasks] = None, attention: bool = False, bn_train: bool = False,
                      pf: Optional[List[List[int]]] = None) -> Tensor:
    """network.py:104-132 (ResUnet3D) over network.py:549-565 (Unet.forward).

    A state dict with ``...norm.weight`` / ``...up.2.weight`` tensors selects the BatchNorm3d variant
    (ResAttrBNUnet3D, network.py:38-69); ``bn_train`` = the module's training flag for those norms (batch statistics,
    running buffers in ``sd`` updated in place).

    encode level L = ResBlockStack with max(L,1) blocks (network.py:116-118); pooling is a
    stride-2 ResBlock (network.py:125-126); decode = ResBlock(2f -> f) (network.py:523-527).
    ``attention=True`` gives ResAttrUnet3D (network.py:72-101); an explicit ``pf`` (list of channel pairs) gives the
    nets with hand-written widths: ResAttrUnet3D2 (network.py:6-35) = ``pf=ATTR2_FEATURES, attention=True``.
    """
    global BN_TRAIN
    BN_TRAIN = bool(bn_train)
    masks = masks or DropoutMasks(train=False)
    if pf is None:
        pf = paired_features(num_pool, num_features)
    npairs = len(pf)
    num_pool = npairs // 2
    x = F.conv3d(x, sd["net.conv.weight"], sd["net.conv.bias"], padding=1)   # network.py:550, no norm
    skips = []
    for i in range(num_pool):
        x = res_block_stack(sd, f"net.encode_blocks.{i}.", x, pf[i][0], pf[i][1], max(i, 1), masks)
        skips.append(x)
        x = res_block(sd, f"net.pool_blocks.{i}.", x, pf[i][1], pf[i + 1][0], 2, masks)
    x = res_block_stack(sd, f"net.encode_blocks.{num_pool}.", x, pf[num_pool][0], pf[num_pool][1],
                        max(num_pool, 1), masks)
    for i in range(num_pool - 1, -1, -1):
        x = up_concat(sd, f"net.up_blocks.{i}.", x, skips[i], attention)
        cin = pf[npairs - i - 1][0] + pf[i][1]
        x = res_block(sd, f"net.decode_blocks.{i}.", x, cin, pf[npairs - i - 1][1], 1, masks)
    return F.conv3d(x, sd["net.fc.weight"], sd["net.fc.bias"])               # network.py:563


def plain_unet_forward(sd: Dict[str, Tensor], x: Tensor, pf: List[List[int]],
                       num_stacks: int = 2, masks: Optional[DropoutMasks] = None) -> Tensor:
    """Generic ``Unet(...)`` with its default blocks (network.py:470-487): MaxPoolBlock k2 s2
    (network.py:452-463) and ConvBlockStack(num_stacks=2) encode/decode.  State-dict keys carry
    the extra ``net.``-less prefix of a bare ``Unet`` module."""
    masks = masks or DropoutMasks(train=False)
    npairs = len(pf)
    num_pool = npairs // 2
    x = F.conv3d(x, sd["conv.weight"], sd["conv.bias"], padding=1)
    skips = []
    for i in range(num_pool):
        x = conv_block_stack(sd, f"encode_blocks.{i}.", x, num_stacks, masks)
        skips.append(x)
        x = F.max_pool3d(x, kernel_size=2, stride=2)
    x = conv_block_stack(sd, f"encode_blocks.{num_pool}.", x, num_stacks, masks)
    for i in range(num_pool - 1, -1, -1):
        x = up_concat(sd, f"up_blocks.{i}.", x, skips[i])
        x = conv_block_stack(sd, f"decode_blocks.{i}.", x, num_stacks, masks)
    return F.conv3d(x, sd["fc.weight"], sd["fc.bias"])


# --------------------------------------------------------------------------------------
# loss.py
# --------------------------------------------------------------------------------------
def _probs(logits: Tensor) -> Tensor:
    """loss.py:7-11."""
    return torch.softmax(logits, dim=1) if logits.size(1) > 1 else torch.sigmoid(logits)


def _flatten(inp: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
    """loss.py:14-29: (N,C,...) -> (N*V, C); target -> one-hot (N*V, C)."""
    c = inp.size(1)
    flat = inp.reshape(inp.size(0), c, -1).transpose(1, 2).reshape(-1, c)
    onehot = F.one_hot(target, num_classes=c).reshape(-1, c)
    return flat, onehot


def dice(p: Tensor, g: Tensor, alpha: float = 0.5, beta: float = 0.5, smooth: float = 1e-7) -> Tensor:
    """loss.py:32-48 (Tversky form, sums over the whole batch)."""
    p = p.reshape(-1)
    g = g.reshape(-1)
    tp = (p * g).sum()
    fn = ((1 - p) * g).sum()
    fp = (p * (1 - g)).sum()
    return (tp + smooth) / (tp + alpha * fn + beta * fp + smooth)


def _weights(c: int, weight_v) -> Tensor:
    """loss.py:151-155 / 64-68: weight_c and the class mask are computed then overwritten;
    only L1-normalised weight_v survives (reference behaviour, kept bug-compatible)."""
    w = torch.ones(c) if weight_v is None else torch.tensor(weight_v)
    return F.normalize(w.type(torch.float), p=1, dim=0)


def dice_per_class(logits: Tensor, target: Tensor, alpha=0.5, beta=0.5, smooth=1e-7) -> Tensor:
    p, g = _flatten(_probs(logits), target)
    return torch.stack([dice(p[:, i], g[:, i], alpha, beta, smooth) for i in range(p.size(1))])


def dice_loss(logits, target, weight_v=None, alpha=0.5, beta=0.5, smooth=1e-7) -> Tensor:
    """loss.py:143-166 (DiceLoss.forward)."""
    d = dice_per_class(logits, target, alpha, beta, smooth)
    return (_weights(logits.size(1), weight_v) * (1 - d)).sum()


def dice_metric(logits, target, weight_v=None, alpha=0.5, beta=0.5, smooth=1e-7) -> Tensor:
    """loss.py:104-120 (Dice.forward)."""
    d = dice_per_class(logits, target, alpha, beta, smooth)
    return (_weights(logits.size(1), weight_v) * d).sum()


def focal_per_class(logits: Tensor, target: Tensor, gamma: float = 2) -> Tensor:
    """loss.py:70-80: C * mean_v( -(1-pt)^gamma * onehot * log pt ) per class."""
    x, g = _flatten(logits, target)
    c = x.size(-1)
    if c > 1:
        logpt = F.log_softmax(x, -1)
        pt = logpt.exp()
    else:
        pt = torch.sigmoid(x)
        logpt = torch.log(pt)
    return c * (-(1 - pt) ** gamma * g * logpt).mean(dim=0)


def focal_loss(logits, target, gamma=2, weight_v=None) -> Tensor:
    """loss.py:186-193 (FocalLoss.forward) -> loss.py:51-82."""
    return (_weights(logits.size(1), weight_v) * focal_per_class(logits, target, gamma)).sum()


def hybrid_loss(logits, target, gamma=2, weight_v=None, alpha=0.5, beta=0.5, smooth=1e-7) -> Tensor:
    """loss.py:218-254 (HybirdLoss.forward): sum_c w_c (1 - dice_c + focal_c)."""
    d = dice_per_class(logits, target, alpha, beta, smooth)
    f = focal_per_class(logits, target, gamma)
    return (_weights(logits.size(1), weight_v) * (1 - d + f)).sum()


# --------------------------------------------------------------------------------------
# transform.py helpers used by predict_per_patch (pure integer index math, bit-exact)
# --------------------------------------------------------------------------------------
def center_bbox(crop_size: Sequence[int], shape: Sequence[int]) -> List[List[int]]:
    """transform.py:403-420 (gen_bbox_for_crop, crop_mode='center', margin 0)."""
    bbox = []
    for i, s in enumerate(shape):
        if i < len(crop_size):
            lo = (s - crop_size[i]) // 2          # floor division, may be negative (=> pad)
            bbox.append([lo, lo + crop_size[i]])
        else:
            bbox.append([0, s])
    return bbox


def crop_pad_to_bbox(a: np.ndarray, bbox, cval=0) -> np.ndarray:
    """transform.py:423-437: crop to the in-range part, then constant-pad the rest."""
    shape = a.shape
    sl = tuple(slice(max(0, b[0]), min(b[1], shape[d])) for d, b in enumerate(bbox))
    out = a[sl]
    pw = [[abs(min(0, b[0])), abs(min(0, shape[d] - b[1]))] for d, b in enumerate(bbox)]
    if any(v > 0 for p in pw for v in p):
        out = np.pad(out, pw, "constant", constant_values=cval)
    return out.astype(a.dtype)


def crop_pad(a: np.ndarray, size: Sequence[int]) -> np.ndarray:
    """transform.py:393-400 (crop_pad, centre mode)."""
    return crop_pad_to_bbox(a, center_bbox(size, a.shape))


def pad_to(a: np.ndarray, size: Sequence[int]) -> np.ndarray:
    """transform.py:387-390 (pad): grow each leading axis to at least ``size``."""
    tgt = [max(a.shape[d], size[d]) for d in range(len(size))]
    return crop_pad(a, tgt)


# --------------------------------------------------------------------------------------
# trainer.py:17-98 predict_per_patch
# --------------------------------------------------------------------------------------
def tile_centres(extent: int, patch: int, step_per_patch: int) -> np.ndarray:
    """trainer.py:29-40 for one axis, including the float-step truncation quirk
    (SURVEY.md Q1): ``np.arange(start, stop, float_step, dtype=int)`` yields
    ``start + i * (int(start + step) - start)`` for ``ceil((stop - start) / step)`` items."""
    start = patch // 2
    end = extent - patch // 2
    num_steps = math.ceil((end - start) / (patch / step_per_patch))
    step = (end - start) / (num_steps + 1e-8)
    if step == 0:
        step = 9999999
    stop = end + 1e-8
    count = int(math.ceil((stop - start) / step))
    delta = int(start + step) - start
    return np.array([start + i * delta for i in range(count)], dtype=np.int64)


def tile_slices(shape: Sequence[int], patch: Sequence[int], step_per_patch: int):
    """trainer.py:53-65: x outermost, z fastest; slice = centre -/+ patch//2."""
    cs = [tile_centres(shape[d], patch[d], step_per_patch) for d in range(3)]
    out = []
    for x in cs[0]:
        for y in cs[1]:
            for z in cs[2]:
                out.append(tuple(slice(int(c) - patch[d] // 2, int(c) + patch[d] // 2)
                                 for d, c in enumerate((x, y, z))))
    return out


def gaussian_window(patch: Sequence[int], sigma_scale: float = 0.125) -> np.ndarray:
    """Separable Gaussian importance map (north-star extension; no reference code).
    w(i) = exp(-0.5 ((i - (p-1)/2) / (sigma_scale p))^2), product over axes, fp32."""
    ax = []
    for p in patch:
        i = np.arange(p, dtype=np.float64)
        ax.append(np.exp(-0.5 * ((i - (p - 1) / 2.0) / (sigma_scale * p)) ** 2))
    w = ax[0][:, None, None] * ax[1][None, :, None] * ax[2][None, None, :]
    return w.astype(np.float32)


def predict_per_patch(inp: np.ndarray, model_fn: Callable[[Tensor], Tensor], num_classes: int = 3,
                      patch_size=(96, 96, 96), step_per_patch: int = 4, one_hot: bool = False,
                      window: Optional[np.ndarray] = None) -> np.ndarray:
    """trainer.py:17-98.  ``inp`` is (X,Y,Z,C_in) float32; ``model_fn`` maps (1,C_in,px,py,pz)
    -> logits (1,C,px,py,pz).  ``window=None`` is the reference's uniform blending
    (``result += p; result_n += 1``); a (px,py,pz) window gives weighted blending."""
    orig = inp.shape[:3]
    inp = pad_to(inp, patch_size)
    shape = inp.shape[:3]
    result = torch.zeros([num_classes] + list(shape))
    result_n = torch.zeros_like(result)
    x = torch.from_numpy(np.ascontiguousarray(np.moveaxis(inp, -1, 0))[None])
    w = None if window is None else torch.from_numpy(window)
    with torch.no_grad():
        for sl in tile_slices(shape, patch_size, step_per_patch):
            out = model_fn(x[(slice(None), slice(None)) + sl])
            out = torch.sigmoid(out) if num_classes == 1 else torch.softmax(out, dim=1)
            if w is None:
                result[(slice(None),) + sl] += out[0]
                result_n[(slice(None),) + sl] += 1
            else:
                result[(slice(None),) + sl] += out[0] * w
                result_n[(slice(None),) + sl] += w
    result = result / result_n          # uncovered voxels: 0/0 = NaN
    if one_hot:
        res = np.moveaxis(result.numpy(), 0, -1).astype(np.float32)
    else:
        if num_classes == 1:
            result = result.squeeze(0)
        else:
            result = torch.argmax(torch.softmax(result, dim=0), dim=0)
        res = np.round(result.numpy()).astype(np.uint8)
    return crop_pad(res, orig)

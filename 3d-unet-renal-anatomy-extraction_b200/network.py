"""Drop-in replacement for the reference's ``network.py`` model classes on the hot path.

Same constructor signatures, attribute names, submodule tree, registration order (so ``parameters()`` -- and with
it optimizer ``state_dict``s -- line up) and ``state_dict`` keys as the reference's generic ``Unet``
(network.py:470-565), its blocks (``ConvBlock[Stack]`` :153-214, ``ConvTrans3D`` :298-320, ``UpConcat`` :323-350,
``AttBlock`` :353-371, ``ResBlock[Stack]`` :374-449, ``MaxPoolBlock`` :452-463) and the concrete nets built from
them (``ResUnet3D`` :104-132, ``ResAttrUnet3D`` :72-101, ``ResAttrUnet3D2`` :6-35), so reference checkpoints load
unchanged and ``trainer.py`` can drive them.  The parameters are ordinary fp32 ``nn.Parameter``s held by
never-called ``nn.Conv3d`` / ``nn.ConvTranspose3d`` containers (which gives PyTorch's default initialisation);
``forward`` does not run a single PyTorch op on the activations: it hands the whole forward / backward pass to
:mod:`engine`, which enqueues the library's sm_100a kernels.  There is no fallback -- without the CUDA library
or on a CPU tensor, ``forward`` raises.

The op/kwargs hooks of the reference's blocks (``conv_op``, ``norm_op``, ``dropout_kwargs`` ...) are accepted
with the reference's defaults; values the kernels do not implement raise ``NotImplementedError`` at
construction time instead of silently computing something else.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from . import engine


def generate_paired_features(num_pool: int, num_features: int) -> List[List[int]]:
    """Channel plan f, 2f, ... on the way down, 2^num_pool f at the bottom, mirrored on the way up
    (same result as network.py:135-141)."""
    widths = [num_features * (1 << i) for i in range(num_pool + 1)]
    return [[w, w] for w in widths] + [[w, w] for w in reversed(widths[:-1])]


def generate_paired_features2(num_pool: int, num_features: int) -> List[List[int]]:
    """The plan for pooling blocks that keep the width (max-pool): level i goes f·2^i -> f·2^(i+1)
    (same result as network.py:144-150)."""
    down = [[num_features << i, num_features << (i + 1)] for i in range(num_pool)]
    up = [[num_features << i, num_features << i] for i in range(num_pool - 1, -1, -1)]
    return down + [[num_features << num_pool, num_features << num_pool]] + up


_CONV_KW = {'kernel_size': 3, 'padding': 1}
_DROP_KW = {'p': 0.5, 'inplace': True}
_NONLIN_KW = {'inplace': True}


def _check_ops(conv_op, conv_kwargs, dropout_op, dropout_kwargs, norm_op, norm_kwargs, nonlin_op, nonlin_kwargs):
    """The kernels implement exactly the reference's default layer choices."""
    def as3(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v,) * 3
    if conv_op is not nn.Conv3d:
        raise NotImplementedError("conv_op must be nn.Conv3d")
    ck = dict(conv_kwargs)
    if as3(ck.pop('kernel_size', None)) != (3, 3, 3) or as3(ck.pop('padding', 0)) != (1, 1, 1) or ck:
        raise NotImplementedError(f"conv_kwargs {conv_kwargs}: only kernel_size=3, padding=1 is built")
    if dropout_op not in (nn.Dropout3d, None):
        raise NotImplementedError("dropout_op must be nn.Dropout3d or None")
    if norm_op not in (nn.InstanceNorm3d, nn.BatchNorm3d) or norm_kwargs:
        raise NotImplementedError("norm_op must be nn.InstanceNorm3d or nn.BatchNorm3d with default arguments")
    if nonlin_op is not nn.LeakyReLU or nonlin_kwargs.get('negative_slope', 0.01) != 0.01:
        raise NotImplementedError("nonlin_op must be nn.LeakyReLU(negative_slope=0.01)")


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("blocks are parameter containers; run the enclosing Unet")


class ConvBlock(_Container):
    """network.py:153-182: conv k3 -> Dropout3d -> InstanceNorm3d -> LeakyReLU."""

    def __init__(self, in_channels, out_channels, conv_op=nn.Conv3d, conv_kwargs=_CONV_KW, dropout_op=nn.Dropout3d,
                 dropout_kwargs=_DROP_KW, norm_op=nn.InstanceNorm3d, norm_kwargs={}, nonlin_op=nn.LeakyReLU,
                 nonlin_kwargs=_NONLIN_KW):
        super().__init__()
        _check_ops(conv_op, conv_kwargs, dropout_op, dropout_kwargs, norm_op, norm_kwargs, nonlin_op, nonlin_kwargs)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv = conv_op(in_channels, out_channels, **conv_kwargs)
        self.dropout = dropout_op(**dropout_kwargs) if dropout_op else None
        self.norm = norm_op(out_channels, **norm_kwargs)
        self.nonlin = nonlin_op(**nonlin_kwargs)

    @property
    def dropout_p(self) -> float:
        return float(self.dropout.p) if self.dropout is not None else 0.0


class ConvBlockStack(_Container):
    """network.py:185-214: num_stacks ConvBlocks, the first one changing the width."""

    def __init__(self, in_channels, out_channels, num_stacks=2, **block_kwargs):
        super().__init__()
        self.conv_blocks = nn.ModuleList(
            [ConvBlock(in_channels if i == 0 else out_channels, out_channels, **block_kwargs) for i in range(num_stacks)])


class ConvTrans3D(_Container):
    """network.py:298-320: ConvTranspose3d(k3,s2,p1) -> zero pad to 2x -> InstanceNorm -> LeakyReLU.
    Only the transposed conv has parameters; it sits at ``up.0`` like in the reference."""

    def __init__(self, in_channels, out_channels, norm_op=nn.InstanceNorm3d, norm_kwargs={}, nonlin_op=nn.LeakyReLU,
                 nonlin_kwargs=_NONLIN_KW):
        super().__init__()
        _check_ops(nn.Conv3d, _CONV_KW, None, {}, norm_op, norm_kwargs, nonlin_op, nonlin_kwargs)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.up = nn.Sequential(
            nn.ConvTranspose3d(in_channels, out_channels, kernel_size=3, stride=2, padding=1),
            nn.ConstantPad3d(padding=(0, 1, 0, 1, 0, 1), value=0),
            norm_op(out_channels, **norm_kwargs),
            nonlin_op(**nonlin_kwargs))


class AttBlock(_Container):
    """network.py:353-371: ONE 1x1x1 conv applied to the skip, the gate and their activated sum;
    result = conv(skip) * sigmoid(conv(lrelu(conv(skip) + conv(gate))))."""

    def __init__(self, out_channels, conv_op=nn.Conv3d, nonlin_op=nn.LeakyReLU, nonlin_kwargs=_NONLIN_KW):
        super().__init__()
        _check_ops(conv_op, _CONV_KW, None, {}, nn.InstanceNorm3d, {}, nonlin_op, nonlin_kwargs)
        self.conv = conv_op(out_channels, out_channels, kernel_size=1)
        self.lrelu = nonlin_op(**nonlin_kwargs)
        self.active = nn.Sigmoid()


class UpConcat(_Container):
    """network.py:323-350: upsample, optionally gate the skip, then concat [upsampled, skip]."""

    def __init__(self, in_channels, out_channels, conv_trans_op=ConvTrans3D, attention=False, att_conv_op=nn.Conv3d,
                 norm_op=nn.InstanceNorm3d, norm_kwargs={}, nonlin_op=nn.LeakyReLU, nonlin_kwargs=_NONLIN_KW):
        super().__init__()
        if conv_trans_op is not ConvTrans3D:
            raise NotImplementedError("conv_trans_op must be ConvTrans3D")
        self.in_channels, self.out_channels, self.attention = in_channels, out_channels, attention
        self.conv_trans = conv_trans_op(in_channels, out_channels, norm_op=norm_op, norm_kwargs=norm_kwargs,
                                        nonlin_op=nonlin_op, nonlin_kwargs=nonlin_kwargs)
        if attention:
            self.att_gate = AttBlock(out_channels, conv_op=att_conv_op, nonlin_op=nonlin_op, nonlin_kwargs=nonlin_kwargs)


class ResBlock(_Container):
    """network.py:374-416: conv1 (k3, stride s) -> dropout -> IN -> LReLU -> conv2 (k3) -> IN -> (+skip) -> LReLU.
    skip_conv (k1, stride s) is only used when in != out or s != 1 but always present, as in the reference,
    so state_dicts match."""

    def __init__(self, in_channels, out_channels, stride=1, conv_op=nn.Conv3d, conv_kwargs=_CONV_KW,
                 dropout_op=nn.Dropout3d, dropout_kwargs=_DROP_KW, norm_op=nn.InstanceNorm3d, norm_kwargs={},
                 nonlin_op=nn.LeakyReLU, nonlin_kwargs=_NONLIN_KW):
        super().__init__()
        _check_ops(conv_op, conv_kwargs, dropout_op, dropout_kwargs, norm_op, norm_kwargs, nonlin_op, nonlin_kwargs)
        if stride not in (1, 2):
            raise NotImplementedError("stride must be 1 or 2")
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        self.conv1 = conv_op(in_channels, out_channels, stride=stride, **conv_kwargs)
        self.conv2 = conv_op(out_channels, out_channels, **conv_kwargs)
        self.dropout = dropout_op(**dropout_kwargs) if dropout_op else None
        self.norm = norm_op(out_channels, **norm_kwargs)
        self.nonlin = nonlin_op(**nonlin_kwargs)
        self.skip_conv = conv_op(in_channels, out_channels, kernel_size=1, stride=stride)

    @property
    def uses_skip_conv(self) -> bool:
        return self.in_channels != self.out_channels or self.stride != 1

    @property
    def dropout_p(self) -> float:
        return float(self.dropout.p) if self.dropout is not None else 0.0


class ResBlockStack(_Container):
    """network.py:419-449: num_stacks residual blocks; the first one changes the width (and strides)."""

    def __init__(self, in_channels, out_channels, stride=1, num_stacks=2, **block_kwargs):
        super().__init__()
        self.res_blocks = nn.ModuleList(
            [ResBlock(in_channels if i == 0 else out_channels, out_channels, stride=stride if i == 0 else 1,
                      **block_kwargs) for i in range(num_stacks)])


class MaxPoolBlock(_Container):
    """network.py:452-463: nn.MaxPool3d(kernel_size=2, stride=2); the channel arguments are ignored."""

    def __init__(self, in_channels, out_channels, pool_op=nn.MaxPool3d, pool_kwargs={'kernel_size': 2, 'stride': 2}):
        super().__init__()
        if pool_op is not nn.MaxPool3d or dict(pool_kwargs) != {'kernel_size': 2, 'stride': 2}:
            raise NotImplementedError("only nn.MaxPool3d(kernel_size=2, stride=2) is built")
        self.pool = pool_op(**pool_kwargs)


def none_fn(level):
    return {}


class Unet(nn.Module):
    """The reference's generic encoder / decoder (network.py:470-565): same constructor, same wiring.

    forward(x: (N, in_channels, D, H, W) float) -> (N, out_channels, D, H, W) fp32 logits, computed by
    :class:`engine.UNetEngine`.  ``precision`` ("bf16" | "fp16") selects the 16-bit storage format."""

    def __init__(self, in_channels, out_channels, paired_features, pool_block=MaxPoolBlock, pool_kwargs={},
                 pool_kwargs_fn=none_fn, up_block=UpConcat, up_kwargs={}, up_kwargs_fn=none_fn,
                 encode_block=ConvBlockStack, encode_kwargs={}, encode_kwargs_fn=none_fn, decode_block=ConvBlockStack,
                 decode_kwargs={}, decode_kwargs_fn=none_fn, conv_op=nn.Conv3d):
        super().__init__()
        num_pairs = len(paired_features)
        assert (num_pairs % 2) == 1, 'Number of paired features must be odd number.'      # network.py:491
        self.num_pool = num_pairs // 2
        assert self.num_pool > 0, 'At least one pool.'                                      # network.py:493
        if conv_op is not nn.Conv3d:
            raise NotImplementedError("conv_op must be nn.Conv3d")
        pf = paired_features
        pools, ups, encs, decs = [], [], [], []
        for i in range(self.num_pool):
            pools.append(pool_block(pf[i][1], pf[i + 1][0], **pool_kwargs_fn(i), **pool_kwargs))
            ups.append(up_block(pf[num_pairs - i - 2][1], pf[num_pairs - i - 1][0], **up_kwargs_fn(i), **up_kwargs))
            encs.append(encode_block(pf[i][0], pf[i][1], **encode_kwargs_fn(i), **encode_kwargs))
            decs.append(decode_block(pf[num_pairs - i - 1][0] + pf[i][1], pf[num_pairs - i - 1][1],
                                     **decode_kwargs_fn(i), **decode_kwargs))
        encs.append(encode_block(pf[self.num_pool][0], pf[self.num_pool][1], **encode_kwargs_fn(self.num_pool),
                                 **encode_kwargs))
        # registration order of the reference (network.py:536-547): pool, up, encode, decode, conv, fc
        self.pool_blocks = nn.ModuleList(pools)
        self.up_blocks = nn.ModuleList(ups)
        self.encode_blocks = nn.ModuleList(encs)
        self.decode_blocks = nn.ModuleList(decs)
        self.conv = conv_op(in_channels, pf[0][0], kernel_size=3, padding=1)
        self.fc = conv_op(pf[num_pairs - 1][1], out_channels, kernel_size=1)
        self.precision = "bf16"
        self.last_dropout_masks = None      # masks drawn by the most recent train-mode forward (for tests)
        self._engine = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._run(x, self)

    def invalidate_weights(self):
        """See engine.UNetEngine.invalidate_weights: call after in-place parameter writes autograd cannot see."""
        if self._engine is not None:
            self._engine.invalidate_weights()

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_weights()          # load_state_dict copies into .data: the 16-bit packs are stale

    def _run(self, x, owner):
        if self._engine is None or self._engine.owner is not owner:
            self._engine = engine.UNetEngine(self, owner)
        return self._engine.run(x)


def _encode_kwargs_fn(level):           # network.py:116-118
    return {'num_stacks': max(level, 1)}


class _ResNetBase(nn.Module):
    """Shared shape of the concrete residual nets: attributes, ``.net``, forward."""

    def _build(self, in_channels, out_channels, paired_features, attention, norm_op=None):
        self.in_channels = in_channels
        self.out_channels = out_channels
        nk = {} if norm_op is None else {'norm_op': norm_op}
        self.net = Unet(in_channels=in_channels, out_channels=out_channels, paired_features=paired_features,
                        pool_block=ResBlock, pool_kwargs={'stride': 2, **nk},
                        up_kwargs={**({'attention': True} if attention else {}), **nk},
                        encode_block=ResBlockStack, encode_kwargs=dict(nk), encode_kwargs_fn=_encode_kwargs_fn,
                        decode_block=ResBlock, decode_kwargs=dict(nk))
        self.precision = "bf16"           # "bf16" | "fp16": 16-bit storage of activations / packed weights / gradients
        self.last_dropout_masks = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net._run(x, self)

    def invalidate_weights(self):
        """Call after in-place parameter writes autograd cannot see (EMA / SWA, ``p.data.copy_``): drops the cached 16-bit
        weight packs and captured inference graphs."""
        self.net.invalidate_weights()


class ResUnet3D(_ResNetBase):
    """``ResUnet3D(num_pool=4, num_features=30, in_channels=1, out_channels=1)`` -- network.py:104-132."""

    def __init__(self, num_pool: int = 4, num_features: int = 30, in_channels: int = 1, out_channels: int = 1):
        super().__init__()
        self.num_pool = num_pool
        self.num_features = num_features
        self._build(in_channels, out_channels, generate_paired_features(num_pool, num_features), attention=False)


class ResAttrUnet3D(_ResNetBase):
    """``ResUnet3D`` with attention gates on the skips (network.py:72-101); the coarse model of the cascade."""

    def __init__(self, num_pool: int = 4, num_features: int = 30, in_channels: int = 1, out_channels: int = 1):
        super().__init__()
        self.num_pool = num_pool
        self.num_features = num_features
        self._build(in_channels, out_channels, generate_paired_features(num_pool, num_features), attention=True)


class ResAttrBNUnet3D(_ResNetBase):
    """``ResAttrUnet3D`` with BatchNorm3d (affine, running statistics) in place of InstanceNorm3d -- network.py:38-69.
    As in the reference, each ResBlock owns ONE BatchNorm3d module that normalises both of its convolutions."""

    def __init__(self, num_pool: int = 4, num_features: int = 30, in_channels: int = 1, out_channels: int = 1):
        super().__init__()
        self.num_pool = num_pool
        self.num_features = num_features
        self._build(in_channels, out_channels, generate_paired_features(num_pool, num_features), attention=True,
                    norm_op=nn.BatchNorm3d)


class ResAttrUnet3D2(_ResNetBase):
    """network.py:6-35: five poolings, widths 30/60/120/240/320/320, attention gates."""

    def __init__(self, in_channels: int = 1, out_channels: int = 1):
        super().__init__()
        down = [[30, 30], [60, 60], [120, 120], [240, 240], [320, 320]]
        self._build(in_channels, out_channels, down + [[320, 320]] + [list(p) for p in reversed(down)], attention=True)


UNet3D = ResUnet3D      # the name BASELINE.json's north_star uses

"""Drop-in replacement for the reference's ``network.py`` model classes on the hot path.

Same constructor signatures, attribute names, submodule tree and ``state_dict`` keys as
``network.ResUnet3D`` (network.py:104-132) built on ``network.Unet`` (network.py:470-565), so reference
checkpoints load unchanged and ``trainer.py`` can drive it.  The parameters are ordinary fp32
``nn.Parameter``s (held by never-called ``nn.Conv3d`` / ``nn.ConvTranspose3d`` containers, which gives
PyTorch's default initialisation); ``forward`` does not run a single PyTorch op on the activations:
it hands the whole forward / backward pass to :mod:`engine`, which enqueues the library's sm_100a
kernels.  There is no fallback -- without the CUDA library or on a CPU tensor, ``forward`` raises.
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


class ResBlock(nn.Module):
    """Parameter container of one residual block (network.py:374-416): conv1 (k3, stride s),
    conv2 (k3), skip_conv (k1, stride s; only used when in != out or s != 1, but always present,
    as in the reference, so state_dicts match)."""

    def __init__(self, in_channels: int, out_channels: int, stride: int = 1, dropout_p: float = 0.5):
        super().__init__()
        self.in_channels, self.out_channels, self.stride, self.dropout_p = in_channels, out_channels, stride, dropout_p
        self.conv1 = nn.Conv3d(in_channels, out_channels, 3, stride=stride, padding=1)
        self.conv2 = nn.Conv3d(out_channels, out_channels, 3, padding=1)
        self.skip_conv = nn.Conv3d(in_channels, out_channels, 1, stride=stride)

    @property
    def uses_skip_conv(self) -> bool:
        return self.in_channels != self.out_channels or self.stride != 1

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("blocks are parameter containers; run the enclosing network")


class ResBlockStack(nn.Module):
    """network.py:419-449: num_stacks residual blocks, the first one changing the width."""

    def __init__(self, in_channels: int, out_channels: int, num_stacks: int = 1):
        super().__init__()
        self.res_blocks = nn.ModuleList(
            [ResBlock(in_channels if j == 0 else out_channels, out_channels) for j in range(num_stacks)])


class ConvTrans3D(nn.Module):
    """network.py:298-320: ConvTranspose3d(k3,s2,p1) -> zero pad to 2x -> InstanceNorm -> LeakyReLU.
    Only the transposed conv has parameters; it sits at ``up.0`` like in the reference."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.up = nn.Sequential(nn.ConvTranspose3d(in_channels, out_channels, 3, stride=2, padding=1))


class UpConcat(nn.Module):
    """network.py:323-350 (attention=False): upsample, then concat [upsampled, skip]."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv_trans = ConvTrans3D(in_channels, out_channels)


class Unet(nn.Module):
    """Residual U-Net wiring of network.py:470-565 specialised to the blocks ResUnet3D selects
    (encode = ResBlockStack, pool = stride-2 ResBlock, decode = ResBlock, up = UpConcat)."""

    def __init__(self, in_channels: int, out_channels: int, paired_features, encode_stacks):
        super().__init__()
        assert len(paired_features) % 2 == 1, "need an odd number of feature pairs"     # network.py:491
        num_pool = len(paired_features) // 2
        assert num_pool >= 1                                                                # network.py:493
        pf = paired_features
        self.num_pool = num_pool
        self.conv = nn.Conv3d(in_channels, pf[0][0], 3, padding=1)
        self.encode_blocks = nn.ModuleList()
        self.pool_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        self.decode_blocks = nn.ModuleList()
        for i in range(num_pool):
            self.pool_blocks.append(ResBlock(pf[i][1], pf[i + 1][0], stride=2))
        for i in range(num_pool + 1):
            self.encode_blocks.append(ResBlockStack(pf[i][0], pf[i][1], encode_stacks(i)))
        n = len(pf)
        for i in range(num_pool):
            self.up_blocks.append(UpConcat(pf[n - i - 2][1] if i + 1 < num_pool else pf[num_pool][1], pf[n - i - 1][0]))
            self.decode_blocks.append(ResBlock(pf[n - i - 1][0] + pf[i][1], pf[n - i - 1][1]))
        self.fc = nn.Conv3d(pf[-1][1], out_channels, 1)


class ResUnet3D(nn.Module):
    """``ResUnet3D(num_pool=4, num_features=30, in_channels=1, out_channels=1)`` -- network.py:104-132.

    forward(x: (N, in_channels, D, H, W) float) -> (N, out_channels, D, H, W) fp32 logits."""

    def __init__(self, num_pool: int = 4, num_features: int = 30, in_channels: int = 1, out_channels: int = 1):
        super().__init__()
        self.num_pool = num_pool
        self.num_features = num_features
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.net = Unet(in_channels, out_channels, generate_paired_features(num_pool, num_features),
                        encode_stacks=lambda level: max(level, 1))                          # network.py:116-118
        self.precision = "bf16"           # "bf16" | "fp16": 16-bit storage of forward activations / weights
        self._engine = None
        self.last_dropout_masks = None      # masks drawn by the most recent train-mode forward (for tests)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._engine is None:
            self._engine = engine.ResUNetEngine(self)
        return self._engine.run(x)


UNet3D = ResUnet3D      # the name BASELINE.json's north_star uses

"""MRLA-light layer — drop-in for resnet/models/modules/mrla_light_module.py:9-74 of the reference
(same class name, constructor arguments, parameter names/shapes and forward(x) signature), executed
by the hand-written sm_100a kernels in mrla_b200/csrc instead of ~6 ATen launches.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _lib
from ..ops import LightCfg, light_tail


def eca_kernel_size(channels: int) -> int:
    """Adaptive ECA kernel size, odd (reference: mrla_light_module.py:40-42)."""
    t = int(abs((math.log(channels, 2) + 1) / 2.0))  # same expression as the reference (float log)
    return t + 1 - (t % 2)


def resolve_heads(channels: int, heads, dim_perhead) -> int:
    if heads is None and dim_perhead is None:
        raise ValueError("arguments heads and dim_perhead cannot be None at the same time !")
    return int(channels / dim_perhead) if dim_perhead is not None else heads


class mrla_light_layer(nn.Module):
    """gate(x) * dwconv3x3(x) with gate = sigmoid(<Wq*gap(x), Wk*gap(x)>_head / sqrt(d)).

    Parameters live in `Wq`/`Wk` (Conv1d 1->1, k taps, no bias) and `Wv` (depthwise Conv2d 3x3, no
    bias) exactly as in the reference so checkpoints load with strict=True and model-level init
    loops (`isinstance(m, nn.Conv2d)`) see the same modules; the nn modules are parameter holders
    only — the arithmetic runs in the fused CUDA op.
    """

    act = _lib.ACT_NONE

    def __init__(self, input_dim, heads=None, dim_perhead=None, k_size=None):
        super().__init__()
        self.input_dim = input_dim
        self.heads = resolve_heads(input_dim, heads, dim_perhead)
        if self.heads < 1 or input_dim % self.heads:
            raise ValueError(f"input_dim={input_dim} is not divisible into {self.heads} heads")
        self.k_size = eca_kernel_size(input_dim) if k_size is None else k_size
        self.avg_pool = nn.AdaptiveAvgPool2d(1)  # kept for module-tree parity; not called
        self.Wq = nn.Conv1d(1, 1, kernel_size=self.k_size, padding=(self.k_size - 1) // 2, bias=False)
        self.Wk = nn.Conv1d(1, 1, kernel_size=self.k_size, padding=(self.k_size - 1) // 2, bias=False)
        self.Wv = nn.Conv2d(input_dim, input_dim, kernel_size=3, stride=1, padding=1, groups=input_dim, bias=False)
        self._norm_fact = 1 / math.sqrt(input_dim / self.heads)
        self.sigmoid = nn.Sigmoid()

    @property
    def dim_perhead(self) -> int:
        return self.input_dim // self.heads

    def cfg(self, **kw) -> LightCfg:
        return LightCfg(dim_perhead=self.dim_perhead, k_size=self.k_size, act=self.act, **kw)

    def forward(self, x):
        return light_tail(x, None, self.Wq.weight, self.Wk.weight, self.Wv.weight, cfg=self.cfg())

"""MRLA-base layer — drop-in for resnet/models/modules/mrla_base_module.py:10-89 of the reference (same class
name, constructor arguments, parameter names/shapes, `forward(x, prev_K, prev_V) -> (out, K, V)`), executed by the
hand-written CUDA kernels in mrla_b200/csrc/base_kernels.cuh with an in-place stage cache (no torch.cat)."""
from __future__ import annotations

import math

import torch.nn as nn

from ..ops import BaseCfg, base_tail
from .mrla_light_module import eca_kernel_size, resolve_heads


class mrla_base_layer(nn.Module):
    def __init__(self, input_dim, heads=None, dim_perhead=None, k_size=None, init_cell=False):
        super().__init__()
        self.input_dim = input_dim
        self.init_cell = init_cell
        self.heads = resolve_heads(input_dim, heads, dim_perhead)
        if self.heads < 1 or input_dim % self.heads:
            raise ValueError(f"input_dim={input_dim} is not divisible into {self.heads} heads")
        self.k_size = eca_kernel_size(input_dim) if k_size is None else k_size
        self.avg_pool = nn.AdaptiveAvgPool2d(1)  # module-tree parity only
        self.Wq = nn.Conv1d(1, 1, kernel_size=self.k_size, padding=(self.k_size - 1) // 2, bias=False)
        self.Wk = nn.Conv1d(1, 1, kernel_size=self.k_size, padding=(self.k_size - 1) // 2, bias=False)
        self.Wv = nn.Conv2d(input_dim, input_dim, kernel_size=3, stride=1, padding=1, groups=input_dim, bias=False)
        self._norm_fact = 1 / math.sqrt(input_dim / self.heads)
        self.softmax = nn.Softmax(dim=-1)
        self._cap_hint = 4  # learned stage depth (grows to the deepest t seen)

    @property
    def dim_perhead(self) -> int:
        return self.input_dim // self.heads

    def cfg(self, **kw) -> BaseCfg:
        return BaseCfg(dim_perhead=self.dim_perhead, k_size=self.k_size, **kw)

    def run(self, x, prev_K, prev_V, cfg, gamma=None, beta=None, running_mean=None, running_var=None,
            drop_scale=None, out=None):
        y, K, V = base_tail(x, prev_K, prev_V, self.Wq.weight, self.Wk.weight, self.Wv.weight, gamma, beta,
                            running_mean, running_var, drop_scale, init_cell=self.init_cell, cfg=cfg,
                            cap_hint=self._cap_hint, out=out)
        cache = K._mrla_cache
        if self.init_cell:
            cache.owner = self
        owner = getattr(cache, "owner", None)
        if owner is not None and cache.t > owner._cap_hint:
            owner._cap_hint = cache.t
        return y, K, V

    def forward(self, x, prev_K, prev_V):
        return self.run(x, prev_K, prev_V, self.cfg())

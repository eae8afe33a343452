"""DeiT + MRLA-base (token layout) — mirror of `mrlab_layer` / `mrlab_module` of deit/deit_mrla_base.py:120-243
(reference).  The K/V cache of a 4-block group lives in the in-place StageCache of mrla_b200.ops."""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .deit_mrla_light import tokens_as_image
from .modules.mrla_base_module import mrla_base_layer
from .ops import deit_base_module

__all__ = ["mrlab_layer", "mrlab_module"]


class mrlab_layer(mrla_base_layer):
    """Identical arithmetic to mrla_base_layer (reference deit_mrla_base.py:166-201)."""


class mrlab_module(nn.Module):
    def __init__(self, input_dim, dim_perhead, init_cell=False, channel_wise=False,
                 norm_layer=partial(nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.dim_perhead = 1 if channel_wise else dim_perhead
        self.init_cell = init_cell
        self.normx = norm_layer(input_dim)
        self.mrla = mrlab_layer(input_dim=input_dim, dim_perhead=self.dim_perhead, init_cell=init_cell)

    def forward(self, xt, prev_k, prev_v):
        if self.init_cell:
            prev_k = prev_v = None
        ln = self.normx
        if type(ln) is nn.LayerNorm and ln.elementwise_affine and ln.bias is not None \
                and tuple(ln.normalized_shape) == (xt.shape[-1],):
            m = self.mrla
            res = deit_base_module(xt, prev_k, prev_v, ln.weight, ln.bias, m.Wq.weight, m.Wk.weight, m.Wv.weight,
                                   init_cell=self.init_cell, cfg=m.cfg(), eps=ln.eps, cap_hint=m._cap_hint)
            if res is not None:   # one autograd node, no library launches (token LayerNorm kernel + MRLA-base tail kernels)
                out, kt, vt = res
                cache = kt._mrla_cache
                if self.init_cell:
                    cache.owner = m
                owner = getattr(cache, "owner", None)
                if owner is not None and cache.t > owner._cap_hint:
                    owner._cap_hint = cache.t
                return out, kt, vt
        xn = self.normx(xt)
        img, kt, vt = self.mrla(tokens_as_image(xn[:, 1:]), prev_k, prev_v)
        b, c, s, _ = img.shape
        tokens = img.permute(0, 2, 3, 1).reshape(b, s * s, c)
        return torch.cat((xn[:, :1], tokens), dim=1), kt, vt

"""DeiT + MRLA-light (token layout) — host-side mirror of the MRLA pieces of deit/deit_mrla_light.py (reference):
`mrlal_layer` (:117-180, GELU on V), `mrlal_module` (:183-209: LayerNorm of x and o, cls-token pass-through, lambda
recurrence on the 14x14 token image) with identical class names, constructor arguments and parameter names.
The token image [B, n-1, C] is read in place as an NHWC activation with batch stride n*C (no permute / copy);
round 2: the whole module (both LayerNorms, cls pass-through, gate, GELU-V, lambda) is ONE kernel per direction
(csrc/deit_fused.cuh, one CTA per sample) wherever a sample fits a CTA's shared memory; other shapes keep the round-1 path."""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .modules.mrla_light_module import mrla_light_layer
from .ops import deit_light_module, deit_light_supported, light_tail

__all__ = ["mrlal_layer", "mrlal_module", "tokens_as_image"]


def tokens_as_image(tok: torch.Tensor) -> torch.Tensor:
    """[B, m, C] token slice -> logical [B, C, s, s] view (channels-last strides, zero copies); m must be a square
    (the reference takes int(sqrt(n-1)), deit_mrla_light.py:201)."""
    b, m, c = tok.shape
    s = int(math.sqrt(m))
    if s * s != m:
        raise ValueError(f"token count {m} is not a perfect square")
    return tok.as_strided((b, c, s, s), (tok.stride(0), 1, s * tok.stride(1), tok.stride(1)), tok.storage_offset())


class mrlal_layer(mrla_light_layer):
    """MRLA-light layer with V = GELU(dwconv3x3(x)) (reference deit_mrla_light.py:153,166-167)."""
    act = _lib.ACT_GELU

    def __init__(self, input_dim, heads=None, dim_perhead=None, k_size=None):
        super().__init__(input_dim, heads=heads, dim_perhead=dim_perhead, k_size=k_size)
        self.act_v = nn.GELU()  # module-tree parity; the activation is applied inside the fused kernel

    def forward(self, x):
        if x.dim() == 4 and x.is_contiguous() and x.shape[1] > 1 and x.shape[2] * x.shape[3] > 1:
            x = x.contiguous(memory_format=torch.channels_last)  # GELU variant is implemented for NHWC only
        return super().forward(x)


class mrlal_module(nn.Module):
    def __init__(self, input_dim, dim_perhead, norm_layer=partial(nn.LayerNorm, eps=1e-6)):
        super().__init__()
        self.dim_perhead = dim_perhead
        self.mrla = mrlal_layer(input_dim=input_dim, dim_perhead=self.dim_perhead)
        self.lambda_t = nn.Parameter(torch.randn(input_dim))
        self.normx = norm_layer(input_dim)
        self.normo = norm_layer(input_dim)

    def forward(self, xt, ot_1):
        m = self.mrla
        plain_ln = all(type(ln) is nn.LayerNorm and ln.elementwise_affine and ln.bias is not None
                       and tuple(ln.normalized_shape) == (xt.shape[-1],) for ln in (self.normx, self.normo))
        if (plain_ln and self.normx.eps == self.normo.eps and ot_1.shape == xt.shape and ot_1.dtype == xt.dtype
                and xt.is_contiguous() and ot_1.is_contiguous() and deit_light_supported(xt, self.dim_perhead, m.k_size)):
            # one kernel per direction: LN_x, LN_o, cls pass-through, gate, GELU-V, lambda (csrc/deit_fused.cuh)
            return deit_light_module(xt, ot_1, self.normx.weight, self.normx.bias, self.normo.weight, self.normo.bias,
                                     m.Wq.weight, m.Wk.weight, m.Wv.weight, self.lambda_t, dim_perhead=self.dim_perhead,
                                     k_size=m.k_size, eps=self.normx.eps)
        # other norm layers / shapes the fused kernel does not take (C % 64 != 0, a sample larger than one CTA's shared
        # memory): LayerNorms through the library, MRLA through the generic tail kernels
        xn = self.normx(xt)
        on = self.normo(ot_1)
        m = self.mrla
        tokens = light_tail(tokens_as_image(xn[:, 1:]), tokens_as_image(on[:, 1:]), m.Wq.weight, m.Wk.weight,
                            m.Wv.weight, self.lambda_t, cfg=m.cfg())  # [B,C,s,s], NHWC memory
        b, c, s, _ = tokens.shape
        tokens = tokens.permute(0, 2, 3, 1).reshape(b, s * s, c)
        return torch.cat((xn[:, :1], tokens), dim=1)

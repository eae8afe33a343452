"""Stochastic depth for the MRLA branch.

Same contract as the reference's `DropPath` (resnet/models/utils/drop.py:7-35): in training,
each sample's branch is kept with probability 1-p and rescaled by 1/(1-p).  Here the Bernoulli
draw is exposed as a per-sample scale vector `m_b` so the fused tail kernel can consume it
(the mask multiply never becomes a separate pass over the activation).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


def drop_scale(x: torch.Tensor, drop_prob: Optional[float], training: bool) -> Optional[torch.Tensor]:
    """[B] fp32 scale (0 or 1/keep), or None when DropPath is the identity.

    Draws `torch.rand((B,1,...,1), dtype=x.dtype, device=x.device)` — the very call the reference
    makes (drop.py:21) — so both implementations consume the RNG stream identically."""
    if not drop_prob or not training:
        return None
    keep = 1.0 - drop_prob
    u = torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)
    return ((u + keep).floor_().float() / keep).reshape(-1)


class DropPath(nn.Module):
    """Drop-in for the reference module; usable standalone (elementwise) or fused via `drop_scale`."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def scale(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        return drop_scale(x, self.drop_prob, self.training)

    def forward(self, x):
        m = self.scale(x)
        if m is None:
            return x
        return x * m.to(x.dtype).view((-1,) + (1,) * (x.ndim - 1))

    def extra_repr(self):
        return f"drop_prob={self.drop_prob}"

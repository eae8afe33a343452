#!/usr/bin/env python
"""Run the fused DeiT mrlal_module fwd+bwd a few times at the DeiT-tiny shape — ncu target for k_deit_light_*."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrla_b200.deit_mrla_light import mrlal_module
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mod = mrlal_module(192, 16).to(dev).bfloat16()
x = torch.randn(B, 197, 192, device=dev, dtype=torch.bfloat16).requires_grad_()
o = torch.randn(B, 197, 192, device=dev, dtype=torch.bfloat16).requires_grad_()
dy = torch.randn(B, 197, 192, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    y = x + mod(x, o)
    y.backward(dy)
torch.cuda.synchronize()

#!/usr/bin/env python
"""Run the folded MRLA-light tail op (as resnet50_mrlal calls it) a few times at one stage shape — ncu target."""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrla_b200 import _lib
from mrla_b200.ops import LightCfg, light_tail
ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=256); ap.add_argument("--HW", type=int, default=56)
ap.add_argument("--B", type=int, default=256); ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--no-bn3", action="store_true", help="z is the bn3 OUTPUT (sweep-1 MODE 5) instead of the raw conv3 output + bn3 coefficients (MODE 6)")
a = ap.parse_args()
dev = torch.device("cuda:0")
k = 7 if a.C == 2048 else 5
mk = lambda: torch.randn(a.B, a.C, a.HW, a.HW, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
z, idt, dy = mk().requires_grad_(), mk().requires_grad_(), mk()
P = [torch.randn(k, device=dev).requires_grad_(), torch.randn(k, device=dev).requires_grad_(),
     (torch.randn(a.C, 1, 3, 3, device=dev) * 0.05).requires_grad_(), torch.randn(a.C, 1, 1, device=dev).requires_grad_(),
     torch.ones(a.C, device=dev).requires_grad_(), torch.zeros(a.C, device=dev).requires_grad_()]
rm, rv = torch.zeros(a.C, device=dev), torch.ones(a.C, device=dev)
cfg = LightCfg(dim_perhead=32, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True, fuse_add_relu=True)
zc = None if a.no_bn3 else torch.stack([torch.rand(a.C, device=dev) + 0.5, torch.randn(a.C, device=dev) * 0.1]).contiguous()
for _ in range(a.iters):
    y = light_tail(z, idt, *P, rm, rv, None, cfg=cfg, z_coef=zc)
    y.backward(dy)
    z.grad = None; idt.grad = None
torch.cuda.synchronize()

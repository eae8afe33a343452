#!/usr/bin/env python
"""Run ops.bn_act (channels-last BatchNorm + ReLU, fwd + bwd) a few times at one activation shape — ncu target for the
k_bn_* kernels (bn1/bn2 of a bottleneck: --relu; bn3 / downsample: no ReLU)."""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrla_b200.ops import bn_act
ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=64); ap.add_argument("--HW", type=int, default=56)
ap.add_argument("--B", type=int, default=256); ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--relu", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
x = torch.randn(a.B, a.C, a.HW, a.HW, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
dy = torch.randn_like(x)
bn = torch.nn.BatchNorm2d(a.C).to(dev).train()
for _ in range(a.iters):
    y = bn_act(x, bn, relu=a.relu)
    y.backward(dy)
    x.grad = None
torch.cuda.synchronize()

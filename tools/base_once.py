#!/usr/bin/env python
"""Run one MRLA-base stage (T blocks threading K/V through ops.base_tail, fwd + bwd) a few times at one ResNet stage shape —
ncu target for the k_base_* kernels (tools/profile_base.sh)."""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrla_b200 import _lib
from mrla_b200.modules.mrla_light_module import eca_kernel_size
from mrla_b200.ops import BaseCfg, base_tail
ap = argparse.ArgumentParser()
ap.add_argument("--C", type=int, default=1024); ap.add_argument("--HW", type=int, default=14)
ap.add_argument("--B", type=int, default=256); ap.add_argument("--T", type=int, default=6)
ap.add_argument("--d", type=int, default=16); ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, C, HW, T, d = a.B, a.C, a.HW, a.T, a.d
k = eca_kernel_size(C)
mk = lambda: torch.randn(B, C, HW, HW, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
xs = [torch.relu(mk()).requires_grad_() for _ in range(T)]
dys = [mk() for _ in range(T)]
Ps = [dict(wq=torch.randn(k, device=dev), wk=torch.randn(k, device=dev), wv=torch.randn(C, 1, 3, 3, device=dev) * 0.3,
           gamma=torch.ones(C, device=dev), beta=torch.zeros(C, device=dev)) for _ in range(T)]
for P in Ps:
    for v in P.values():
        v.requires_grad_()
cfg = BaseCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, relu=True, residual=True)
for _ in range(a.iters):
    kk = vv = None
    ys = []
    for t in range(T):
        P = Ps[t]
        y, kk, vv = base_tail(xs[t], kk, vv, P["wq"], P["wk"], P["wv"], P["gamma"], P["beta"], torch.zeros(C, device=dev),
                              torch.ones(C, device=dev), None, init_cell=(t == 0), cfg=cfg, cap_hint=T)
        ys.append(y)
    torch.autograd.backward(ys, dys)
    for x in xs:
        x.grad = None
torch.cuda.synchronize()

#!/usr/bin/env python
"""Device time of the stem max pooling (256 x 64 x 112 x 112 bf16 NHWC) forward / backward against at::max_pool2d."""
import os, sys, torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrla_b200.ops import max_pool
dev = torch.device("cuda:0")
x = torch.relu(torch.randn(256, 64, 112, 112, device=dev)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
pool = nn.MaxPool2d(3, 2, 1)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, f in (("mrla_b200", lambda t: max_pool(t, pool)), ("aten", lambda t: F.max_pool2d(t, 3, 2, 1))):
    xr = x.clone().requires_grad_()
    y = f(xr)
    dy = torch.randn_like(y)
    tf = timeit(lambda: f(xr))
    tb = timeit(lambda: torch.autograd.grad(y, xr, dy, retain_graph=True))
    print(f"{name}: fwd {tf*1e3:.0f} us  bwd {tb*1e3:.0f} us")

#!/usr/bin/env python
"""Poison the caching allocator with NaN so that any element a kernel forgets to write shows up in the outputs
(initcheck reported an uninitialised read inside an ATen float->double copy of one of our results)."""
import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
from conftest import golden_names, load_golden
import test_light_gpu as T
dev = torch.device("cuda:0")
def poison():
    bufs = [torch.full((n,), float("nan"), device=dev) for n in (1 << 24, 1 << 20, 1 << 16, 1 << 12, 257, 64)] * 4
    del bufs
for name in golden_names("light_tail") + golden_names("light_layer"):
    g = load_golden(name)
    for layout in ("nchw", "nhwc"):
        if layout == "nhwc" and g["C"] % 4:
            continue
        for dtype in (torch.float32, torch.bfloat16):
            poison()
            if name.startswith("light_tail"):
                x, o, y, P, rm, rv = T._run_tail(g, dtype, layout, dev)
                outs = dict(y=y, dx=x.grad, do=o.grad, rm=rm, rv=rv, **{k: v.grad for k, v in P.items()})
            else:
                continue
            bad = [k for k, v in outs.items() if v is not None and not torch.isfinite(v.float()).all()]
            if bad:
                print(name, layout, dtype, "NaN in", bad)
print("done")

#!/usr/bin/env python
"""Kernel-level timing of the fused MRLA-light tail at the ResNet-50 stage shapes.

Reports fwd / bwd / total ms per block and algorithmic GB/s = 8*N*sizeof / time against the measured
HBM copy peak in MEASURED_PEAKS.json.  L2 is flushed between iterations (cold) unless --warm."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrla_b200 import _lib  # noqa: E402
from mrla_b200.ops import LightCfg, light_tail  # noqa: E402

STAGES = [(256, 56, 3), (512, 28, 4), (1024, 14, 6), (2048, 7, 3)]


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--layout", default="nhwc")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warm", action="store_true")
    ap.add_argument("--stages", default="0,1,2,3")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dtype = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[args.dtype]
    peak, how = peak_gbs()
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    for si in [int(s) for s in args.stages.split(",")]:
        C, HW, nblk = STAGES[si]
        B, d = args.batch, 32
        k = 7 if C == 2048 else 5
        mk = lambda: torch.randn(B, C, HW, HW, device=dev, dtype=dtype)
        x, o, dy = torch.relu(mk()), mk(), mk()
        if args.layout == "nhwc":
            x, o, dy = (t.contiguous(memory_format=torch.channels_last) for t in (x, o, dy))
        x.requires_grad_(); o.requires_grad_()
        P = [torch.randn(k, device=dev).requires_grad_(), torch.randn(k, device=dev).requires_grad_(),
             (torch.randn(C, 1, 3, 3, device=dev) * 0.47).requires_grad_(), torch.randn(C, 1, 1, device=dev).requires_grad_(),
             torch.ones(C, device=dev).requires_grad_(), torch.zeros(C, device=dev).requires_grad_()]
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        cfg = LightCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True)
        tf, tb = [], []
        for it in range(args.iters + 5):
            if not args.warm:
                flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            y = light_tail(x, o, *P, rm, rv, None, cfg=cfg)
            e1.record()
            if not args.warm:
                flush.zero_()
            e1b = torch.cuda.Event(enable_timing=True)
            e1b.record()
            y.backward(dy)
            e2.record()
            torch.cuda.synchronize()
            if it >= 5:
                tf.append(e0.elapsed_time(e1)); tb.append(e1b.elapsed_time(e2))
            x.grad = None; o.grad = None
        tf.sort(); tb.sort()
        f, b = tf[len(tf) // 2], tb[len(tb) // 2]
        N = B * C * HW * HW
        es = x.element_size()
        gb = 8 * N * es / 1e9
        row = dict(stage=si + 1, C=C, HW=HW, B=B, dtype=args.dtype, layout=args.layout, fwd_ms=round(f, 4), bwd_ms=round(b, 4),
                   total_ms=round(f + b, 4), alg_GB=round(gb, 3), alg_GBps=round(gb / (f + b) * 1e3, 1),
                   frac_of_peak=round(gb / (f + b) * 1e3 / peak, 3), fwd_GBps=round(3 * N * es / f / 1e6, 1),
                   bwd_GBps=round(5 * N * es / b / 1e6, 1), blocks=nblk, cold=not args.warm)
        rows.append(row)
        print(json.dumps(row), flush=True)
        del x, o, dy, y
    tot_ms = sum(r["total_ms"] * r["blocks"] for r in rows)
    tot_gb = sum(r["alg_GB"] * r["blocks"] for r in rows)
    summ = dict(all_tails_ms=round(tot_ms, 3), alg_GB=round(tot_gb, 2), alg_GBps=round(tot_gb / tot_ms * 1e3, 1),
                frac=round(tot_gb / tot_ms * 1e3 / peak, 3), peak_GBps=peak, peak_kind=how)
    print(json.dumps(summ))
    if args.json:
        json.dump(dict(rows=rows, summary=summ), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()

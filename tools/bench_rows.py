#!/usr/bin/env python
"""Kernel-level measurements for every SURVEY.md §8 row, next to eager PyTorch on the same B200.

"eager" = the oracle restatement evaluated on the GPU in the same dtype: it issues the same ATen calls as the
reference modules (adaptive-avg-pool / conv1d / depthwise conv2d / einsum / batch_norm / ...), so it stands in for
"the unmodified reference module on B200" (the reference tree itself does not travel to the GPU box).
Only this script and tests use the oracle; the product path never does.

    python tools/bench_rows.py --json gpurun_out/rows.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrla_b200 import _lib  # noqa: E402
from mrla_b200.ops import BaseCfg, LightCfg, base_tail, light_tail  # noqa: E402
from oracle import mrla_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
FLUSH = None


def timeit(fn, iters=10, warm=3):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(warm + iters):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def eca_k(C):
    return O.eca_kernel_size(C)


def light_case(B, C, H, W, d, dt, layout="nhwc", act=None, bn=True):
    k = eca_k(C)
    mk = lambda: torch.randn(B, C, H, W, device=dev, dtype=dt)
    x, o, dy = torch.relu(mk()), mk(), mk()
    if layout == "nhwc":
        x, o, dy = (t.contiguous(memory_format=torch.channels_last) for t in (x, o, dy))
    x.requires_grad_(); o.requires_grad_()
    P = dict(wq=torch.randn(k, device=dev), wk=torch.randn(k, device=dev), wv=torch.randn(C, 1, 3, 3, device=dev) * 0.3,
             lam=torch.randn(C, 1, 1, device=dev), gamma=torch.ones(C, device=dev), beta=torch.zeros(C, device=dev))
    for v in P.values():
        v.requires_grad_()
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    cfg = LightCfg(dim_perhead=d, k_size=k, act=_lib.ACT_GELU if act else _lib.ACT_NONE,
                   bn_mode=_lib.BN_TRAIN if bn else _lib.BN_NONE, residual=bn)

    def mine():
        y = light_tail(x, o, P["wq"], P["wk"], P["wv"], P["lam"], P["gamma"] if bn else None, P["beta"] if bn else None,
                       rm if bn else None, rv if bn else None, None, cfg=cfg)
        y.backward(dy)
        x.grad = None; o.grad = None

    Pl = {n: v.detach().to(dt).requires_grad_() for n, v in P.items()}

    def eager():
        if bn:
            y, _, _ = O.light_tail(x, o, Pl["wq"], Pl["wk"], Pl["wv"], Pl["lam"], C // d, Pl["gamma"], Pl["beta"],
                                   rm.to(dt), rv.to(dt), training=True)
        else:
            y = O.light_layer(x, Pl["wq"], Pl["wk"], Pl["wv"], C // d, act=act) + Pl["lam"].view(1, C, 1, 1) * o
        y.backward(dy)
        x.grad = None; o.grad = None

    tm, te = timeit(mine), timeit(eager)
    nbytes = 8 * B * C * H * W * x.element_size()
    return dict(B=B, C=C, H=H, W=W, d=d, dtype=str(dt).split(".")[-1], layout=layout, ms=round(tm, 4),
                eager_ms=round(te, 4), speedup_vs_eager=round(te / tm, 2), alg_GBps=round(nbytes / tm / 1e6, 1))


def base_case(B, C, HW, d, T, dt, layout="nhwc"):
    k = eca_k(C)
    mk = lambda: torch.randn(B, C, HW, HW, device=dev, dtype=dt)
    xs = [torch.relu(mk()) for _ in range(T)]
    dys = [mk() for _ in range(T)]
    if layout == "nhwc":
        xs = [t.contiguous(memory_format=torch.channels_last) for t in xs]
        dys = [t.contiguous(memory_format=torch.channels_last) for t in dys]
    for t in xs:
        t.requires_grad_()
    Ps = [dict(wq=torch.randn(k, device=dev), wk=torch.randn(k, device=dev), wv=torch.randn(C, 1, 3, 3, device=dev) * 0.3,
               gamma=torch.ones(C, device=dev), beta=torch.zeros(C, device=dev)) for _ in range(T)]
    for P in Ps:
        for v in P.values():
            v.requires_grad_()
    cfg = BaseCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, relu=True, residual=True)

    def mine():
        kk = vv = None
        ys = []
        for t in range(T):
            P = Ps[t]
            y, kk, vv = base_tail(xs[t], kk, vv, P["wq"], P["wk"], P["wv"], P["gamma"], P["beta"],
                                  torch.zeros(C, device=dev), torch.ones(C, device=dev), None, init_cell=(t == 0),
                                  cfg=cfg, cap_hint=T)
            ys.append(y)
        torch.autograd.backward(ys, dys)
        for t_ in xs:
            t_.grad = None

    Pls = [{n: v.detach().to(dt).requires_grad_() for n, v in P.items()} for P in Ps]

    def eager():
        kk = vv = None
        ys = []
        for t in range(T):
            P = Pls[t]
            y, kk, vv, _, _ = O.base_tail(xs[t], kk, vv, P["wq"], P["wk"], P["wv"], C // d, t == 0, P["gamma"], P["beta"],
                                          torch.zeros(C, device=dev, dtype=dt), torch.ones(C, device=dev, dtype=dt))
            ys.append(y)
        torch.autograd.backward(ys, dys)
        for t_ in xs:
            t_.grad = None

    tm, te = timeit(mine, iters=5, warm=2), timeit(eager, iters=5, warm=2)
    N = B * C * HW * HW
    nbytes = sum((4 * t + 3) * N * xs[0].element_size() for t in range(1, T + 1))
    return dict(B=B, C=C, HW=HW, d=d, T=T, dtype=str(dt).split(".")[-1], layout=layout, ms_stage=round(tm, 4),
                eager_ms_stage=round(te, 4), speedup_vs_eager=round(te / tm, 2), alg_GBps=round(nbytes / tm / 1e6, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None, help="base: only the MRLA-base stage rows; wide: only the W > 56 rows")
    args = ap.parse_args()
    bf = torch.bfloat16
    out = {}
    if args.only == "base":
        out["A6_base_stage_resnet50_bf16_nhwc_B256"] = [base_case(256, C, H, 16, T, bf) for C, H, T in
                                                        ((256, 56, 3), (512, 28, 4), (1024, 14, 6), (2048, 7, 3))]
        print(json.dumps(out["A6_base_stage_resnet50_bf16_nhwc_B256"]), flush=True)
        if args.json:
            json.dump(out, open(args.json, "w"), indent=1)
        return
    if args.only == "wide":
        # feature maps wider than 56: a detection backbone at 800x1216 with 2 images per GPU (mmdetection's
        # resnet_mrlal.py) and a 448^2 classification batch
        out["F4_light_tail_mmdet_800x1216_bf16_nhwc_B2"] = [light_case(2, C, H, W, 32, bf) for C, H, W in
                                                            ((256, 200, 304), (512, 100, 152), (1024, 50, 76), (2048, 25, 38))]
        print(json.dumps(out["F4_light_tail_mmdet_800x1216_bf16_nhwc_B2"]), flush=True)
        out["A4_light_tail_resnet50_448px_bf16_nhwc_B64"] = [light_case(64, C, H, H, 32, bf) for C, H in
                                                             ((256, 112), (512, 56), (1024, 28), (2048, 14))]
        print(json.dumps(out["A4_light_tail_resnet50_448px_bf16_nhwc_B64"]), flush=True)
        if args.json:
            json.dump(out, open(args.json, "w"), indent=1)
        return
    out["A4_light_tail_resnet50_bf16_nhwc_B256"] = [light_case(256, C, H, H, 32, bf) for C, H in
                                                    ((256, 56), (512, 28), (1024, 14), (2048, 7))]
    print(json.dumps(out["A4_light_tail_resnet50_bf16_nhwc_B256"]), flush=True)
    out["A4_light_tail_resnet50_fp32_nchw_B32"] = [light_case(32, C, H, H, 32, torch.float32, "nchw") for C, H in
                                                   ((256, 56), (512, 28), (1024, 14), (2048, 7))]
    print(json.dumps(out["A4_light_tail_resnet50_fp32_nchw_B32"]), flush=True)
    if not args.quick:
        out["A9_light_tail_efficientnet_b0_shapes_bf16_nhwc_B384"] = [
            light_case(384, C, H, H, 8, bf) for C, H in ((16, 112), (24, 56), (40, 28), (80, 14), (112, 14), (192, 7), (320, 7))]
        print(json.dumps(out["A9_light_tail_efficientnet_b0_shapes_bf16_nhwc_B384"]), flush=True)
        out["A7_deit_tiny_token_image_bf16_B256"] = [light_case(256, 192, 14, 14, 16, bf, act="gelu", bn=False)]
        print(json.dumps(out["A7_deit_tiny_token_image_bf16_B256"]), flush=True)
        out["A6_base_stage_resnet50_bf16_nhwc_B256"] = [base_case(256, C, H, 16, T, bf) for C, H, T in
                                                        ((256, 56, 3), (512, 28, 4), (1024, 14, 6), (2048, 7, 3))]
        print(json.dumps(out["A6_base_stage_resnet50_bf16_nhwc_B256"]), flush=True)
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()

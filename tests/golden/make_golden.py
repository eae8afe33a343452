#!/usr/bin/env python
"""Freeze golden vectors from the UNMODIFIED reference (run in the dev container only).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.pt.gz

The reference ships no tests or fixtures, so parity is pinned by executing its own
classes (imported from /root/reference through oracle/ref_loader.py) in fp64 on seeded
inputs and committing inputs, parameters, outputs, gradients and BN buffers.  All
inputs are rounded to bf16-representable values first so the same fixtures drive the
fp32 and the bf16 GPU parity tests without any input rounding error.

Reference entry points exercised (paths relative to /root/reference):
  light_layer  resnet/models/modules/mrla_light_module.py:9-74
  light_tail   resnet/models/resnet_mrla_light.py:32-43,85-86,116 + utils/drop.py:7-35
  base_stage   resnet/models/modules/mrla_base_module.py:10-89, resnet_mrla_base.py:32-51,124-127
  deit_light   deit/deit_mrla_light.py:117-209,234
  deit_base    deit/deit_mrla_base.py:120-243,274
"""
import gzip
import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402


def q(t):
    """Round to bf16-representable values, return fp64."""
    return t.float().bfloat16().double()


def randq(*shape):
    return q(torch.randn(*shape))


def set_params_q(mod):
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(q(p))


def grads_of(mod):
    return {n: p.grad.clone() for n, p in mod.named_parameters()}


def light_layer_case(seed, B, C, H, W, d):
    torch.manual_seed(seed)
    L = ref_loader.light_layer_mod().mrla_light_layer(C, dim_perhead=d).double()
    set_params_q(L)
    x = torch.relu(randq(B, C, H, W)).requires_grad_()
    dy = randq(B, C, H, W)
    y = L(x)
    (y * dy).sum().backward()
    return dict(kind="light_layer", B=B, C=C, H=H, W=W, d=d, k=L.k_size, x=x.detach(), dy=dy, y=y.detach(),
                params={n: p.detach().clone() for n, p in L.named_parameters()},
                dx=x.grad.clone(), dparams=grads_of(L))


def light_tail_case(seed, B, C, H, W, d, drop_path, training, relu_x=True):
    torch.manual_seed(seed)
    rl = ref_loader.resnet_light()

    class M(rl.mrla_module):
        dim_perhead = d

    mod = M(C).double()
    bn = nn.BatchNorm2d(C).double()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.5 * torch.randn(C))
        bn.bias.copy_(0.3 * torch.randn(C))
        bn.running_mean.copy_(0.2 * torch.randn(C))
        bn.running_var.copy_(0.5 + torch.rand(C))
    set_params_q(mod)
    set_params_q(bn)
    with torch.no_grad():
        bn.running_mean.copy_(q(bn.running_mean))
        bn.running_var.copy_(q(bn.running_var))
    dp = ref_loader.drop_mod().DropPath(drop_path) if drop_path > 0 else nn.Identity()
    for m in (mod, bn, dp):
        m.train(training)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    x = randq(B, C, H, W)
    if relu_x:
        x = torch.relu(x)
    x.requires_grad_()
    o = randq(B, C, H, W).requires_grad_()
    dy = randq(B, C, H, W)
    # DropPath draws torch.rand((B,1,1,1), dtype=x.dtype) (utils/drop.py:21); replay it to record m_b
    drop_scale = None
    state = torch.random.get_rng_state()
    if drop_path > 0 and training:
        keep = 1 - drop_path
        drop_scale = ((keep + torch.rand((B, 1, 1, 1), dtype=torch.float64)).floor_() / keep).reshape(B)
    torch.random.set_rng_state(state)
    y = x + dp(bn(mod(x, o)))
    (y * dy).sum().backward()
    params = {n: p.detach().clone() for n, p in mod.named_parameters()}
    params.update({"bn." + n: p.detach().clone() for n, p in bn.named_parameters()})
    dparams = grads_of(mod)
    dparams.update({"bn." + n: g for n, g in grads_of(bn).items()})
    return dict(kind="light_tail", B=B, C=C, H=H, W=W, d=d, k=mod.mrla.k_size, training=training,
                drop_path=drop_path, drop_scale=drop_scale, eps=bn.eps, momentum=bn.momentum,
                x=x.detach(), o=o.detach(), dy=dy, y=y.detach(), params=params,
                running_mean0=rm0, running_var0=rv0,
                running_mean1=bn.running_mean.clone(), running_var1=bn.running_var.clone(),
                num_batches_tracked1=int(bn.num_batches_tracked),
                dx=x.grad.clone(), do=o.grad.clone(), dparams=dparams)


def base_stage_case(seed, B, C, H, W, d, T, drop_path, training=True):
    """T consecutive MRLA-base block tails of one stage (first has init_cell=True)."""
    torch.manual_seed(seed)
    rb = ref_loader.resnet_base()

    class M(rb.mrla_module):
        dim_perhead = d

    mods, bns = [], []
    for t in range(T):
        m = M(C, init_cell=(t == 0)).double()
        bn = nn.BatchNorm2d(C).double()
        with torch.no_grad():
            bn.weight.copy_(1 + 0.5 * torch.randn(C))
            bn.bias.copy_(0.3 * torch.randn(C))
        set_params_q(m)
        set_params_q(bn)
        m.train(training)
        bn.train(training)
        mods.append(m)
        bns.append(bn)
    dp = ref_loader.drop_mod().DropPath(drop_path) if drop_path > 0 else nn.Identity()
    dp.train(training)
    xs = [torch.relu(randq(B, C, H, W)).requires_grad_() for _ in range(T)]
    dys = [randq(B, C, H, W) for _ in range(T)]
    drop_scales = []
    k = v = None
    ys = []
    loss = 0
    for t in range(T):
        state = torch.random.get_rng_state()
        if drop_path > 0 and training:
            keep = 1 - drop_path
            drop_scales.append(((keep + torch.rand((B, 1, 1, 1), dtype=torch.float64)).floor_() / keep).reshape(B))
        torch.random.set_rng_state(state)
        # resnet_mrla_base.py:124-127
        attn, k, v = mods[t](xs[t], k, v)
        attn = torch.relu(bns[t](attn))
        y = xs[t] + dp(attn)
        ys.append(y)
        loss = loss + (y * dys[t]).sum()
    loss.backward()
    out = dict(kind="base_stage", B=B, C=C, H=H, W=W, d=d, T=T, k=mods[0].mrla.k_size, training=training,
               drop_path=drop_path, drop_scales=drop_scales, eps=bns[0].eps, momentum=bns[0].momentum,
               xs=[x.detach() for x in xs], dys=dys, ys=[y.detach() for y in ys],
               K=k.detach(), V=v.detach(), dxs=[x.grad.clone() for x in xs], blocks=[])
    for t in range(T):
        params = {n: p.detach().clone() for n, p in mods[t].named_parameters()}
        params.update({"bn." + n: p.detach().clone() for n, p in bns[t].named_parameters()})
        dparams = grads_of(mods[t])
        dparams.update({"bn." + n: g for n, g in grads_of(bns[t]).items()})
        out["blocks"].append(dict(params=params, dparams=dparams,
                                  running_mean1=bns[t].running_mean.clone(),
                                  running_var1=bns[t].running_var.clone()))
    return out


def deit_light_case(seed, B, S, C, d):
    torch.manual_seed(seed)
    dl = ref_loader.deit_light()
    mod = dl.mrlal_module(C, d).double()
    with torch.no_grad():
        for ln in (mod.normx, mod.normo):
            ln.weight.copy_(1 + 0.3 * torch.randn(C))
            ln.bias.copy_(0.2 * torch.randn(C))
    set_params_q(mod)
    n = S * S + 1
    x = randq(B, n, C).requires_grad_()
    o = randq(B, n, C).requires_grad_()
    dy = randq(B, n, C)
    y = x + mod(x, o)  # deit_mrla_light.py:234
    (y * dy).sum().backward()
    return dict(kind="deit_light", B=B, S=S, C=C, d=d, x=x.detach(), o=o.detach(), dy=dy, y=y.detach(),
                params={n_: p.detach().clone() for n_, p in mod.named_parameters()},
                dx=x.grad.clone(), do=o.grad.clone(), dparams=grads_of(mod))


def deit_base_case(seed, B, S, C, d, T):
    torch.manual_seed(seed)
    db = ref_loader.deit_base()
    mods = []
    for t in range(T):
        m = db.mrlab_module(C, d, init_cell=(t == 0)).double()
        with torch.no_grad():
            m.normx.weight.copy_(1 + 0.3 * torch.randn(C))
            m.normx.bias.copy_(0.2 * torch.randn(C))
        set_params_q(m)
        mods.append(m)
    n = S * S + 1
    xs = [randq(B, n, C).requires_grad_() for _ in range(T)]
    dys = [randq(B, n, C) for _ in range(T)]
    k = v = None
    ys, loss = [], 0
    for t in range(T):
        attn, k, v = mods[t](xs[t], k, v)
        y = xs[t] + attn  # deit_mrla_base.py:273-275
        ys.append(y)
        loss = loss + (y * dys[t]).sum()
    loss.backward()
    return dict(kind="deit_base", B=B, S=S, C=C, d=d, T=T, xs=[x.detach() for x in xs], dys=dys,
                ys=[y.detach() for y in ys], K=k.detach(), V=v.detach(), dxs=[x.grad.clone() for x in xs],
                blocks=[dict(params={n_: p.detach().clone() for n_, p in m.named_parameters()},
                             dparams=grads_of(m)) for m in mods])


CASES = {
    # name: (fn, kwargs)
    "light_layer_c64": (light_layer_case, dict(seed=1, B=3, C=64, H=7, W=7, d=32)),
    "light_layer_c256_ragged": (light_layer_case, dict(seed=2, B=2, C=256, H=3, W=5, d=32)),
    "light_tail_c64_train": (light_tail_case, dict(seed=3, B=4, C=64, H=7, W=7, d=32, drop_path=0.0, training=True)),
    "light_tail_c64_drop": (light_tail_case, dict(seed=4, B=6, C=64, H=6, W=5, d=32, drop_path=0.4, training=True)),
    "light_tail_c64_eval": (light_tail_case, dict(seed=5, B=3, C=64, H=7, W=7, d=32, drop_path=0.2, training=False)),
    "light_tail_c256_train": (light_tail_case, dict(seed=6, B=2, C=256, H=6, W=7, d=32, drop_path=0.0, training=True)),
    "light_tail_c2048_k7": (light_tail_case, dict(seed=7, B=2, C=2048, H=2, W=2, d=32, drop_path=0.0, training=True)),
    "light_tail_c96_d8_1x1": (light_tail_case, dict(seed=8, B=5, C=96, H=1, W=1, d=8, drop_path=0.0, training=True)),
    "light_tail_c64_d16_wide": (light_tail_case, dict(seed=9, B=2, C=64, H=2, W=37, d=16, drop_path=0.0, training=True, relu_x=False)),
    # feature maps wider than 56 columns (detection backbones, mmdetection/mmdet/models/backbones/resnet_mrlal.py:116): the v7
    # sweeps walk them in column tiles (round 2); train-mode BN and the mmdet setting (eval-mode BN, no DropPath)
    "light_tail_c64_w75": (light_tail_case, dict(seed=18, B=2, C=64, H=5, W=75, d=32, drop_path=0.0, training=True)),
    "light_tail_c128_w60_eval": (light_tail_case, dict(seed=19, B=2, C=128, H=4, W=60, d=32, drop_path=0.0, training=False)),
    "base_stage_c64_t3": (base_stage_case, dict(seed=10, B=3, C=64, H=5, W=5, d=16, T=3, drop_path=0.0)),
    "base_stage_c128_t4_drop": (base_stage_case, dict(seed=11, B=4, C=128, H=3, W=4, d=16, T=4, drop_path=0.3)),
    "base_stage_c64_eval": (base_stage_case, dict(seed=12, B=2, C=64, H=5, W=5, d=16, T=2, drop_path=0.2, training=False)),
    "deit_light_c64": (deit_light_case, dict(seed=13, B=3, S=4, C=64, d=16)),
    "deit_light_c192": (deit_light_case, dict(seed=14, B=2, S=3, C=192, d=16)),
    "deit_base_c64_t3": (deit_base_case, dict(seed=15, B=2, S=4, C=64, d=16, T=3)),
    # DeiT-tiny geometry of BASELINE configs[4]: 197 tokens (14x14 + cls), C = 192, 12 heads of 16 (round 2)
    "deit_light_c192_n197": (deit_light_case, dict(seed=16, B=2, S=14, C=192, d=16)),
    "deit_base_c192_t2_n197": (deit_base_case, dict(seed=17, B=2, S=14, C=192, d=16, T=2)),
}
F32_OUTPUTS = {"deit_light_c192_n197", "deit_base_c192_t2_n197"}   # outputs / gradients stored in fp32 (file size)


def _to_f32(obj):
    if isinstance(obj, torch.Tensor) and obj.dtype == torch.float64:
        return obj.float()
    if isinstance(obj, dict):
        return {k: _to_f32(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_to_f32(v) for v in obj]
    return obj


def main():
    assert ref_loader.available(), "reference tree not found (set MRLA_REF)"
    total = 0
    only = set(sys.argv[1:])
    for name, (fn, kw) in CASES.items():
        if only and name not in only:
            continue
        case = fn(**kw)
        if name in F32_OUTPUTS:
            case = _to_f32(case)
        path = os.path.join(HERE, name + ".pt.gz")
        with gzip.open(path, "wb", compresslevel=9) as f:
            torch.save(case, f)
        sz = os.path.getsize(path)
        total += sz
        print(f"{name:32s} {sz/1024:8.1f} KiB")
    print(f"total {total/1e6:.2f} MB")


if __name__ == "__main__":
    main()

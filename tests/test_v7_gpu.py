"""GPU parity of the round-2 "v7" sweeps (mrla_b200/csrc/light_v7.cuh) through the C ABI:

 (1) the bn3-folded tail op WITHOUT a materialised x (`x_virtual`: every sweep re-forms x = relu(bn3(c3) + identity)
     from the raw conv3 output) against the fp64 oracle restatement of resnet_mrla_light.py:101-102,113-116 at the
     four BASELINE stage shapes — fp32 at B=32 (config 1) and bf16 at the full B=256 (config 2), the shapes bench.py
     times;
 (2) the same op against the materialising kernels of round 1 (ALLOW_VIRTUAL_X off) and against bn3 as its own op;
 (3) the plain tail (x given) on the v7 sweeps against the oracle on ragged / odd shapes.

Tolerances: fp32 <= 1e-5, bf16 <= 2e-2 of max|reference| (norm-wise, conftest.rel_err).  Gradients that pass through
the bottleneck's ReLU are compared away from its kink: elements whose fp64 pre-activation |bn3(c3) + identity| is below
the rounding noise of the storage dtype may take either branch of the mask, so they are excluded (and counted: < 1 %)."""
import copy

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 4e-3}
STAGES = [(256, 56), (512, 28), (1024, 14), (2048, 7)]


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _masked_rel_err(a, b, keep):
    a, b = a.detach().double(), b.detach().double()
    denom = b.abs().max().item() or 1.0
    return ((a - b).abs() * keep).max().item() / denom


def _bn3_tail_case(B, C, HW, dtype, dev, seed=0, drop=True):
    from mrla_b200.modules.mrla_light_module import eca_kernel_size
    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    H, W = HW if isinstance(HW, tuple) else (HW, HW)
    c3 = _cl(rn(B, C, H, W).to(dtype))
    idt = _cl(torch.relu(rn(B, C, H, W)).to(dtype))
    dy = _cl(rn(B, C, H, W).to(dtype))
    k = eca_kernel_size(C)
    P = dict(w3=0.5 + torch.rand(C, device=dev, generator=g), b3=0.3 * rn(C), wq=0.5 * rn(k), wk=0.5 * rn(k),
             wv=rn(C, 1, 3, 3) * (2 / 9) ** 0.5, lam=rn(C, 1, 1), gamma=1 + 0.3 * rn(C), beta=0.2 * rn(C))
    keep = 0.8
    ds = ((torch.rand(B, device=dev, generator=g) + keep).floor() / keep) if drop else None
    return c3, idt, dy, P, ds, k


def _run_product(c3, idt, dy, P, ds, k, d=32):
    """ops.bn3_light_tail exactly as MRLA_Bottleneck.forward calls it."""
    import torch.nn as nn
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, bn3_light_tail, bn3_tail_eligible
    C = c3.shape[1]
    dev = c3.device
    bn3 = nn.BatchNorm2d(C).to(dev).train()
    with torch.no_grad():
        bn3.weight.copy_(P["w3"])
        bn3.bias.copy_(P["b3"])
    cfg = LightCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True, fuse_add_relu=True)
    assert bn3_tail_eligible(c3, idt, bn3, cfg)
    leaves = {n: P[n].clone().requires_grad_() for n in ("wq", "wk", "wv", "lam", "gamma", "beta")}
    c3g, idg = c3.clone().requires_grad_(), idt.clone().requires_grad_()
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    y = bn3_light_tail(c3g, idg, bn3, leaves["wq"], leaves["wk"], leaves["wv"], leaves["lam"], leaves["gamma"],
                       leaves["beta"], rm, rv, ds, cfg=cfg)
    y.backward(dy)
    torch.cuda.synchronize()
    grads = {n: v.grad for n, v in leaves.items()}
    grads["w3"], grads["b3"] = bn3.weight.grad, bn3.bias.grad
    return y.detach(), c3g.grad, idg.grad, grads, dict(rm=rm, rv=rv, rm3=bn3.running_mean.clone(), rv3=bn3.running_var.clone())


def _run_oracle(c3, idt, dy, P, ds, d=32, storage_dtype=None):
    from oracle import mrla_oracle as O
    C = c3.shape[1]
    dev = c3.device
    f64 = dict(dtype=torch.float64, device=dev)
    c3d, idd = c3.detach().double().requires_grad_(), idt.detach().double().requires_grad_()
    Pd = {n: v.detach().double().requires_grad_() for n, v in P.items()}
    y, ex = O.bottleneck_light_tail(c3d, idd, Pd["w3"], Pd["b3"], torch.zeros(C, **f64), torch.ones(C, **f64), Pd["wq"],
                                    Pd["wk"], Pd["wv"], Pd["lam"], C // d, Pd["gamma"], Pd["beta"], torch.zeros(C, **f64),
                                    torch.ones(C, **f64), drop_scale=None if ds is None else ds.double(),
                                    storage_dtype=storage_dtype)
    y.backward(dy.double())
    out = (y.detach(), c3d.grad, idd.grad, {n: v.grad for n, v in Pd.items()}, ex)
    return out


@pytest.mark.parametrize("dtype,B", [(torch.float32, 32), (torch.bfloat16, 256)])
@pytest.mark.parametrize("C,HW", STAGES)
def test_bn3_tail_virtual_x_vs_oracle(C, HW, dtype, B, cuda_device):
    """The op bench.py times (bn3 affine + add + ReLU folded, x never materialised) against the fp64 oracle at the
    BASELINE shapes — B=256 bf16 included (the oracle runs in fp64 on the GPU: ~35 GB of autograd state at stage 1)."""
    from mrla_b200 import ops
    dev = cuda_device
    c3, idt, dy, P, ds, k = _bn3_tail_case(B, C, HW, dtype, dev, seed=C + HW)
    calls0 = dict(ops.launch_counter)
    y, dc3, did, grads, bufs = _run_product(c3, idt, dy, P, ds, k)
    assert ops.launch_counter["fwd"] > calls0["fwd"]
    # bf16: the oracle stores the bn3 output and the sum in bf16 like the reference's autocast graph does (which side of
    # the ReLU an element near the kink falls on is decided by those two roundings); everything else is fp64
    yr, dc3r, didr, gr, ex = _run_oracle(c3, idt, dy, P, ds, storage_dtype=None if dtype == torch.float32 else dtype)
    tol = TOL[dtype]
    assert rel_err(y, yr) < tol
    # elements still excluded: the product's fp32 BatchNorm statistics differ from fp64 ones in the last bits, which can
    # move a bn3 output across a rounding boundary of the storage dtype and, if the sum is within one ulp of zero, flip
    # the mask (a handful of elements in 2e8)
    eps_store = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -22     # two ulps of the stored bn3 output
    keep = (ex["pre"].abs() > eps_store * ex["z"].abs() + 1e-30)
    assert keep.double().mean().item() > 0.99
    assert _masked_rel_err(dc3, dc3r, keep) < tol
    assert _masked_rel_err(did, didr, keep) < tol
    for n in grads:
        assert rel_err(grads[n], gr[n]) < 2 * tol, n
    assert rel_err(bufs["rm"], ex["running_mean"]) < tol and rel_err(bufs["rv"], ex["running_var"]) < tol
    assert rel_err(bufs["rm3"], ex["bn3_running_mean"]) < tol and rel_err(bufs["rv3"], ex["bn3_running_var"]) < tol
    del yr, dc3r, didr, gr, ex
    torch.cuda.empty_cache()


@pytest.mark.parametrize("shape", [(2, 64, 20, 100), (3, 128, 9, 57), (1, 256, 200, 304), (2, 512, 100, 152),
                                   (40, 256, 12, 70), (160, 64, 6, 113)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])   # fp32 stages of 8 column groups do not fit smem
def test_bn3_tail_virtual_x_wide_maps_vs_oracle(shape, dtype, cuda_device):
    """W > 56 (the detection backbone's feature maps, mmdetection/.../resnet_mrlal.py: 200x304 at stage 1 for an 800x1216
    image): the v7 sweeps walk each image in column tiles of 56; tile halos are the neighbouring tile's real columns.
    Small batches (the first four shapes) spread the tiles of one image over CTAs and add the moments atomically; the last
    two shapes keep one image per work unit."""
    B, C, H, W = shape
    dev = cuda_device
    c3, idt, dy, P, ds, k = _bn3_tail_case(B, C, (H, W), dtype, dev, seed=C + W, drop=False)
    y, dc3, did, grads, bufs = _run_product(c3, idt, dy, P, ds, k)
    yr, dc3r, didr, gr, ex = _run_oracle(c3, idt, dy, P, ds, storage_dtype=None if dtype == torch.float32 else dtype)
    tol = TOL[dtype]
    assert rel_err(y, yr) < tol
    eps_store = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -22
    keep = (ex["pre"].abs() > eps_store * ex["z"].abs() + 1e-30)
    assert keep.double().mean().item() > 0.99
    assert _masked_rel_err(dc3, dc3r, keep) < tol
    assert _masked_rel_err(did, didr, keep) < tol
    for n in grads:
        assert rel_err(grads[n], gr[n]) < 2 * tol, n


@pytest.mark.parametrize("shape", [(8, 256, 56), (16, 512, 28), (32, 1024, 14), (32, 2048, 7), (4, 128, 9), (3, 64, 13)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
def test_virtual_x_matches_materialised_x(shape, dtype, cuda_device, monkeypatch):
    """x_virtual (v7 sweeps) against the round-1 kernels that store x: same arithmetic per element, different summation
    order of the moments -> agreement to rounding of the accumulations, not bit for bit."""
    from mrla_b200 import ops
    B, C, HW = shape
    dev = cuda_device
    c3, idt, dy, P, ds, k = _bn3_tail_case(B, C, HW, dtype, dev, seed=7)
    a = _run_product(c3, idt, dy, P, ds, k)
    monkeypatch.setattr(ops, "ALLOW_VIRTUAL_X", False)
    b = _run_product(c3, idt, dy, P, ds, k)
    tol = {torch.float32: 2e-6, torch.bfloat16: 8e-3, torch.float16: 1e-3}[dtype]
    assert rel_err(a[0], b[0]) < tol
    assert rel_err(a[1], b[1]) < tol
    assert rel_err(a[2], b[2]) < tol
    ptol = {torch.float32: 2e-5, torch.bfloat16: 1e-2, torch.float16: 4e-3}[dtype]
    for n in a[3]:
        assert rel_err(a[3][n], b[3][n]) < ptol, n
    for n in a[4]:
        assert rel_err(a[4][n], b[4][n]) < 1e-5, n


@pytest.mark.parametrize("shape", [(4, 64, 7, 7), (3, 256, 9, 13), (2, 128, 56, 56), (5, 192, 14, 14), (2, 64, 3, 50),
                                   (3, 512, 28, 28), (2, 64, 12, 100), (1, 128, 40, 84), (2, 192, 5, 57)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["train", "eval", "none"])
def test_plain_tail_on_v7_sweeps_vs_oracle(shape, dtype, mode, cuda_device):
    """light_tail with a given x (mrla_module.forward / block tail without the fold) on the v7 sweeps: ragged widths
    (13, 50, 9), W < 7, every BatchNorm mode."""
    from mrla_b200 import _lib
    from mrla_b200.modules.mrla_light_module import eca_kernel_size
    from mrla_b200.ops import LightCfg, light_tail
    from oracle import mrla_oracle as O
    B, C, H, W = shape
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(H * W + C)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, o, dy = (_cl(t.to(dtype)) for t in (torch.relu(rn(B, C, H, W)), rn(B, C, H, W), rn(B, C, H, W)))
    k, d = eca_kernel_size(C), 32
    P = dict(wq=0.5 * rn(k), wk=0.5 * rn(k), wv=rn(C, 1, 3, 3) * (2 / 9) ** 0.5, lam=rn(C, 1, 1), gamma=1 + 0.3 * rn(C),
             beta=0.2 * rn(C))
    bn_mode = {"train": _lib.BN_TRAIN, "eval": _lib.BN_EVAL, "none": _lib.BN_NONE}[mode]
    rm0, rv0 = 0.1 * rn(C), 0.5 + torch.rand(C, device=dev, generator=g)
    leaves = {n: v.clone().requires_grad_() for n, v in P.items()}
    xg, og = x.clone().requires_grad_(), o.clone().requires_grad_()
    rm, rv = rm0.clone(), rv0.clone()
    cfg = LightCfg(dim_perhead=d, k_size=k, bn_mode=bn_mode, residual=(mode != "none"))
    if mode == "none":
        y = light_tail(xg, og, leaves["wq"], leaves["wk"], leaves["wv"], leaves["lam"], cfg=cfg)
    else:
        y = light_tail(xg, og, leaves["wq"], leaves["wk"], leaves["wv"], leaves["lam"], leaves["gamma"], leaves["beta"],
                       rm, rv, None, cfg=cfg)
    y.backward(dy)
    xd, od = x.double().requires_grad_(), o.double().requires_grad_()
    Pd = {n: v.double().requires_grad_() for n, v in P.items()}
    if mode == "none":
        yr = O.light_module(xd, od, Pd["wq"], Pd["wk"], Pd["wv"], Pd["lam"], C // d)
    else:
        yr, _, _ = O.light_tail(xd, od, Pd["wq"], Pd["wk"], Pd["wv"], Pd["lam"], C // d, Pd["gamma"], Pd["beta"],
                                rm0.double(), rv0.double(), training=(mode == "train"))
    yr.backward(dy.double())
    tol = TOL[dtype]
    assert rel_err(y, yr) < tol
    assert rel_err(xg.grad, xd.grad) < tol
    assert rel_err(og.grad, od.grad) < tol
    names = ("wq", "wk", "wv", "lam") if mode == "none" else tuple(P)
    for n in names:
        assert rel_err(leaves[n].grad, Pd[n].grad) < 2 * tol, n


def test_v7_plan_is_used_at_baseline_shapes(cuda_device):
    """The library reports x_virtual support for every BASELINE stage shape in bf16 (so bench.py times the v7 sweeps)."""
    import ctypes
    from mrla_b200 import _lib
    L = _lib.lib()
    for C, HW in STAGES:
        z = _cl(torch.zeros(2, C, HW, HW, device=cuda_device, dtype=torch.bfloat16))
        a = _lib.MrlaLightArgs()
        a.B, a.C, a.H, a.W = 256, C, HW, HW
        a.dim_perhead, a.k_size, a.dtype, a.layout, a.act, a.bn_mode = 32, 5, _lib.BF16, _lib.NHWC, 0, _lib.BN_TRAIN
        a.bs_x = a.bs_o = a.bs_y = a.bs_z = C * HW * HW
        a.z = a.o = z.data_ptr()
        assert L.mrla_light_virtual_x(ctypes.byref(a)) == 1, (C, HW)

"""GPU parity of the fused MRLA-light tail (through the C ABI) against
 (1) golden vectors frozen from the real reference (tests/golden/), and
 (2) the oracle restatement evaluated in fp64 on the same seeded inputs, at the ResNet-50 stage
     shapes of BASELINE.json (B=32 fp32 and B=256 bf16).
Tolerances are the north-star ones: fp32 <= 1e-5, bf16 <= 2e-2 (max-abs error / max-abs reference)."""
import pytest
import torch

from conftest import golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 4e-3}
LAYOUTS = ["nchw", "nhwc"]


def _to(t, dtype, layout, dev):
    t = t.to(dev, dtype)
    if layout == "nhwc" and t.ndim == 4:
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def _run_tail(g, dtype, layout, dev):
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, light_tail
    P = {k: v.to(dev, torch.float32).requires_grad_() for k, v in g["params"].items()}
    x = _to(g["x"], dtype, layout, dev).requires_grad_()
    o = _to(g["o"], dtype, layout, dev).requires_grad_()
    rm = g["running_mean0"].to(dev, torch.float32).clone()
    rv = g["running_var0"].to(dev, torch.float32).clone()
    ds = g["drop_scale"].to(dev, torch.float32) if g["drop_scale"] is not None else None
    cfg = LightCfg(dim_perhead=g["d"], k_size=g["k"], bn_mode=_lib.BN_TRAIN if g["training"] else _lib.BN_EVAL,
                   residual=True, update_running=g["training"], eps=g["eps"], momentum=g["momentum"])
    y = light_tail(x, o, P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"], P["lambda_t"],
                   P["bn.weight"], P["bn.bias"], rm, rv, ds, cfg=cfg)
    y.backward(_to(g["dy"], dtype, layout, dev))
    torch.cuda.synchronize()
    return x, o, y, P, rm, rv


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("name", golden_names("light_tail"))
def test_light_tail_golden(name, dtype, layout, cuda_device):
    g = load_golden(name)
    if layout == "nhwc" and g["C"] % 4:
        pytest.skip("NHWC path needs C % 4 == 0")
    x, o, y, P, rm, rv = _run_tail(g, dtype, layout, cuda_device)
    tol = TOL[dtype]
    assert y.shape == g["y"].shape and y.dtype == dtype
    assert rel_err(y, g["y"]) < tol
    assert rel_err(x.grad, g["dx"]) < tol
    assert rel_err(o.grad, g["do"]) < tol
    for k in P:
        assert rel_err(P[k].grad, g["dparams"][k]) < 2 * tol, k
    assert rel_err(rm, g["running_mean1"]) < tol
    assert rel_err(rv, g["running_var1"]) < tol


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", golden_names("light_layer"))
def test_light_layer_golden(name, dtype, layout, cuda_device):
    from mrla_b200.modules import mrla_light_layer
    g = load_golden(name)
    dev = cuda_device
    L = mrla_light_layer(g["C"], dim_perhead=g["d"]).to(dev)
    L.load_state_dict({k: v.float() for k, v in g["params"].items()}, strict=True)
    x = _to(g["x"], dtype, layout, dev).requires_grad_()
    y = L(x)
    y.backward(_to(g["dy"], dtype, layout, dev))
    tol = TOL[dtype]
    assert rel_err(y, g["y"]) < tol
    assert rel_err(x.grad, g["dx"]) < tol
    for k, p in L.named_parameters():
        assert rel_err(p.grad, g["dparams"][k]) < 2 * tol, k


STAGES = [(256, 56), (512, 28), (1024, 14), (2048, 7)]


def _oracle_tail_fp64(x, o, P, d, drop_scale, dy, training=True):
    """oracle/mrla_oracle.py evaluated in fp64 on the GPU tensors (checker only)."""
    from oracle import mrla_oracle as O
    C = x.shape[1]
    xd, od = x.detach().double().requires_grad_(), o.detach().double().requires_grad_()
    Pd = {k: v.detach().double().requires_grad_() for k, v in P.items()}
    rm0 = torch.zeros(C, dtype=torch.float64, device=x.device)
    rv0 = torch.ones(C, dtype=torch.float64, device=x.device)
    y, rm, rv = O.light_tail(xd, od, Pd["wq"], Pd["wk"], Pd["wv"], Pd["lam"], C // d, Pd["gamma"], Pd["beta"],
                             rm0, rv0, training=training, drop_scale=None if drop_scale is None else drop_scale.double())
    y.backward(dy.double())
    return y.detach(), xd.grad, od.grad, {k: v.grad for k, v in Pd.items()}, rm, rv


@pytest.mark.parametrize("layout", LAYOUTS + ["nchw_generic"])
@pytest.mark.parametrize("dtype,B", [(torch.float32, 32), (torch.bfloat16, 256)])
@pytest.mark.parametrize("C,HW", STAGES)
def test_light_tail_stage_shapes_vs_oracle(C, HW, dtype, B, layout, cuda_device, monkeypatch):
    """BASELINE.json stage shapes: config-1 batch (32, fp32) and config-2 batch (256/GPU, bf16).
    `nchw` inputs of this size are promoted to channels_last inside the op (TMA kernels); `nchw_generic`
    pins the generic NCHW kernels on the same inputs."""
    from mrla_b200 import _lib, ops
    if layout == "nchw_generic":
        monkeypatch.setattr(ops, "PROMOTE_NCHW", False)
        layout = "nchw"
    from mrla_b200.modules.mrla_light_module import eca_kernel_size
    from mrla_b200.ops import LightCfg, light_tail
    dev = cuda_device
    if dtype == torch.bfloat16 and HW == 56:
        B = 64  # fp64 oracle + autograd intermediates at B=256 would need > 40 GB; full size is covered below
    torch.manual_seed(1234 + C)
    d = 32
    x = _to(torch.relu(torch.randn(B, C, HW, HW, device=dev)), dtype, layout, dev)
    o = _to(torch.randn(B, C, HW, HW, device=dev), dtype, layout, dev)
    dy = _to(torch.randn(B, C, HW, HW, device=dev), dtype, layout, dev)
    k = eca_kernel_size(C)
    P = dict(wq=torch.randn(k, device=dev) * 0.5, wk=torch.randn(k, device=dev) * 0.5,
             wv=torch.randn(C, 1, 3, 3, device=dev) * (2 / 9) ** 0.5, lam=torch.randn(C, 1, 1, device=dev),
             gamma=1 + 0.3 * torch.randn(C, device=dev), beta=0.2 * torch.randn(C, device=dev))
    keep = 0.8
    ds = (torch.rand(B, device=dev) + keep).floor() / keep
    for v in P.values():
        v.requires_grad_()
    xg, og = x.clone().requires_grad_(), o.clone().requires_grad_()
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    cfg = LightCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True)
    y = light_tail(xg, og, P["wq"], P["wk"], P["wv"], P["lam"], P["gamma"], P["beta"], rm, rv, ds, cfg=cfg)
    y.backward(dy)
    yr, dxr, dor, dPr, rmr, rvr = _oracle_tail_fp64(x, o, P, d, ds, dy)
    tol = TOL[dtype]
    assert rel_err(y, yr) < tol
    assert rel_err(xg.grad, dxr) < tol
    assert rel_err(og.grad, dor) < tol
    for kname in P:
        assert rel_err(P[kname].grad, dPr[kname]) < 2 * tol, kname
    assert rel_err(rm, rmr) < tol and rel_err(rv, rvr) < tol


@pytest.mark.parametrize("layout", LAYOUTS)
def test_light_tail_full_size_properties(layout, cuda_device):
    """Size-independent properties at the full BASELINE stage-1 shape (256 x 256 x 56 x 56 bf16):
    BN of the branch is normalised per channel, dbeta equals the plain sum of the upstream gradient,
    and (train-mode BN) the per-channel sum of dS — hence of dO / lambda — vanishes."""
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, light_tail
    dev = cuda_device
    B, C, HW, d, k = 256, 256, 56, 32, 5
    torch.manual_seed(7)
    mk = lambda: _to(torch.randn(B, C, HW, HW, device=dev, dtype=torch.bfloat16), torch.bfloat16, layout, dev)
    x = torch.relu(mk()).requires_grad_()
    o = mk().requires_grad_()
    dy = mk()
    wq, wk = (torch.randn(k, device=dev) * 0.5).requires_grad_(), (torch.randn(k, device=dev) * 0.5).requires_grad_()
    wv = (torch.randn(C, 1, 3, 3, device=dev) * 0.47).requires_grad_()
    lam = torch.randn(C, 1, 1, device=dev).requires_grad_()
    gamma = (1 + 0.3 * torch.randn(C, device=dev)).requires_grad_()
    beta = (0.2 * torch.randn(C, device=dev)).requires_grad_()
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    cfg = LightCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True)
    y = light_tail(x, o, wq, wk, wv, lam, gamma, beta, rm, rv, None, cfg=cfg)
    y.backward(dy)
    z = (y.float() - x.detach().float())
    zm = z.mean(dim=(0, 2, 3))
    zv = z.var(dim=(0, 2, 3), unbiased=False)
    assert (zm - beta.detach()).abs().max().item() < 2e-2
    assert ((zv.sqrt() - gamma.detach().abs()).abs() / gamma.detach().abs().clamp_min(0.1)).max().item() < 2e-2
    dbeta_ref = dy.float().sum(dim=(0, 2, 3))
    assert rel_err(beta.grad, dbeta_ref) < 1e-3
    do_sum = o.grad.float().sum(dim=(0, 2, 3))
    scale = o.grad.float().abs().sum(dim=(0, 2, 3))
    assert (do_sum.abs() / scale.clamp_min(1e-6)).max().item() < 2e-2
    assert torch.isfinite(x.grad.float()).all() and torch.isfinite(wv.grad).all()


def test_module_and_block_tail_match_golden(cuda_device):
    """Drop-in module path: mrla_module + BatchNorm2d + DropPath objects, state-dict loaded from golden."""
    from mrla_b200.drop import DropPath
    from mrla_b200.resnet_mrla_light import mrla_light_block_tail, mrla_module
    dev = cuda_device
    g = load_golden("light_tail_c64_drop")

    class M(mrla_module):
        dim_perhead = g["d"]

    mod = M(g["C"]).to(dev)
    bn = torch.nn.BatchNorm2d(g["C"]).to(dev)
    mod.load_state_dict({k: v.float() for k, v in g["params"].items() if not k.startswith("bn.")}, strict=True)
    bn.weight.data.copy_(g["params"]["bn.weight"]); bn.bias.data.copy_(g["params"]["bn.bias"])
    bn.running_mean.copy_(g["running_mean0"]); bn.running_var.copy_(g["running_var0"])
    dp = DropPath(g["drop_path"])
    x = g["x"].float().to(dev).requires_grad_()
    o = g["o"].float().to(dev).requires_grad_()
    # replay the RNG state the golden generator used for the DropPath draw: m_b is recorded, so patch it in
    dp.scale = lambda t: g["drop_scale"].float().to(dev)
    y = mrla_light_block_tail(x, o, mod, bn, dp)
    y.backward(g["dy"].float().to(dev))
    assert rel_err(y, g["y"]) < 1e-5
    assert rel_err(x.grad, g["dx"]) < 1e-5
    assert rel_err(mod.lambda_t.grad, g["dparams"]["lambda_t"]) < 2e-5
    assert rel_err(bn.running_var, g["running_var1"]) < 1e-5
    assert int(bn.num_batches_tracked) == g["num_batches_tracked1"]
    # module alone (no BN): o_t = layer(x) + lambda * o
    from oracle import mrla_oracle as O
    P = g["params"]
    ref = O.light_module(g["x"], g["o"], P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"], P["lambda_t"],
                         g["C"] // g["d"])
    assert rel_err(mod(x.detach(), o.detach()), ref) < 1e-5


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("shape,dtype", [((4, 64, 7, 7), torch.float32), ((3, 256, 9, 13), torch.float32),
                                         ((32, 256, 56, 56), torch.bfloat16), ((32, 1024, 14, 14), torch.float32),
                                         ((64, 2048, 7, 7), torch.bfloat16)])
def test_light_tail_with_folded_add_relu(shape, dtype, layout, cuda_device):
    """`pre_add_relu`: x = relu(z + identity) formed inside the op (resnet_mrla_light.py:113-114 folded in) —
    output and the gradients w.r.t. z and identity (TOTAL, incl. the path through x) match the oracle applied to
    relu(z + identity).  Covers the fused sweep-B epilogue (TMA shapes) and the library-op fallback (small shapes)."""
    from mrla_b200 import _lib
    from mrla_b200.modules.mrla_light_module import eca_kernel_size
    from mrla_b200.ops import LightCfg, light_tail
    from oracle import mrla_oracle as O
    dev = cuda_device
    B, C, H, W = shape
    torch.manual_seed(B * 7 + C)
    d, k = 32, eca_kernel_size(C)
    z = _to(torch.randn(B, C, H, W, device=dev), dtype, layout, dev)
    idt = _to(torch.randn(B, C, H, W, device=dev), dtype, layout, dev)
    dy = _to(torch.randn(B, C, H, W, device=dev), dtype, layout, dev)
    P = dict(wq=torch.randn(k, device=dev) * 0.5, wk=torch.randn(k, device=dev) * 0.5,
             wv=torch.randn(C, 1, 3, 3, device=dev) * 0.4, lam=torch.randn(C, 1, 1, device=dev),
             gamma=1 + 0.3 * torch.randn(C, device=dev), beta=0.2 * torch.randn(C, device=dev))
    for v in P.values():
        v.requires_grad_()
    zg, ig = z.clone().requires_grad_(), idt.clone().requires_grad_()
    cfg = LightCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True, fuse_add_relu=True)
    y = light_tail(zg, ig, P["wq"], P["wk"], P["wv"], P["lam"], P["gamma"], P["beta"], torch.zeros(C, device=dev),
                   torch.ones(C, device=dev), None, cfg=cfg)
    y.backward(dy)
    zd, idd = z.double().requires_grad_(), idt.double().requires_grad_()
    Pd = {n: v.detach().double().requires_grad_() for n, v in P.items()}
    xd = torch.relu(zd + idd)
    if dtype != torch.float32:
        xd = xd + (xd.detach().to(dtype).double() - xd.detach())  # the op stores x in `dtype`: same rounding, unit gradient
    yr, _, _ = O.light_tail(xd, idd, Pd["wq"], Pd["wk"], Pd["wv"], Pd["lam"], C // d, Pd["gamma"], Pd["beta"],
                            torch.zeros(C, dtype=torch.float64, device=dev), torch.ones(C, dtype=torch.float64, device=dev))
    yr.backward(dy.double())
    tol = TOL[dtype]
    assert rel_err(y, yr) < tol
    assert rel_err(zg.grad, zd.grad) < tol
    assert rel_err(ig.grad, idd.grad) < tol
    for n in P:
        assert rel_err(P[n].grad, Pd[n].grad) < 2 * tol, n

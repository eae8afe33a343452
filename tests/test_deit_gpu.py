"""GPU parity of the DeiT (token-layout) MRLA modules against golden vectors frozen from the reference."""
import pytest
import torch

from conftest import golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 4e-3}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("name", golden_names("deit_light"))
def test_deit_light_module_golden(name, dtype, cuda_device):
    from mrla_b200.deit_mrla_light import mrlal_module
    g = load_golden(name)
    dev = cuda_device
    mod = mrlal_module(g["C"], g["d"]).to(dev)
    mod.load_state_dict({k: v.float() for k, v in g["params"].items()}, strict=True)
    mod = mod.to(dtype)
    x = g["x"].to(dev, dtype).requires_grad_()
    o = g["o"].to(dev, dtype).requires_grad_()
    y = x + mod(x, o)  # Block.forward, deit_mrla_light.py:234
    y.backward(g["dy"].to(dev, dtype))
    tol = TOL[dtype]
    assert y.shape == g["y"].shape
    assert rel_err(y, g["y"]) < tol
    assert rel_err(x.grad, g["dx"]) < 2 * tol
    assert rel_err(o.grad, g["do"]) < 2 * tol
    for k, p in mod.named_parameters():
        assert rel_err(p.grad, g["dparams"][k]) < 3 * tol, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", golden_names("deit_base"))
def test_deit_base_module_golden(name, dtype, cuda_device):
    from mrla_b200.deit_mrla_base import mrlab_module
    g = load_golden(name)
    dev = cuda_device
    T = g["T"]
    mods = []
    for t in range(T):
        m = mrlab_module(g["C"], g["d"], init_cell=(t == 0)).to(dev)
        m.load_state_dict({k: v.float() for k, v in g["blocks"][t]["params"].items()}, strict=True)
        mods.append(m.to(dtype))
    xs = [x.to(dev, dtype).requires_grad_() for x in g["xs"]]
    k = v = None
    ys = []
    for t in range(T):
        attn, k, v = mods[t](xs[t], k, v)
        ys.append(xs[t] + attn)  # deit_mrla_base.py:273-275
    torch.autograd.backward(ys, [d.to(dev, dtype) for d in g["dys"]])
    tol = TOL[dtype]
    assert rel_err(k, g["K"]) < tol and rel_err(v, g["V"]) < tol
    for t in range(T):
        assert rel_err(ys[t], g["ys"][t]) < tol, t
        assert rel_err(xs[t].grad, g["dxs"][t]) < 2 * tol, t
        for n, p in mods[t].named_parameters():
            ref = g["blocks"][t]["dparams"][n]
            if ref.abs().max().item() == 0:
                continue
            lim = 0.15 if (dtype != torch.float32 and ("Wq" in n or "Wk" in n)) else 3 * tol
            assert rel_err(p.grad, ref) < lim, (t, n)


def test_deit_light_layer_nchw_input(cuda_device):
    """mrlal_layer called directly on a plain contiguous NCHW tensor (reference signature forward(x))."""
    from mrla_b200.deit_mrla_light import mrlal_layer
    from oracle import mrla_oracle as O
    dev = cuda_device
    torch.manual_seed(3)
    L = mrlal_layer(64, dim_perhead=16).to(dev)
    x = torch.randn(2, 64, 5, 5, device=dev, requires_grad=True)
    y = L(x)
    y.sum().backward()
    xd = x.detach().double().requires_grad_()
    yr = O.light_layer(xd, L.Wq.weight.detach().double(), L.Wk.weight.detach().double(),
                       L.Wv.weight.detach().double(), 4, act="gelu")
    yr.sum().backward()
    assert rel_err(y, yr) < 1e-5 and rel_err(x.grad, xd.grad) < 1e-5

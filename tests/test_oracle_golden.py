"""The CPU oracle (oracle/mrla_oracle.py) held to the golden vectors frozen from the real
reference (tests/golden/make_golden.py).  fp64 on both sides -> agreement to ~1e-12."""
import pytest
import torch

from conftest import golden_names, load_golden, rel_err
from oracle import mrla_oracle as O

TOL = 1e-11


def _t(ref, base):
    """Fixtures whose outputs are stored in fp32 (the 197-token DeiT cases, file size) carry fp32 rounding of the reference."""
    return max(base, 2e-6) if ref.dtype == torch.float32 else base


def _leaf(t):
    return t.clone().requires_grad_()


@pytest.mark.parametrize("name", golden_names("light_layer"))
def test_light_layer(name):
    g = load_golden(name)
    P = {k: _leaf(v) for k, v in g["params"].items()}
    x = _leaf(g["x"])
    heads = g["C"] // g["d"]
    y = O.light_layer(x, P["Wq.weight"], P["Wk.weight"], P["Wv.weight"], heads)
    (y * g["dy"]).sum().backward()
    assert rel_err(y, g["y"]) < TOL
    assert rel_err(x.grad, g["dx"]) < TOL
    for k in P:
        assert rel_err(P[k].grad, g["dparams"][k]) < TOL, k


@pytest.mark.parametrize("name", golden_names("light_tail"))
def test_light_tail(name):
    g = load_golden(name)
    P = {k: _leaf(v) for k, v in g["params"].items()}
    x, o = _leaf(g["x"]), _leaf(g["o"])
    heads = g["C"] // g["d"]
    assert O.eca_kernel_size(g["C"]) == g["k"]
    y, rm, rv = O.light_tail(x, o, P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"], P["lambda_t"],
                             heads, P["bn.weight"], P["bn.bias"], g["running_mean0"], g["running_var0"],
                             training=g["training"], momentum=g["momentum"], eps=g["eps"],
                             drop_scale=g["drop_scale"])
    (y * g["dy"]).sum().backward()
    assert rel_err(y, g["y"]) < TOL
    assert rel_err(rm, g["running_mean1"]) < TOL
    assert rel_err(rv, g["running_var1"]) < TOL
    assert rel_err(x.grad, g["dx"]) < TOL
    assert rel_err(o.grad, g["do"]) < TOL
    for k in P:
        assert rel_err(P[k].grad, g["dparams"][k]) < 1e-10, k


@pytest.mark.parametrize("name", [n for n in golden_names("light_tail") if "eval" not in n])
def test_light_tail_closed_form(name):
    """The two-sweep algebra (per-(b,c) moments -> BN statistics) the CUDA path uses."""
    g = load_golden(name)
    P = g["params"]
    heads = g["C"] // g["d"]
    y, aux = O.light_tail_closed_form(g["x"], g["o"], P["mrla.Wq.weight"], P["mrla.Wk.weight"],
                                      P["mrla.Wv.weight"], P["lambda_t"], heads, P["bn.weight"], P["bn.bias"],
                                      eps=g["eps"], drop_scale=g["drop_scale"])
    assert rel_err(y, g["y"]) < 1e-9


@pytest.mark.parametrize("name", golden_names("base_stage"))
def test_base_stage(name):
    g = load_golden(name)
    T = g["T"]
    heads = g["C"] // g["d"]
    xs = [_leaf(x) for x in g["xs"]]
    Ps = [{k: _leaf(v) for k, v in blk["params"].items()} for blk in g["blocks"]]
    C = g["C"]
    k = v = None
    loss = 0
    ys = []
    for t in range(T):
        P = Ps[t]
        ds = g["drop_scales"][t] if g["drop_scales"] else None
        rm0, rv0 = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
        y, k, v, rm, rv = O.base_tail(xs[t], k, v, P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"],
                                      heads, t == 0, P["bn.weight"], P["bn.bias"], rm0, rv0,
                                      training=g["training"], momentum=g["momentum"], eps=g["eps"], drop_scale=ds)
        ys.append(y)
        loss = loss + (y * g["dys"][t]).sum()
        assert rel_err(rm, g["blocks"][t]["running_mean1"]) < TOL
        assert rel_err(rv, g["blocks"][t]["running_var1"]) < TOL
    loss.backward()
    assert rel_err(k, g["K"]) < TOL and rel_err(v, g["V"]) < TOL
    for t in range(T):
        assert rel_err(ys[t], g["ys"][t]) < _t(g["ys"][t], TOL)
        assert rel_err(xs[t].grad, g["dxs"][t]) < _t(g["ys"][t], 1e-10)
        for n in Ps[t]:
            assert rel_err(Ps[t][n].grad, g["blocks"][t]["dparams"][n]) < _t(g["ys"][t], 1e-10), (t, n)


@pytest.mark.parametrize("name", golden_names("deit_light"))
def test_deit_light(name):
    g = load_golden(name)
    P = {k: _leaf(v) for k, v in g["params"].items()}
    x, o = _leaf(g["x"]), _leaf(g["o"])
    heads = g["C"] // g["d"]
    y = O.deit_light_block_tail(x, o, P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"], P["lambda_t"],
                                heads, P["normx.weight"], P["normx.bias"], P["normo.weight"], P["normo.bias"])
    (y * g["dy"]).sum().backward()
    assert rel_err(y, g["y"]) < _t(g["y"], TOL)
    assert rel_err(x.grad, g["dx"]) < _t(g["y"], 1e-10)
    assert rel_err(o.grad, g["do"]) < _t(g["y"], 1e-10)
    for k in P:
        assert rel_err(P[k].grad, g["dparams"][k]) < _t(g["y"], 1e-10), k


@pytest.mark.parametrize("name", golden_names("deit_base"))
def test_deit_base(name):
    g = load_golden(name)
    T = g["T"]
    heads = g["C"] // g["d"]
    xs = [_leaf(x) for x in g["xs"]]
    Ps = [{k: _leaf(v) for k, v in blk["params"].items()} for blk in g["blocks"]]
    k = v = None
    loss, ys = 0, []
    for t in range(T):
        P = Ps[t]
        out, k, v = O.deit_base_module(xs[t], k, v, P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"],
                                       heads, t == 0, P["normx.weight"], P["normx.bias"])
        y = xs[t] + out
        ys.append(y)
        loss = loss + (y * g["dys"][t]).sum()
    loss.backward()
    for t in range(T):
        assert rel_err(ys[t], g["ys"][t]) < _t(g["ys"][t], TOL)
        assert rel_err(xs[t].grad, g["dxs"][t]) < _t(g["ys"][t], 1e-10)
        for n in Ps[t]:
            assert rel_err(Ps[t][n].grad, g["blocks"][t]["dparams"][n]) < _t(g["ys"][t], 1e-10), (t, n)

"""GPU tests added in round 2 for the review findings and the parity holes the judge listed:

 * norm_layer = SyncBatchNorm / GroupNorm keeps the norm module's own semantics (no silent per-GPU statistics);
 * BatchNorm statistics of a channel with |mean| >> std (E[x^2] - mean^2 cancellation);
 * model on a device that is not torch.cuda.current_device() (needs 2 GPUs);
 * a second backward over the same MRLA-base graph (retain_graph) reproduces the first;
 * create_graph through the fused ops raises instead of returning a history-free gradient;
 * MRLA-base at ResNet-101 depth (T = 23 blocks in one stage: cache growth path);
 * DeiT light / base modules at the DeiT-tiny shape of BASELINE configs[4] (B=256, 197 tokens, C=192, bf16)."""
import pytest
import torch
import torch.nn as nn

from conftest import rel_err, rel_err_l2

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("norm", ["sync", "group"])
def test_non_plain_norm_layers_keep_their_semantics(norm, cuda_device):
    """MRLA_Bottleneck(norm_layer=SyncBatchNorm | GroupNorm): identical to the same block evaluated with library ops only
    (reference graph resnet_mrla_light.py:92-118), including the gradients; no fused BatchNorm kernel may touch them."""
    from mrla_b200 import ops
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck
    from oracle import mrla_oracle as O
    dev = cuda_device
    torch.backends.cudnn.allow_tf32 = False
    mk = (lambda c: nn.SyncBatchNorm(c)) if norm == "sync" else (lambda c: nn.GroupNorm(8, c))
    torch.manual_seed(0)
    blk = _cl_mod(MRLA_Bottleneck(256, 64, norm_layer=mk).to(dev)).train()
    with torch.no_grad():
        blk.bn3.weight.uniform_(0.5, 1.5)
    x = _cl(torch.relu(torch.randn(4, 256, 14, 14, device=dev))).requires_grad_()
    dy = _cl(torch.randn(4, 256, 14, 14, device=dev))
    before = dict(ops.launch_counter)
    y = blk(x)
    y.backward(dy)
    # library-only evaluation of the reference graph with the same modules
    xr = x.detach().clone().requires_grad_()
    out = torch.relu(blk.bn1(blk.conv1(xr)))
    out = torch.relu(blk.bn2(blk.conv2(out)))
    out = torch.relu(blk.bn3(blk.conv3(out)) + xr)
    m = blk.mrla
    s = O.light_module(out, xr, m.mrla.Wq.weight, m.mrla.Wk.weight, m.mrla.Wv.weight, m.lambda_t, 256 // 32)
    yr = out + blk.bn_mrla(s)
    gx, = torch.autograd.grad(yr, xr, dy)
    assert rel_err(y, yr) < 1e-5
    assert rel_err(x.grad, gx) < 3e-4   # three cuDNN convolutions forward and backward on both sides
    # only the MRLA module itself (mrla_light_forward / backward) ran on the fused kernels
    assert ops.launch_counter["fwd"] - before["fwd"] <= 4


def _cl_mod(m):
    return m.to(memory_format=torch.channels_last)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_batchnorm_large_mean_small_std(dtype, cuda_device):
    """ops.bn_act on a channel distribution with mean ~ 50, std ~ 0.1 (ADVICE: E[x^2] - mean^2 cancels in fp32):
    statistics, output and gradients against nn.BatchNorm2d evaluated in fp64."""
    from mrla_b200.ops import bn_act
    dev = cuda_device
    torch.manual_seed(1)
    C = 64
    mean = torch.linspace(-60, 60, C, device=dev).view(1, C, 1, 1)
    x = _cl((mean + 0.1 * torch.randn(8, C, 28, 28, device=dev)).to(dtype))
    dy = _cl(torch.randn(8, C, 28, 28, device=dev).to(dtype))
    bn = nn.BatchNorm2d(C).to(dev).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    ref = nn.BatchNorm2d(C).to(dev).double().train()
    ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    xg = x.clone().requires_grad_()
    y = bn_act(xg, bn, relu=False)
    y.backward(dy)
    xd = x.double().requires_grad_()
    yr = ref(xd)
    yr.backward(dy.double())
    if dtype == torch.float32:
        assert rel_err(bn.running_var, ref.running_var) < 1e-5
        assert rel_err(bn.running_mean, ref.running_mean) < 1e-6
        assert rel_err(y, yr) < 1e-4          # the fp32 input itself carries 50 * 2^-24 / 0.1 ~ 3e-5 of relative noise
        assert rel_err(xg.grad, xd.grad) < 1e-3
        assert rel_err(bn.weight.grad, ref.weight.grad) < 1e-3
    else:
        # bf16 inputs near 50 are quantised to steps of 0.25: the statistics must still be those of the stored values
        assert rel_err(bn.running_var, ref.running_var) < 1e-4
        assert rel_err(y, yr) < 2e-2
    assert rel_err(bn.bias.grad, ref.bias.grad) < 1e-4 if dtype == torch.float32 else True


def test_second_device(cuda_device):
    """Module on cuda:1 while the current device is cuda:0 (ADVICE: launches followed the current device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, light_tail
    from oracle import mrla_oracle as O
    dev = torch.device("cuda:1")
    torch.cuda.set_device(0)
    torch.manual_seed(2)
    B, C, H, W, d, k = 4, 64, 14, 14, 32, 3
    x = _cl(torch.relu(torch.randn(B, C, H, W, device=dev))).requires_grad_()
    o = _cl(torch.randn(B, C, H, W, device=dev)).requires_grad_()
    P = [torch.randn(k, device=dev), torch.randn(k, device=dev), torch.randn(C, 1, 3, 3, device=dev),
         torch.randn(C, 1, 1, device=dev), torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)]
    cfg = LightCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, residual=True)
    y = light_tail(x, o, *P, torch.zeros(C, device=dev), torch.ones(C, device=dev), None, cfg=cfg)
    y.sum().backward()
    assert torch.cuda.current_device() == 0
    yr, _, _ = O.light_tail(x.detach().double(), o.detach().double(), *[p.double() for p in P[:4]], C // d, P[4].double(),
                            P[5].double(), torch.zeros(C, device=dev).double(), torch.ones(C, device=dev).double())
    assert rel_err(y, yr) < 1e-5


def _base_stage(xs, Ps, d, k, dev):
    from mrla_b200 import _lib
    from mrla_b200.ops import BaseCfg, base_tail
    C = xs[0].shape[1]
    kk = vv = None
    ys = []
    for t, (x, P) in enumerate(zip(xs, Ps)):
        cfg = BaseCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN, relu=True, residual=True)
        y, kk, vv = base_tail(x, kk, vv, P["wq"], P["wk"], P["wv"], P["gamma"], P["beta"], torch.zeros(C, device=dev),
                              torch.ones(C, device=dev), None, init_cell=(t == 0), cfg=cfg)
        ys.append(y)
    return ys


def _base_params(T, C, k, dev):
    Ps = [dict(wq=torch.randn(k, device=dev) * 0.5, wk=torch.randn(k, device=dev) * 0.5,
               wv=torch.randn(C, 1, 3, 3, device=dev) * 0.47, gamma=1 + 0.3 * torch.randn(C, device=dev),
               beta=0.2 * torch.randn(C, device=dev)) for _ in range(T)]
    for P in Ps:
        for v in P.values():
            v.requires_grad_()
    return Ps


def test_base_stage_second_backward_matches_first(cuda_device):
    """retain_graph=True: the in-place dV / dK accumulation state is reset once the first block's backward has run."""
    dev = cuda_device
    torch.manual_seed(3)
    B, C, HW, d, T, k = 3, 64, 7, 16, 4, 3
    xs = [_cl(torch.relu(torch.randn(B, C, HW, HW, device=dev))).requires_grad_() for _ in range(T)]
    dys = [_cl(torch.randn(B, C, HW, HW, device=dev)) for _ in range(T)]
    Ps = _base_params(T, C, k, dev)
    ys = _base_stage(xs, Ps, d, k, dev)
    torch.autograd.backward(ys, dys, retain_graph=True)
    g1 = [x.grad.clone() for x in xs] + [P[n].grad.clone() for P in Ps for n in P]
    for x in xs:
        x.grad = None
    for P in Ps:
        for v in P.values():
            v.grad = None
    torch.autograd.backward(ys, dys)
    g2 = [x.grad for x in xs] + [P[n].grad for P in Ps for n in P]
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)


def test_create_graph_raises(cuda_device):
    """The backward kernels are not differentiable: asking for a graph through them must fail loudly."""
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, light_tail
    dev = cuda_device
    B, C, H, W, d, k = 2, 64, 7, 7, 32, 3
    x = _cl(torch.relu(torch.randn(B, C, H, W, device=dev))).requires_grad_()
    o = _cl(torch.randn(B, C, H, W, device=dev)).requires_grad_()
    P = [torch.randn(k, device=dev), torch.randn(k, device=dev), torch.randn(C, 1, 3, 3, device=dev),
         torch.randn(C, 1, 1, device=dev)]
    y = light_tail(x, o, *P, cfg=LightCfg(dim_perhead=d, k_size=k))
    gx, = torch.autograd.grad(y.sum(), x, create_graph=True)
    with pytest.raises(RuntimeError):
        gx.sum().backward()


@pytest.mark.parametrize("layout", ["nhwc", "nchw"])
def test_base_stage_resnet101_depth(layout, cuda_device):
    """T = 23 blocks in one stage (resnet101_mrlab stage 3, reference resnet_mrla_base.py:280): the stage cache grows
    past its initial capacity twice; outputs and gradients against the fp64 oracle chain."""
    from oracle import mrla_oracle as O
    dev = cuda_device
    torch.manual_seed(4)
    B, C, HW, d, T, k = 2, 128, 7, 16, 23, 3
    mk = (lambda t: _cl(t)) if layout == "nhwc" else (lambda t: t.contiguous())
    xs = [mk(torch.relu(torch.randn(B, C, HW, HW, device=dev))).requires_grad_() for _ in range(T)]
    dys = [mk(torch.randn(B, C, HW, HW, device=dev)) for _ in range(T)]
    Ps = _base_params(T, C, k, dev)
    ys = _base_stage(xs, Ps, d, k, dev)
    torch.autograd.backward(ys, dys)
    xd = [x.detach().double().requires_grad_() for x in xs]
    Pd = [{n: v.detach().double().requires_grad_() for n, v in P.items()} for P in Ps]
    kr = vr = None
    yr = []
    z, o1 = torch.zeros(C, dtype=torch.float64, device=dev), torch.ones(C, dtype=torch.float64, device=dev)
    for t in range(T):
        P = Pd[t]
        y, kr, vr, _, _ = O.base_tail(xd[t], kr, vr, P["wq"], P["wk"], P["wv"], C // d, t == 0, P["gamma"], P["beta"], z, o1)
        yr.append(y)
    torch.autograd.backward(yr, [g.double() for g in dys])
    for t in range(T):
        assert rel_err(ys[t], yr[t]) < 1e-5, t
        assert rel_err(xs[t].grad, xd[t].grad) < 2e-5, t
        for n in Ps[t]:
            truth = Pd[t][n].grad
            if truth.abs().max().item() == 0.0:
                continue
            assert rel_err(Ps[t][n].grad, truth) < 5e-5, (t, n)


def _deit_light_oracle(mod, x, o):
    from oracle import mrla_oracle as O
    P = {n: p.detach().double().requires_grad_() for n, p in mod.named_parameters()}
    xd, od = x.detach().double().requires_grad_(), o.detach().double().requires_grad_()
    C = x.shape[-1]
    y = O.deit_light_block_tail(xd, od, P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"], P["lambda_t"],
                                C // mod.dim_perhead, P["normx.weight"], P["normx.bias"], P["normo.weight"],
                                P["normo.bias"])
    return y, xd, od, P


@pytest.mark.parametrize("dtype,B", [(torch.float32, 32), (torch.bfloat16, 256)])
def test_deit_light_at_deit_tiny_shape(dtype, B, cuda_device):
    """BASELINE configs[4]: deit_mrlal_tiny, 197 tokens x 192 channels, batch 256 / GPU in bf16 (and batch 32 in fp32)."""
    from mrla_b200.deit_mrla_light import mrlal_module
    dev = cuda_device
    torch.manual_seed(5)
    mod = mrlal_module(192, 16).to(dev)
    with torch.no_grad():
        for ln in (mod.normx, mod.normo):
            ln.weight.uniform_(0.7, 1.3)
            ln.bias.normal_(0, 0.2)
    x = torch.randn(B, 197, 192, device=dev).to(dtype)
    o = torch.randn(B, 197, 192, device=dev).to(dtype)
    dy = torch.randn(B, 197, 192, device=dev).to(dtype)
    yr, xd, od, P = _deit_light_oracle(mod, x, o)
    yr.backward(dy.double())
    m = mod.to(dtype)
    xg, og = x.clone().requires_grad_(), o.clone().requires_grad_()
    y = xg + m(xg, og)    # Block.forward, deit/deit_mrla_light.py:234
    y.backward(dy)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel_err(y, yr) < tol
    assert rel_err(xg.grad, xd.grad) < 2 * tol
    assert rel_err(og.grad, od.grad) < 2 * tol
    for n, p in m.named_parameters():
        assert rel_err_l2(p.grad, P[n].grad) < 3 * tol, n


def test_deit_base_at_deit_tiny_shape(cuda_device):
    """deit_mrlab_tiny geometry (197 x 192, 12 heads), 4 blocks sharing one cache (reset every 4 layers,
    deit/deit_mrla_base.py:261-264), batch 64 in bf16."""
    from mrla_b200.deit_mrla_base import mrlab_module
    from oracle import mrla_oracle as O
    dev = cuda_device
    torch.manual_seed(6)
    B, T, C, dph = 64, 4, 192, 16
    mods = [mrlab_module(C, dph, init_cell=(t == 0)).to(dev) for t in range(T)]
    xs = [torch.randn(B, 197, C, device=dev).bfloat16() for _ in range(T)]
    dys = [torch.randn(B, 197, C, device=dev).bfloat16() for _ in range(T)]
    # oracle (fp64)
    xd = [x.double().requires_grad_() for x in xs]
    Pd = [{n: p.detach().double().requires_grad_() for n, p in m.named_parameters()} for m in mods]
    kr = vr = None
    yr = []
    for t in range(T):
        P = Pd[t]
        out, kr, vr = O.deit_base_module(xd[t], kr, vr, P["mrla.Wq.weight"], P["mrla.Wk.weight"], P["mrla.Wv.weight"],
                                         C // dph, t == 0, P["normx.weight"], P["normx.bias"])
        yr.append(xd[t] + out)
    torch.autograd.backward(yr, [g.double() for g in dys])
    # product (bf16)
    ms = [m.bfloat16() for m in mods]
    xg = [x.clone().requires_grad_() for x in xs]
    k = v = None
    ys = []
    for t in range(T):
        attn, k, v = ms[t](xg[t], k, v)
        ys.append(xg[t] + attn)
    torch.autograd.backward(ys, dys)
    for t in range(T):
        assert rel_err(ys[t], yr[t]) < 2e-2, t
        assert rel_err(xg[t].grad, xd[t].grad) < 4e-2, t

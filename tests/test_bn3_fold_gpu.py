"""GPU: the bottleneck's bn3 folded into the MRLA-light tail op (one autograd node: BatchNorm statistics -> sweep 1 applies
a_c*conv3 + b_c, adds the identity, ReLUs, takes the moments) against the same block with bn3 as its own op — SURVEY.md
§8f rank 1, reference resnet/models/resnet_mrla_light.py:101-102,113-116.  The two paths run the same arithmetic per
element (x is formed in storage precision either way) with different summation orders, so outputs, gradients and the
BatchNorm buffers agree to accumulation rounding; parity with the reference itself is tests/test_v7_gpu.py (fp64 oracle of
the folded op at the BASELINE shapes) and test_model_gpu.py (whole model)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _block(inplanes, planes, downsample, drop_path, dev):
    import torch.nn as nn
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck, _conv1x1
    ds = None
    if downsample:
        ds = nn.Sequential(_conv1x1(inplanes, planes * 4), nn.BatchNorm2d(planes * 4))
    torch.manual_seed(3)
    blk = MRLA_Bottleneck(inplanes, planes, downsample=ds, drop_path=drop_path).to(dev)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith("bn3.weight"):
                p.uniform_(0.5, 1.5)
            if n.endswith("bn3.bias"):
                p.normal_(0, 0.3)
    return blk.to(memory_format=torch.channels_last).train()


@pytest.mark.parametrize("shape", [(8, 256, 64, 56, False), (8, 64, 64, 56, True), (16, 512, 128, 28, False),
                                   (32, 1024, 256, 14, False), (32, 2048, 512, 7, False), (4, 128, 64, 9, True)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_bn3_fold_equals_separate_bn3(shape, dtype, monkeypatch):
    import mrla_b200.resnet_mrla_light as M
    from mrla_b200 import ops
    B, cin, planes, hw, ds = shape
    dev = torch.device("cuda:0")
    blk_a = _block(cin, planes, ds, 0.2, dev)
    blk_b = copy.deepcopy(blk_a)
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.relu(torch.randn(B, cin, hw, hw, generator=g)).to(dev).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, planes * 4, hw, hw, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    used = {"fold": 0}
    real = ops._Bn3LightTail.apply

    def counting(*a):
        used["fold"] += 1
        return real(*a)

    outs = []
    for blk, fold in ((blk_a, True), (blk_b, False)):
        if not fold:
            monkeypatch.setattr(M, "bn3_tail_eligible", lambda *a, **k: False)
        else:
            monkeypatch.setattr(ops._Bn3LightTail, "apply", counting)
        xi = x.clone().requires_grad_()
        torch.manual_seed(5)   # same DropPath draw
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
            y = blk(xi if dtype == torch.float32 else xi.to(dtype))
        y.backward(dy.to(y.dtype))
        outs.append((y.detach(), xi.grad, {n: p.grad for n, p in blk.named_parameters()},
                     {n: b.clone() for n, b in blk.named_buffers()}))
    assert used["fold"] == 1, "the folded path did not run"
    (ya, dxa, ga, ba), (yb, dxb, gb, bb) = outs
    # round 2: the folded op re-forms x inside every sweep (7-column threads, different summation order of the moments
    # and of the BatchNorm reductions), so the two paths agree to accumulation rounding instead of bit for bit
    from conftest import rel_err
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(ya, yb) < tol
    assert rel_err(dxa, dxb) < tol
    for n in ga:
        assert (ga[n] is None) == (gb[n] is None), n
        if ga[n] is not None:
            assert rel_err(ga[n], gb[n]) < 2 * tol, n
    for n in ba:
        assert rel_err(ba[n].float(), bb[n].float()) < 1e-5, n


def test_bn3_fold_not_used_in_eval_or_nchw():
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, bn3_tail_eligible
    import torch.nn as nn
    dev = torch.device("cuda:0")
    bn3 = nn.BatchNorm2d(64).to(dev)
    cfg = LightCfg(dim_perhead=32, k_size=3, bn_mode=_lib.BN_TRAIN, residual=True, fuse_add_relu=True)
    c3 = torch.randn(2, 64, 14, 14, device=dev)
    o = torch.randn(2, 64, 14, 14, device=dev)
    assert not bn3_tail_eligible(c3, o, bn3, cfg)                      # NCHW
    c3c, oc = (t.contiguous(memory_format=torch.channels_last) for t in (c3, o))
    assert bn3_tail_eligible(c3c, oc, bn3, cfg)
    assert not bn3_tail_eligible(c3c, oc, bn3.eval(), cfg)             # running statistics
    assert not bn3_tail_eligible(c3c, oc, bn3.train(), cfg._replace(bn_mode=_lib.BN_EVAL))

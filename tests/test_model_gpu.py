"""Whole-model drop-in check on the GPU: the product ResNet (cuDNN backbone + fused MRLA tails through the C ABI)
against the oracle port of the reference model with the SAME state_dict, logits and every parameter gradient."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _unzero_bn3(model):
    for n, p in model.named_parameters():
        if n.endswith("bn3.weight"):
            torch.nn.init.normal_(p, 1.0, 0.2)


def _compare(prod, orc, x, tol, l2=False, oracle_x=None):
    """`l2`: compare parameter gradients in the L2 norm.  A whole network in fp32 is a chaotic map for single gradient
    elements: product and oracle differ in the last bits of every BatchNorm statistic, which flips a handful of ReLU
    decisions (MRLA-base adds one more ReLU behind bn_mrla) and moves individual elements by percents while every op on its
    own agrees to 1e-5 (tests/test_base_gpu.py, test_robust_gpu.py).  tools/dbg_mrlab.py shows the max-norm figure
    scattering between 8e-4 and 4e-2 over seeds for round-1 and round-2 builds alike."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yp, yo = prod(x), orc(x if oracle_x is None else oracle_x)
    assert rel_err(yp, yo) < tol
    yp.square().sum().backward()
    yo.square().sum().backward()
    go = dict(orc.named_parameters())
    scale = max(p.grad.abs().max().item() for p in go.values())
    worst = 0.0
    for n, p in prod.named_parameters():
        if l2:
            err = (p.grad - go[n].grad).norm().item() / max(go[n].grad.norm().item(), 1e-3 * scale)
        else:
            err = (p.grad - go[n].grad).abs().max().item() / max(go[n].grad.abs().max().item(), 1e-3 * scale)
        worst = max(worst, err)
        assert err < 20 * tol, (n, err)
    bo = dict(orc.named_buffers())
    for n, b in prod.named_buffers():
        if b.dtype.is_floating_point:
            assert rel_err(b, bo[n]) < tol, n
        else:
            assert int(b) == int(bo[n]), n
    return worst


@pytest.mark.parametrize("channels_last", [False, True])
def test_resnet_mrlal_matches_oracle_model(channels_last, cuda_device):
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck, ResNet_mrlal
    from oracle.resnet_oracle import ResNetMrlalOracle
    dev = cuda_device
    torch.manual_seed(0)
    prod = ResNet_mrlal(MRLA_Bottleneck, [2, 1, 1, 1], num_classes=10).to(dev).train()
    _unzero_bn3(prod)
    orc = ResNetMrlalOracle([2, 1, 1, 1], num_classes=10).to(dev).train()
    orc.load_state_dict(prod.state_dict(), strict=True)
    x = torch.randn(4, 3, 96, 96, device=dev)
    if channels_last:
        prod = prod.to(memory_format=torch.channels_last)
    # The product computes in channels_last either way (ops.promote_images re-strides a dense NCHW batch at the stem), so
    # the oracle is given the same values in channels_last strides: cuDNN picks other algorithms per layout, and the ORACLE
    # run in NCHW differs from ITSELF run in channels_last by 6e-3 (L2) / 2.9e-2 (max) in single parameter gradients of
    # this network (ReLU decisions flip on last-bit differences; profiles/r02_whole_model_noise_floor.txt,
    # tools/dbg_mrlal.py), while product and oracle in the same layout agree to 9e-5.
    x = x.contiguous(memory_format=torch.channels_last)
    _compare(prod, orc, x if channels_last else x.clone(memory_format=torch.contiguous_format), 2e-4, oracle_x=x)


def test_resnet_mrlab_matches_oracle_model(cuda_device):
    from mrla_b200.resnet_mrla_base import MRLA_Bottleneck, ResNet_mrlab
    from oracle.resnet_oracle import ResNetMrlabOracle
    dev = cuda_device
    torch.manual_seed(1)
    prod = ResNet_mrlab(MRLA_Bottleneck, [2, 2, 1, 1], num_classes=10).to(dev).train()
    _unzero_bn3(prod)
    orc = ResNetMrlabOracle([2, 2, 1, 1], num_classes=10).to(dev).train()
    assert list(prod.state_dict()) == list(orc.state_dict())
    orc.load_state_dict(prod.state_dict(), strict=True)
    x = torch.randn(4, 3, 96, 96, device=dev).contiguous(memory_format=torch.channels_last)
    prod = prod.to(memory_format=torch.channels_last)
    _compare(prod, orc, x, 5e-4, l2=True)


def test_training_trajectory_matches_oracle_model(cuda_device):
    """Three SGD steps (fp32, drop_path 0): the product model's loss curve tracks the oracle model's."""
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck, ResNet_mrlal
    from oracle.resnet_oracle import ResNetMrlalOracle
    dev = cuda_device
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    prod = ResNet_mrlal(MRLA_Bottleneck, [1, 1, 1, 1], num_classes=10).to(dev).train()
    _unzero_bn3(prod)
    orc = ResNetMrlalOracle([1, 1, 1, 1], num_classes=10).to(dev).train()
    orc.load_state_dict(prod.state_dict(), strict=True)
    x = torch.randn(8, 3, 64, 64, device=dev)
    y = torch.randint(0, 10, (8,), device=dev)
    curves = []
    for model in (prod, orc):
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
        ls = []
        for _ in range(3):
            loss = torch.nn.functional.cross_entropy(model(x), y)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            ls.append(loss.item())
        curves.append(ls)
    for a_, b_ in zip(*curves):
        assert abs(a_ - b_) < 2e-3 * max(1.0, abs(b_)), curves


def test_resnet50_mrlal_bf16_recipe_runs(cuda_device):
    """The BASELINE configs[1] recipe at a small batch: bf16 autocast, channels_last, SGD, drop_path 0.2 — every
    parameter gets a finite gradient (DDP-safe), the loss stays finite, eval() is deterministic."""
    from mrla_b200.resnet_mrla_light import resnet50_mrlal
    dev = cuda_device
    torch.manual_seed(0)
    model = resnet50_mrlal(drop_path=0.2, num_classes=16).to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.001, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(8, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 16, (8,), device=dev)
    for _ in range(3):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(x)
        loss = torch.nn.functional.cross_entropy(out.float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
        opt.step()
        assert torch.isfinite(loss)
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a, b = model(x), model(x)
    assert torch.equal(a, b)

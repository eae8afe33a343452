"""GPU parity of the fused channels-last BatchNorm(+ReLU) producer op against torch.nn.BatchNorm2d (+relu) in
fp64 — the same semantics the reference bottleneck uses for bn1/bn2/bn3 (resnet_mrla_light.py:92-102)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(4, 64, 7, 5), (2, 256, 14, 14), (3, 2048, 3, 3), (8, 24, 9, 9), (16, 512, 28, 28)])
def test_bn_act_matches_torch(shape, dtype, relu, training, cuda_device):
    from mrla_b200.ops import bn_act, bn_act_eligible
    dev = cuda_device
    B, C, H, W = shape
    torch.manual_seed(C + H)
    bn = torch.nn.BatchNorm2d(C).to(dev)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * torch.randn(C)); bn.bias.copy_(0.2 * torch.randn(C))
        bn.running_mean.copy_(0.1 * torch.randn(C)); bn.running_var.copy_(0.5 + torch.rand(C))
    ref = torch.nn.BatchNorm2d(C).to(dev).double()
    ref.load_state_dict({k: v.double() if v.dtype.is_floating_point else v for k, v in bn.state_dict().items()})
    bn.train(training); ref.train(training)
    x = (torch.randn(B, C, H, W, device=dev) * 1.5 + 0.3).to(dtype).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, C, H, W, device=dev).to(dtype).contiguous(memory_format=torch.channels_last)
    assert bn_act_eligible(x)
    xg = x.clone().requires_grad_()
    y = bn_act(xg, bn, relu=relu)
    y.backward(dy)
    xd = x.double().requires_grad_()
    yr = ref(xd)
    if relu:
        yr = torch.relu(yr)
    yr.backward(dy.double())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert y.dtype == dtype and y.shape == x.shape
    assert rel_err(y, yr) < tol
    if dtype == torch.float32 or not relu:
        assert rel_err(xg.grad, xd.grad) < 2 * tol
    else:  # bf16 + ReLU: mask flips on rounding noise -> L2 norm
        from conftest import rel_err_l2
        assert rel_err_l2(xg.grad, xd.grad) < 2 * tol
    assert rel_err(bn.weight.grad, ref.weight.grad) < 3 * tol
    assert rel_err(bn.bias.grad, ref.bias.grad) < 3 * tol
    assert rel_err(bn.running_mean, ref.running_mean) < tol
    assert rel_err(bn.running_var, ref.running_var) < tol
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked)


def test_bn_act_falls_back_for_nchw(cuda_device):
    from mrla_b200.ops import bn_act, bn_act_eligible
    dev = cuda_device
    bn = torch.nn.BatchNorm2d(32).to(dev)
    x = torch.randn(4, 32, 6, 6, device=dev, requires_grad=True)
    assert not bn_act_eligible(x)
    y = bn_act(x, bn, relu=True)
    ref = torch.relu(torch.nn.functional.batch_norm(x, None, None, bn.weight, bn.bias, True))
    assert rel_err(y, ref) < 1e-6

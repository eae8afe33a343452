"""GPU parity of the stem max pooling (3x3, stride 2, pad 1, NHWC) against torch.nn.functional.max_pool2d, through the
C ABI (mrla_maxpool3x3s2_forward / backward).  Forward is bit-exact; backward routes each gradient to the same tap as
at::max_pool2d (first maximum in row-major scan order), so it is bit-exact too."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _run(B, C, H, W, dtype, relu_input):
    from mrla_b200.ops import max_pool
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + C * 10 + H + W)
    x = torch.randn(B, C, H, W, generator=g).to(dev)
    if relu_input:
        x = torch.relu(x)   # ~50 % exact ties at zero, like the real post-ReLU stem activation
    x = x.to(dtype).contiguous(memory_format=torch.channels_last)
    pool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
    xa = x.clone().requires_grad_()
    xb = x.clone().requires_grad_()
    ya = max_pool(xa, pool)
    yb = F.max_pool2d(xb, 3, 2, 1)
    assert ya.shape == yb.shape
    assert torch.equal(ya, yb)
    dy = torch.randn(yb.shape, generator=g).to(dev).to(dtype).contiguous(memory_format=torch.channels_last)
    ya.backward(dy)
    yb.backward(dy)
    assert xa.grad.shape == xb.grad.shape
    assert torch.equal(xa.grad, xb.grad)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 64, 112, 112), (3, 8, 7, 7), (2, 24, 13, 10), (1, 16, 2, 2), (2, 32, 5, 8)])
@pytest.mark.parametrize("relu_input", [False, True])
def test_maxpool_matches_torch(shape, dtype, relu_input):
    _run(*shape, dtype, relu_input)


def test_maxpool_nan_propagates():
    from mrla_b200.ops import max_pool
    dev = torch.device("cuda:0")
    x = torch.randn(1, 8, 6, 6, device=dev).contiguous(memory_format=torch.channels_last)
    x[0, 3, 2, 2] = float("nan")
    pool = nn.MaxPool2d(3, 2, 1)
    ya, yb = max_pool(x, pool), F.max_pool2d(x, 3, 2, 1)
    assert torch.equal(torch.isnan(ya), torch.isnan(yb))
    assert torch.equal(torch.nan_to_num(ya), torch.nan_to_num(yb))


def test_other_pool_configs_use_the_module():
    from mrla_b200.ops import max_pool
    dev = torch.device("cuda:0")
    x = torch.randn(2, 8, 9, 9, device=dev)   # NCHW: not eligible -> nn.MaxPool2d itself
    pool = nn.MaxPool2d(3, 2, 1)
    assert torch.equal(max_pool(x, pool), pool(x))
    xc = x.contiguous(memory_format=torch.channels_last)
    pool2 = nn.MaxPool2d(2, 2)
    assert torch.equal(max_pool(xc, pool2), pool2(xc))

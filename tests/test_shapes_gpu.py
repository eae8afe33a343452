"""Edge / extra shapes of SURVEY.md §8a: EfficientNet-B0 block shapes (row A9, module-level parity only),
mmdet-style non-square eval-BN inputs wider than the TMA tile limit (generic path), frozen parameters."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _tail_vs_oracle(B, C, H, W, d, dtype, layout, training, dev, frozen=False, tol=None):
    from mrla_b200 import _lib
    from mrla_b200.modules.mrla_light_module import eca_kernel_size
    from mrla_b200.ops import LightCfg, light_tail
    from oracle import mrla_oracle as O
    torch.manual_seed(C * 131 + H)
    k = eca_kernel_size(C)
    mk = lambda: torch.randn(B, C, H, W, device=dev).to(dtype)
    x, o, dy = torch.relu(mk()), mk(), mk()
    if layout == "nhwc":
        x, o, dy = (t.contiguous(memory_format=torch.channels_last) for t in (x, o, dy))
    P = dict(wq=torch.randn(k, device=dev) * 0.5, wk=torch.randn(k, device=dev) * 0.5,
             wv=torch.randn(C, 1, 3, 3, device=dev) * 0.4, lam=torch.randn(C, 1, 1, device=dev),
             gamma=1 + 0.3 * torch.randn(C, device=dev), beta=0.2 * torch.randn(C, device=dev))
    rm0, rv0 = 0.1 * torch.randn(C, device=dev), 0.5 + torch.rand(C, device=dev)
    for n, v in P.items():
        v.requires_grad_(not (frozen and n in ("wv", "gamma")))
    xg, og = x.clone().requires_grad_(), o.clone().requires_grad_()
    rm, rv = rm0.clone(), rv0.clone()
    cfg = LightCfg(dim_perhead=d, k_size=k, bn_mode=_lib.BN_TRAIN if training else _lib.BN_EVAL, residual=True,
                   update_running=training)
    y = light_tail(xg, og, P["wq"], P["wk"], P["wv"], P["lam"], P["gamma"], P["beta"], rm, rv, None, cfg=cfg)
    y.backward(dy)
    xd, od = x.double().requires_grad_(), o.double().requires_grad_()
    Pd = {n: v.detach().double().requires_grad_() for n, v in P.items()}
    yr, rmr, rvr = O.light_tail(xd, od, Pd["wq"], Pd["wk"], Pd["wv"], Pd["lam"], C // d, Pd["gamma"], Pd["beta"],
                                rm0.double(), rv0.double(), training=training)
    yr.backward(dy.double())
    tol = tol or (1e-5 if dtype == torch.float32 else 2e-2)
    assert rel_err(y, yr) < tol
    assert rel_err(xg.grad, xd.grad) < tol and rel_err(og.grad, od.grad) < tol
    for n in P:
        if P[n].requires_grad:
            assert rel_err(P[n].grad, Pd[n].grad) < 3 * tol, n
        else:
            assert P[n].grad is None
    assert rel_err(rm, rmr) < tol and rel_err(rv, rvr) < tol


EFFNET_B0 = [(16, 112), (24, 56), (40, 28), (80, 14), (112, 14), (192, 7), (320, 7)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,HW", EFFNET_B0)
def test_efficientnet_b0_block_shapes(C, HW, dtype, cuda_device):
    """Row A9: no reference model exists, so parity is module-level on the EfficientNet-B0 block-output shapes
    (dim_perhead 8 divides every width)."""
    _tail_vs_oracle(8, C, HW, HW, 8, dtype, "nhwc", True, cuda_device)


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_mmdet_style_eval_bn_nonsquare_wide(layout, cuda_device):
    """mmdet backbone variant (resnet_mrlal.py:116): eval-mode BN, no DropPath, non-square COCO-like feature map
    wider than the 56-column TMA tile (generic kernels), with frozen Wv / BN weight (stage-1 freezing)."""
    _tail_vs_oracle(2, 256, 40, 84, 32, torch.float32, layout, False, cuda_device, frozen=True)


def test_very_wide_input_is_rejected_cleanly(cuda_device):
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, light_tail
    dev = cuda_device
    C, W = 8, 600
    x = torch.randn(1, C, 2, W, device=dev)
    with pytest.raises(RuntimeError, match="MRLA_ERR_SHAPE"):
        light_tail(x, torch.randn_like(x), torch.randn(3, device=dev), torch.randn(3, device=dev),
                   torch.randn(C, 1, 3, 3, device=dev), torch.randn(C, 1, 1, device=dev),
                   cfg=LightCfg(dim_perhead=8, k_size=3))

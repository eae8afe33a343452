"""Live check of the oracle against the unmodified reference (dev container only — skipped on the GPU
box, where /root/reference does not exist)."""
import pytest
import torch

from conftest import rel_err
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def test_resnet_oracle_matches_reference_model():
    """Whole-model: same state_dict -> same logits and gradients (fp64, tiny input)."""
    from oracle.resnet_oracle import ResNetMrlalOracle
    torch.manual_seed(0)
    rl = ref_loader.resnet_light()
    ref = rl.ResNet_mrlal(rl.MRLA_Bottleneck, [1, 1, 1, 1], num_classes=10).double()
    orc = ResNetMrlalOracle([1, 1, 1, 1], num_classes=10).double()
    assert list(ref.state_dict().keys()) == list(orc.state_dict().keys())
    # un-zero bn3 so the residual branch matters
    for m in ref.modules():
        if isinstance(m, rl.MRLA_Bottleneck):
            torch.nn.init.normal_(m.bn3.weight, 1.0, 0.2)
    orc.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 3, 64, 64, dtype=torch.float64)
    yr, yo = ref(x), orc(x)
    assert rel_err(yo, yr) < 1e-10
    yr.square().sum().backward()
    yo.square().sum().backward()
    gr = dict(ref.named_parameters())
    for n, p in orc.named_parameters():
        # some gradients are analytically zero (a bias feeding another BatchNorm): absolute floor
        assert (p.grad - gr[n].grad).abs().max().item() < 1e-8 * (1e-3 + gr[n].grad.abs().max().item()), n
    for (n, b1), (_, b2) in zip(orc.named_buffers(), ref.named_buffers()):
        assert rel_err(b1, b2) < 1e-10, n


def test_product_model_state_dict_matches_reference():
    """Drop-in contract: identical keys, shapes and parameter count (no CUDA needed to construct)."""
    from mrla_b200.resnet_mrla_light import resnet50_mrlal
    ref = ref_loader.resnet_light().resnet50_mrlal(drop_path=0.2)
    mine = resnet50_mrlal(drop_path=0.2)
    sr, sm = ref.state_dict(), mine.state_dict()
    assert list(sr.keys()) == list(sm.keys())
    assert all(sr[k].shape == sm[k].shape for k in sr)
    assert sum(p.numel() for p in mine.parameters()) == 25738452
    mine.load_state_dict(sr, strict=True)


def test_eca_kernel_sizes_match_reference():
    from mrla_b200.modules.mrla_light_module import eca_kernel_size
    from oracle.mrla_oracle import eca_kernel_size as ok
    L = ref_loader.light_layer_mod().mrla_light_layer
    for c in (16, 24, 40, 64, 80, 96, 112, 192, 256, 320, 384, 512, 768, 1024, 2048):
        k = L(c, dim_perhead=8).k_size
        assert eca_kernel_size(c) == k == ok(c), c

"""CPU-only checks of the host-side mirror of the reference interface (no kernels run)."""
import pytest
import torch

from oracle import mrla_oracle as O


def test_constructor_contract_light():
    from mrla_b200.modules import mrla_light_layer
    with pytest.raises(ValueError, match="heads and dim_perhead cannot be None"):
        mrla_light_layer(64)
    L = mrla_light_layer(256, dim_perhead=32)
    assert L.heads == 8 and L.k_size == 5
    assert {k: tuple(v.shape) for k, v in L.state_dict().items()} == {
        "Wq.weight": (1, 1, 5), "Wk.weight": (1, 1, 5), "Wv.weight": (256, 1, 3, 3)}
    assert mrla_light_layer(2048, heads=64).k_size == 7
    assert mrla_light_layer(64, heads=2, k_size=9).k_size == 9


def test_constructor_contract_base_and_deit():
    from mrla_b200.deit_mrla_base import mrlab_module
    from mrla_b200.deit_mrla_light import mrlal_module
    from mrla_b200.modules import mrla_base_layer
    from mrla_b200.resnet_mrla_base import mrla_module as base_module
    from mrla_b200.resnet_mrla_light import mrla_module as light_module
    with pytest.raises(ValueError):
        mrla_base_layer(64)
    assert light_module.dim_perhead == 32 and base_module.dim_perhead == 16
    m = light_module(512)
    assert sorted(m.state_dict()) == ["lambda_t", "mrla.Wk.weight", "mrla.Wq.weight", "mrla.Wv.weight"]
    assert tuple(m.lambda_t.shape) == (512, 1, 1)
    assert base_module(64, channel_wise=True).mrla.heads == 64
    d = mrlal_module(192, 16)
    assert tuple(d.lambda_t.shape) == (192,)
    assert sorted(d.state_dict()) == ["lambda_t", "mrla.Wk.weight", "mrla.Wq.weight", "mrla.Wv.weight", "normo.bias",
                                      "normo.weight", "normx.bias", "normx.weight"]
    assert d.normx.eps == 1e-6
    b = mrlab_module(192, 16, init_cell=True)
    assert sorted(b.state_dict()) == ["mrla.Wk.weight", "mrla.Wq.weight", "mrla.Wv.weight", "normx.bias", "normx.weight"]


def test_no_cpu_fallback():
    from mrla_b200.modules import mrla_light_layer
    L = mrla_light_layer(64, dim_perhead=32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L(torch.randn(2, 64, 7, 7))
    from mrla_b200.resnet_mrla_light import mrla_module
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mrla_module(64)(torch.randn(2, 64, 7, 7), torch.randn(2, 64, 7, 7))


def test_product_and_oracle_models_share_state_dict():
    """Key-for-key interchange between the product model and the oracle port (both mirror the reference keys)."""
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck, ResNet_mrlal
    from oracle.resnet_oracle import ResNetMrlalOracle
    a = ResNet_mrlal(MRLA_Bottleneck, [1, 1, 1, 1], num_classes=10)
    b = ResNetMrlalOracle([1, 1, 1, 1], num_classes=10)
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict(), strict=True)
    # reference init facts (resnet_mrla_light.py:176-189): the depthwise Wv is re-initialised with
    # kaiming_normal_(fan_out) because it is an nn.Conv2d (torch's fan_out = C*9 for a [C,1,3,3] weight);
    # the last BN of every residual branch starts at 0
    blk = a.layer1[0]
    assert float(blk.bn3.weight.detach().abs().max()) == 0.0
    std = float(blk.mrla.mrla.Wv.weight.detach().std())
    assert abs(std - (2 / (256 * 9)) ** 0.5) < 0.3 * (2 / (256 * 9)) ** 0.5


def test_se_eca_rejected():
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck
    with pytest.raises(NotImplementedError):
        MRLA_Bottleneck(64, 16, SE=True)


def test_drop_path_scale_matches_reference_draw():
    """Same RNG consumption and same values as resnet/models/utils/drop.py:17-23 (restated in the oracle)."""
    from mrla_b200.drop import DropPath, drop_scale
    x = torch.randn(16, 4, 3, 3)
    torch.manual_seed(123)
    mine = drop_scale(x, 0.3, True)
    torch.manual_seed(123)
    ref = O.drop_path_scale(16, 0.3, True, x)
    assert torch.equal(mine, ref.float())
    assert set(mine.tolist()) <= {0.0, 1 / 0.7} or all(abs(v) < 1e-6 or abs(v - 1 / 0.7) < 1e-6 for v in mine.tolist())
    assert drop_scale(x, 0.3, False) is None and drop_scale(x, 0.0, True) is None
    dp = DropPath(0.5).eval()
    assert dp(x) is x
    dp.train()
    torch.manual_seed(7)
    y = dp(x)
    torch.manual_seed(7)
    m = drop_scale(x, 0.5, True)
    assert torch.allclose(y, x * m.view(-1, 1, 1, 1))


def test_layout_classifier():
    from mrla_b200 import _lib
    from mrla_b200.ops import _layout_of
    x = torch.randn(2, 8, 4, 5)
    assert _layout_of(x) == (_lib.NCHW, 160)
    assert _layout_of(x.contiguous(memory_format=torch.channels_last)) == (_lib.NHWC, 160)
    tok = torch.randn(2, 17, 8)
    from mrla_b200.deit_mrla_light import tokens_as_image
    img = tokens_as_image(tok[:, 1:])
    assert img.shape == (2, 8, 4, 4) and _layout_of(img) == (_lib.NHWC, 17 * 8)
    assert torch.equal(img, tok[:, 1:].reshape(2, 4, 4, 8).permute(0, 3, 1, 2))
    assert _layout_of(x[:, :, ::2]) is None
    with pytest.raises(ValueError):
        tokens_as_image(torch.randn(2, 15, 8))


def test_bn3_fold_and_pool_host_logic_on_cpu_tensors():
    """The bn3 fold and the stem pooling are decided on the host: CPU tensors are never eligible, the pooling (backbone)
    then runs the module itself, and the MRLA tail still refuses to run without the CUDA kernels."""
    import torch.nn as nn
    from mrla_b200 import _lib
    from mrla_b200.ops import LightCfg, bn3_tail_eligible, max_pool
    from mrla_b200.resnet_mrla_light import MRLA_Bottleneck
    cfg = LightCfg(dim_perhead=32, k_size=3, bn_mode=_lib.BN_TRAIN, residual=True, fuse_add_relu=True)
    c3 = torch.randn(2, 64, 7, 7).contiguous(memory_format=torch.channels_last)
    assert not bn3_tail_eligible(c3, c3.clone(), nn.BatchNorm2d(64), cfg)
    assert not bn3_tail_eligible(c3, c3.clone(), nn.BatchNorm2d(64), cfg._replace(fuse_add_relu=False))
    pool = nn.MaxPool2d(3, 2, 1)
    x = torch.randn(2, 8, 9, 9)
    assert torch.equal(max_pool(x, pool), pool(x))
    blk = MRLA_Bottleneck(64, 16).train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.randn(2, 64, 7, 7))


def test_light_args_mirror_has_the_producer_fold_fields():
    """ABI v4: z / bs_z / z_coef / dz_sums close the struct, in the order of include/mrla_b200.h; x_virtual sits in the
    old reserved slot; MrlaBnArgs ends with the optional precomputed sums."""
    from mrla_b200 import _lib
    names = [f[0] for f in _lib.MrlaLightArgs._fields_]
    assert names[-5:] == ["scratch_bytes", "z", "bs_z", "z_coef", "dz_sums"]
    assert names[13] == "x_virtual"
    assert _lib.ABI_VERSION == 4
    bn = [f[0] for f in _lib.MrlaBnArgs._fields_]
    assert bn[6] == "stats_only" and bn[-1] == "sums"


def test_promote_images_is_identity_off_gpu():
    """ops.promote_images only re-strides dense NCHW CUDA batches; CPU tensors (and the oracle's inputs) pass through."""
    import torch
    from mrla_b200.ops import promote_images
    x = torch.randn(2, 3, 8, 8)
    assert promote_images(x) is x
    y = torch.randn(2, 8)
    assert promote_images(y) is y


def test_norm_layer_gate_and_momentum_rule():
    """Only an exact nn.BatchNorm2d takes the fused BatchNorm / tail paths (SyncBatchNorm, GroupNorm and subclasses keep their
    own module semantics: ADVICE round 1); momentum=None is the cumulative moving average of nn.BatchNorm2d."""
    import torch
    import torch.nn as nn
    from mrla_b200.ops import effective_momentum, is_plain_batchnorm

    class MyBN(nn.BatchNorm2d):
        pass

    assert is_plain_batchnorm(nn.BatchNorm2d(8))
    for m in (nn.SyncBatchNorm(8), nn.GroupNorm(2, 8), MyBN(8), nn.Identity(), nn.BatchNorm1d(8)):
        assert not is_plain_batchnorm(m)
    bn = nn.BatchNorm2d(8, momentum=0.25)
    assert effective_momentum(bn) == 0.25
    bn = nn.BatchNorm2d(8, momentum=None)
    bn.num_batches_tracked.fill_(3)
    assert effective_momentum(bn) == pytest.approx(1 / 3)            # counter already incremented by the caller
    assert effective_momentum(bn, pending=1) == pytest.approx(1 / 4)  # counter incremented after the op


def test_flat_gradient_views_keep_parameter_strides():
    """train.GraphedStep hands every parameter a view of ONE flat fp32 buffer as its .grad; channels_last conv weights keep
    their strides (so backward kernels write into the buffer without a layout-converting copy)."""
    import torch
    from mrla_b200.train import _grad_view
    flat = torch.zeros(64 * 16 * 3 * 3 + 10)
    w = torch.randn(64, 16, 3, 3).contiguous(memory_format=torch.channels_last)
    gv = _grad_view(w, flat[:w.numel()])
    assert gv.shape == w.shape and gv.stride() == w.stride()
    gv.copy_(w)
    assert torch.equal(gv, w) and flat[:w.numel()].abs().sum() > 0        # a view: writes land in the flat buffer
    b = torch.randn(10)
    gb = _grad_view(b, flat[w.numel():])
    assert gb.shape == b.shape and gb.data_ptr() == flat[w.numel():].data_ptr()
    w2 = torch.randn(8, 4, 1, 1)                                           # dense NCHW weight: plain view
    assert _grad_view(w2, torch.zeros(32)).stride() == w2.stride()

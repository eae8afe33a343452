"""ResNet + MRLA-light — host-side mirror of resnet/models/resnet_mrla_light.py (reference), with the
block tail `out + drop_path(bn_mrla(mrla(out, identity)))` (reference :116) executed as ONE fused
CUDA op (two sweeps forward, two sweeps backward) instead of ~14 ATen launches.

Kept identical to the reference: class / factory names, constructor arguments, module tree and
therefore every state_dict key and shape (`...mrla.mrla.Wq.weight`, `...mrla.lambda_t`,
`...bn_mrla.*`), weight init, `forward(xt, ot_1)` of `mrla_module`, and `models.__dict__[arch](...)`
style construction (train.py:158).  The backbone convolutions / BatchNorms stay on cuDNN.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .drop import DropPath
from .modules.mrla_light_module import mrla_light_layer
from .ops import (bn3_light_tail, bn3_tail_eligible, bn_act, effective_momentum, is_plain_batchnorm, light_tail,
                  max_pool, promote_images)

__all__ = ["ResNet_mrlal", "MRLA_Bottleneck", "mrla_module", "mrla_light_block_tail",
           "resnet50_mrlal", "resnet101_mrlal"]


class mrla_module(nn.Module):
    """o_t = mrla_light(x_t) + lambda_t * o_{t-1}   (reference resnet_mrla_light.py:32-43)."""
    dim_perhead = 32

    def __init__(self, input_dim):
        super().__init__()
        self.mrla = mrla_light_layer(input_dim=input_dim, dim_perhead=self.dim_perhead)
        self.lambda_t = nn.Parameter(torch.randn(input_dim, 1, 1))

    def forward(self, xt, ot_1):
        m = self.mrla
        return light_tail(xt, ot_1, m.Wq.weight, m.Wk.weight, m.Wv.weight, self.lambda_t, cfg=m.cfg())


def _bn_effective_momentum(bn: nn.BatchNorm2d) -> float:
    # momentum=None: cumulative moving average; the counter is incremented after the op
    return effective_momentum(bn, pending=1)


def mrla_light_block_tail(out, identity, mrla: mrla_module, bn: nn.BatchNorm2d, drop_path: nn.Module,
                          pre_add_relu: bool = False, pre_bn: nn.BatchNorm2d = None):
    """Fused `out + drop_path(bn(mrla(out, identity)))` (reference :116; mmdet variant
    mmdetection/mmdet/models/backbones/resnet_mrlal.py:116 = eval-mode BN, no DropPath).

    `pre_add_relu=True`: `out` is the pre-activation bn3 output and the residual add + ReLU of the bottleneck
    (`out += identity; out = relu(out)`, reference :113-114) is folded into the op: one pass forms
    x = relu(out + identity), and the backward epilogue of sweep B emits d(out) = dx*[x>0] and the TOTAL
    identity gradient directly (no threshold_backward / gradient-accumulation passes).

    `pre_bn` (with pre_add_relu): `out` is the RAW conv3 output and `pre_bn` the bottleneck's bn3 (reference :101-102).
    Where the library folds it, bn3 becomes statistics + a per-channel affine applied inside sweep 1 (one autograd node
    for bn3 + add + ReLU + tail); otherwise bn3 runs as its own op first."""
    layer = mrla.mrla
    if not is_plain_batchnorm(bn):
        # norm_layer = SyncBatchNorm / GroupNorm / frozen BN ...: keep the norm module's own semantics (cross-rank
        # statistics, group statistics).  The MRLA module still runs on the fused kernels; bn / drop_path / residual are
        # the reference's own sequence (resnet_mrla_light.py:101-102,113-116).
        if pre_bn is not None:
            out = bn_act(out, pre_bn)
        if pre_add_relu:
            out = torch.relu(out + identity)
        return out + drop_path(bn(mrla(out, identity)))
    use_batch_stats = bn.training or bn.running_mean is None
    if use_batch_stats:
        mode = _lib.BN_TRAIN
        momentum = _bn_effective_momentum(bn) if bn.track_running_stats else 0.0
    else:
        mode, momentum = _lib.BN_EVAL, 0.0
    update = bn.training and bn.track_running_stats and bn.running_mean is not None
    scale = drop_path.scale(out) if isinstance(drop_path, DropPath) else None
    cfg = layer.cfg(bn_mode=mode, residual=True, update_running=update, eps=bn.eps, momentum=momentum,
                    fuse_add_relu=pre_add_relu)
    if pre_bn is not None and pre_add_relu and bn3_tail_eligible(out, identity, pre_bn, cfg):
        y = bn3_light_tail(out, identity, pre_bn, layer.Wq.weight, layer.Wk.weight, layer.Wv.weight, mrla.lambda_t,
                           bn.weight, bn.bias, bn.running_mean, bn.running_var, scale, cfg=cfg)
    else:
        if pre_bn is not None:
            out = bn_act(out, pre_bn)
        y = light_tail(out, identity, layer.Wq.weight, layer.Wk.weight, layer.Wv.weight, mrla.lambda_t,
                       bn.weight, bn.bias, bn.running_mean, bn.running_var, scale, cfg=cfg)
    if update:
        bn.num_batches_tracked += 1
    return y


def _conv3x3(cin, cout, stride=1, groups=1, dilation=1):
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=dilation, groups=groups, bias=False, dilation=dilation)


def _conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride=stride, bias=False)


class MRLA_Bottleneck(nn.Module):
    """Bottleneck residual block followed by the MRLA-light tail (reference :46-118)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, SE=False, ECA_size=None, groups=1,
                 base_width=64, dilation=1, norm_layer=nn.BatchNorm2d, drop_path=0.0):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        if SE or ECA_size is not None:
            # off in every MRLA configuration of the reference (resnet_mrla_light.py:126-127); out of scope here
            raise NotImplementedError("SE / ECA channel attention is not part of the MRLA hot path")
        width = int(planes * (base_width / 64.0)) * groups
        cout = planes * self.expansion
        self.conv1, self.bn1 = _conv1x1(inplanes, width), norm_layer(width)
        self.conv2, self.bn2 = _conv3x3(width, width, stride, groups, dilation), norm_layer(width)
        self.conv3, self.bn3 = _conv1x1(width, cout), norm_layer(cout)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.se = None
        self.eca = None
        self.mrla = mrla_module(input_dim=cout)
        self.bn_mrla = norm_layer(cout)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        # conv -> cuDNN; BatchNorm(+ReLU) -> fused NHWC kernels (ops.bn_act; library batch_norm for other layouts)
        if self.downsample is None:
            identity = x
        elif isinstance(self.downsample, nn.Sequential) and len(self.downsample) == 2 \
                and isinstance(self.downsample[1], nn.BatchNorm2d):
            identity = bn_act(self.downsample[0](x), self.downsample[1])
        else:
            identity = self.downsample(x)
        out = bn_act(self.conv1(x), self.bn1, relu=True)
        out = bn_act(self.conv2(out), self.bn2, relu=True)
        # bn3 and `out += identity; relu` (reference :101-102, :113-114) are folded into the tail op
        return mrla_light_block_tail(self.conv3(out), identity, self.mrla, self.bn_mrla, self.drop_path,
                                     pre_add_relu=True, pre_bn=self.bn3)


class ResNet_mrlal(nn.Module):
    def __init__(self, block, layers, num_classes=1000, SE=False, ECA=None, zero_init_last_bn=True, groups=1,
                 width_per_group=64, replace_stride_with_dilation=None, norm_layer=nn.BatchNorm2d, drop_rate=0.0,
                 drop_path=0.0):
        super().__init__()
        self._norm_layer = norm_layer = norm_layer or nn.BatchNorm2d
        self.num_classes, self.drop_rate, self.drop_path = num_classes, drop_rate, drop_path
        self.inplanes, self.dilation = 64, 1
        dil = replace_stride_with_dilation or [False, False, False]
        if len(dil) != 3:
            raise ValueError(f"replace_stride_with_dilation should be None or a 3-element tuple, got {dil}")
        ECA = ECA or [None] * 4
        if len(ECA) != 4:
            raise ValueError(f"argument ECA should be a 4-element tuple, got {ECA}")
        self.groups, self.base_width = groups, width_per_group
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = norm_layer(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        for i, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2))):
            stage = self._make_layer(block, planes, layers[i], SE, ECA[i], stride, dil[i - 1] if i else False)
            setattr(self, f"layer{i + 1}", stage)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        # reference init (:176-189): every Conv2d (incl. the depthwise Wv) kaiming fan_out; BN 1/0; last BN of
        # each residual branch zero
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if zero_init_last_bn:
            for m in self.modules():
                if isinstance(m, MRLA_Bottleneck):
                    nn.init.zeros_(m.bn3.weight)

    def _make_layer(self, block, planes, blocks, SE, ECA_size, stride=1, dilate=False):
        prev_dilation = self.dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        cout = planes * block.expansion
        downsample = None
        if stride != 1 or self.inplanes != cout:
            downsample = nn.Sequential(_conv1x1(self.inplanes, cout, stride), self._norm_layer(cout))
        common = dict(SE=SE, ECA_size=ECA_size, groups=self.groups, base_width=self.base_width,
                      norm_layer=self._norm_layer, drop_path=self.drop_path)
        seq = [block(self.inplanes, planes, stride, downsample, dilation=prev_dilation, **common)]
        self.inplanes = cout
        seq += [block(cout, planes, dilation=self.dilation, **common) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def forward_features(self, x):
        x = max_pool(bn_act(self.conv1(promote_images(x)), self.bn1, relu=True), self.maxpool)
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))

    def forward(self, x):
        x = torch.flatten(self.avgpool(self.forward_features(x)), 1)
        if self.drop_rate:
            x = F.dropout(x, p=float(self.drop_rate), training=self.training)
        return self.fc(x)


def resnet50_mrlal(**kwargs):
    return ResNet_mrlal(MRLA_Bottleneck, [3, 4, 6, 3], **kwargs)


def resnet101_mrlal(**kwargs):
    return ResNet_mrlal(MRLA_Bottleneck, [3, 4, 23, 3], **kwargs)

"""ResNet + MRLA-base — host-side mirror of resnet/models/resnet_mrla_base.py (reference): deep 3-conv stem,
per-stage K/V threading `x, k, v = layer(x, k, v)` (reference :254-259), block tail
`out + drop_path(relu(bn_mrla(mrla(out, k, v))))` (reference :124-127) as one fused op over an in-place stage cache.
Same class / factory names, constructor arguments and state_dict keys as the reference."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .drop import DropPath
from .modules.mrla_base_module import mrla_base_layer
from .ops import bn_act, is_plain_batchnorm, max_pool, promote_images
from .resnet_mrla_light import _bn_effective_momentum, _conv1x1, _conv3x3

__all__ = ["ResNet_mrlab", "MRLA_Bottleneck", "mrla_module", "mrla_base_block_tail", "resnet50_mrlab",
           "resnet101_mrlab"]


class mrla_module(nn.Module):
    """Reference resnet_mrla_base.py:32-51 (`dim_perhead = 16`; `channel_wise` -> 1)."""
    dim_perhead = 16

    def __init__(self, input_dim, init_cell=False, channel_wise=False):
        super().__init__()
        if channel_wise:
            self.dim_perhead = 1
        self.mrla = mrla_base_layer(input_dim=input_dim, dim_perhead=self.dim_perhead, init_cell=init_cell)
        self.init_cell = init_cell

    def forward(self, xt, prev_k, prev_v):
        if self.init_cell:
            prev_k = prev_v = None
        return self.mrla(xt, prev_k, prev_v)


def mrla_base_block_tail(out, prev_k, prev_v, mrla: mrla_module, bn: nn.BatchNorm2d, drop_path: nn.Module,
                         relu: bool = True):
    """Fused `out + drop_path(relu(bn(mrla(out, k, v))))` -> (y, k, v)."""
    layer = mrla.mrla
    if mrla.init_cell:
        prev_k = prev_v = None
    if not is_plain_batchnorm(bn):
        # SyncBatchNorm / GroupNorm / ...: the norm module keeps its own semantics (reference :124-127 unfused)
        s, k, v = mrla(out, prev_k, prev_v)
        z = bn(s)
        return out + drop_path(torch.relu(z) if relu else z), k, v
    use_batch_stats = bn.training or bn.running_mean is None
    if use_batch_stats:
        mode = _lib.BN_TRAIN
        momentum = _bn_effective_momentum(bn) if bn.track_running_stats else 0.0
    else:
        mode, momentum = _lib.BN_EVAL, 0.0
    update = bn.training and bn.track_running_stats and bn.running_mean is not None
    scale = drop_path.scale(out) if isinstance(drop_path, DropPath) else None
    cfg = layer.cfg(bn_mode=mode, relu=relu, residual=True, update_running=update, eps=bn.eps, momentum=momentum)
    y, k, v = layer.run(out, prev_k, prev_v, cfg, bn.weight, bn.bias, bn.running_mean, bn.running_var, scale)
    if update:
        bn.num_batches_tracked += 1
    return y, k, v


class MRLA_Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, SE=False, ECA_size=None, groups=1, base_width=64,
                 dilation=1, norm_layer=nn.BatchNorm2d, drop_path=0.0, init_cell=False, channel_wise_mrla=False):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        if SE or ECA_size is not None:
            raise NotImplementedError("SE / ECA channel attention is not part of the MRLA hot path")
        width = int(planes * (base_width / 64.0)) * groups
        cout = planes * self.expansion
        self.conv1, self.bn1 = _conv1x1(inplanes, width), norm_layer(width)
        self.conv2, self.bn2 = _conv3x3(width, width, stride, groups, dilation), norm_layer(width)
        self.conv3, self.bn3 = _conv1x1(width, cout), norm_layer(cout)
        self.relu = nn.ReLU(inplace=True)
        self.downsample, self.stride = downsample, stride
        self.se = self.eca = None
        self.mrla = mrla_module(input_dim=cout, init_cell=init_cell, channel_wise=channel_wise_mrla)
        self.bn_mrla = norm_layer(cout)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x, prev_k, prev_v):
        if self.downsample is None:
            identity = x
        elif isinstance(self.downsample, nn.Sequential) and len(self.downsample) == 2 \
                and isinstance(self.downsample[1], nn.BatchNorm2d):
            identity = bn_act(self.downsample[0](x), self.downsample[1])
        else:
            identity = self.downsample(x)
        out = bn_act(self.conv1(x), self.bn1, relu=True)
        out = bn_act(self.conv2(out), self.bn2, relu=True)
        out = self.relu(bn_act(self.conv3(out), self.bn3) + identity)
        return mrla_base_block_tail(out, prev_k, prev_v, self.mrla, self.bn_mrla, self.drop_path, relu=True)


class ResNet_mrlab(nn.Module):
    def __init__(self, block, layers, num_classes=1000, SE=False, ECA=None, zero_init_last_bn=True, groups=1,
                 width_per_group=64, replace_stride_with_dilation=None, norm_layer=nn.BatchNorm2d, drop_rate=0.0,
                 drop_path=0.0, channel_wise_mrla=False):
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
        sw = 32  # deep stem (reference :176-185)
        self.conv1 = nn.Sequential(
            nn.Conv2d(3, sw, 3, stride=2, padding=1, bias=False), norm_layer(sw), nn.ReLU(inplace=True),
            nn.Conv2d(sw, sw, 3, stride=1, padding=1, bias=False), norm_layer(sw), nn.ReLU(inplace=True),
            nn.Conv2d(sw, self.inplanes, 3, stride=1, padding=1, bias=False))
        self.bn1 = norm_layer(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        cfgs = ((64, 1, False), (128, 2, dil[0]), (256, 2, dil[1]), (512, 2, dil[2]))
        self.stages = nn.ModuleList(
            self._make_layer(block, planes, layers[i], SE, ECA[i], stride, dilate, channel_wise_mrla)
            for i, (planes, stride, dilate) in enumerate(cfgs))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
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

    def _make_layer(self, block, planes, blocks, SE, ECA_size, stride, dilate, channel_wise):
        prev_dilation = self.dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        cout = planes * block.expansion
        downsample = None
        if stride != 1 or self.inplanes != cout:
            downsample = nn.Sequential(_conv1x1(self.inplanes, cout, stride), self._norm_layer(cout))
        common = dict(SE=SE, ECA_size=ECA_size, groups=self.groups, base_width=self.base_width,
                      norm_layer=self._norm_layer, drop_path=self.drop_path, channel_wise_mrla=channel_wise)
        seq = [block(self.inplanes, planes, stride, downsample, dilation=prev_dilation, init_cell=True, **common)]
        self.inplanes = cout
        seq += [block(cout, planes, dilation=self.dilation, init_cell=False, **common) for _ in range(1, blocks)]
        return nn.ModuleList(seq)

    def forward_features(self, x):
        x = max_pool(bn_act(self.conv1(promote_images(x)), self.bn1, relu=True), self.maxpool)
        k = v = None
        for stage in self.stages:
            for blk in stage:
                x, k, v = blk(x, k, v)
        return x

    def forward(self, x):
        x = torch.flatten(self.avgpool(self.forward_features(x)), 1)
        if self.drop_rate:
            x = F.dropout(x, p=float(self.drop_rate), training=self.training)
        return self.fc(x)


def resnet50_mrlab(**kwargs):
    return ResNet_mrlab(MRLA_Bottleneck, [3, 4, 6, 3], **kwargs)


def resnet101_mrlab(**kwargs):
    return ResNet_mrlab(MRLA_Bottleneck, [3, 4, 23, 3], **kwargs)

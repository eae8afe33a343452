"""CPU port of the reference `resnet50_mrlal` / `resnet101_mrlal` (TEST INFRASTRUCTURE + CPU baseline).

Plain eager PyTorch with the reference's module tree and state_dict keys, so weights can be copied
key-for-key between the reference, this port and the product model (mrla_b200.resnet_mrla_light).
The MRLA tail is evaluated through the functional restatement in oracle/mrla_oracle.py — i.e. the same
~14 ATen calls per block the reference issues (resnet/models/resnet_mrla_light.py:89-118).

Used by: tests (whole-model parity: reference == this port on CPU, product == this port on GPU) and by
`bench.py`'s cpu_baseline / `--impl reference` legs, which time it on the GPU box's host cores because
/root/reference itself is not available there.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import mrla_oracle as O


class _LightLayerParams(nn.Module):
    """Parameter holder with the reference names (mrla_light_module.py:45-48)."""

    def __init__(self, c, d):
        super().__init__()
        k = O.eca_kernel_size(c)
        self.heads = int(c / d)
        self.Wq = nn.Conv1d(1, 1, k, padding=(k - 1) // 2, bias=False)
        self.Wk = nn.Conv1d(1, 1, k, padding=(k - 1) // 2, bias=False)
        self.Wv = nn.Conv2d(c, c, 3, 1, 1, groups=c, bias=False)


class _MrlaModule(nn.Module):
    def __init__(self, c, d=32):
        super().__init__()
        self.mrla = _LightLayerParams(c, d)
        self.lambda_t = nn.Parameter(torch.randn(c, 1, 1))

    def forward(self, xt, ot_1):
        m = self.mrla
        return O.light_module(xt, ot_1, m.Wq.weight, m.Wk.weight, m.Wv.weight, self.lambda_t, m.heads)


class _Block(nn.Module):
    def __init__(self, cin, planes, stride, downsample, drop_path):
        super().__init__()
        cout = planes * 4
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.downsample = downsample
        self.mrla = _MrlaModule(cout)
        self.bn_mrla = nn.BatchNorm2d(cout)
        self.drop_prob = drop_path

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = F.relu(self.bn3(self.conv3(out)) + idt)
        z = self.bn_mrla(self.mrla(out, idt))
        m = O.drop_path_scale(z.shape[0], self.drop_prob, self.training, z)
        if m is not None:
            z = z * m.view(-1, 1, 1, 1)
        return out + z


class ResNetMrlalOracle(nn.Module):
    def __init__(self, layers, num_classes=1000, drop_path=0.0):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if i == 0 else 2
            blocks = []
            for j in range(n):
                ds = None
                if j == 0 and (stride != 1 or cin != planes * 4):
                    ds = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
                blocks.append(_Block(cin, planes, stride if j == 0 else 1, ds, drop_path))
                cin = planes * 4
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        self.fc = nn.Linear(cin, num_classes)
        for m in self.modules():  # resnet_mrla_light.py:176-189
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        for m in self.modules():
            if isinstance(m, _Block):
                nn.init.zeros_(m.bn3.weight)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))


def resnet50_mrlal_oracle(**kw):
    return ResNetMrlalOracle([3, 4, 6, 3], **kw)


def resnet101_mrlal_oracle(**kw):
    return ResNetMrlalOracle([3, 4, 23, 3], **kw)


# ----------------------------------------------------------------------------------------------- MRLA-base port
class _BaseLayerParams(nn.Module):
    """Parameter holder with the reference names (mrla_base_module.py:46-50)."""

    def __init__(self, c, d, init_cell):
        super().__init__()
        k = O.eca_kernel_size(c)
        self.heads = int(c / d)
        self.init_cell = init_cell
        self.Wq = nn.Conv1d(1, 1, k, padding=(k - 1) // 2, bias=False)
        self.Wk = nn.Conv1d(1, 1, k, padding=(k - 1) // 2, bias=False)
        self.Wv = nn.Conv2d(c, c, 3, 1, 1, groups=c, bias=False)


class _MrlaBaseModule(nn.Module):
    def __init__(self, c, init_cell, d=16):
        super().__init__()
        self.mrla = _BaseLayerParams(c, d, init_cell)


class _BaseBlock(nn.Module):
    """resnet/models/resnet_mrla_base.py:54-129."""

    def __init__(self, cin, planes, stride, downsample, drop_path, init_cell):
        super().__init__()
        cout = planes * 4
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.downsample = downsample
        self.mrla = _MrlaBaseModule(cout, init_cell)
        self.bn_mrla = nn.BatchNorm2d(cout)
        self.drop_prob = drop_path

    def forward(self, x, k, v):
        idt = x if self.downsample is None else self.downsample(x)
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = F.relu(self.bn3(self.conv3(out)) + idt)
        m = self.mrla.mrla
        s, k, v = O.base_layer(out, k, v, m.Wq.weight, m.Wk.weight, m.Wv.weight, m.heads, m.init_cell)
        z = F.relu(self.bn_mrla(s))
        dm = O.drop_path_scale(z.shape[0], self.drop_prob, self.training, z)
        if dm is not None:
            z = z * dm.view(-1, 1, 1, 1)
        return out + z, k, v


class ResNetMrlabOracle(nn.Module):
    def __init__(self, layers, num_classes=1000, drop_path=0.0):
        super().__init__()
        sw = 32
        self.conv1 = nn.Sequential(nn.Conv2d(3, sw, 3, 2, 1, bias=False), nn.BatchNorm2d(sw), nn.ReLU(inplace=True),
                                   nn.Conv2d(sw, sw, 3, 1, 1, bias=False), nn.BatchNorm2d(sw), nn.ReLU(inplace=True),
                                   nn.Conv2d(sw, 64, 3, 1, 1, bias=False))
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        stages = []
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if i == 0 else 2
            blocks = []
            for j in range(n):
                ds = None
                if j == 0 and (stride != 1 or cin != planes * 4):
                    ds = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
                blocks.append(_BaseBlock(cin, planes, stride if j == 0 else 1, ds, drop_path, init_cell=(j == 0)))
                cin = planes * 4
            stages.append(nn.ModuleList(blocks))
        self.stages = nn.ModuleList(stages)
        self.fc = nn.Linear(cin, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        for m in self.modules():
            if isinstance(m, _BaseBlock):
                nn.init.zeros_(m.bn3.weight)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        k = v = None
        for stage in self.stages:
            for blk in stage:
                x, k, v = blk(x, k, v)
        return self.fc(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))

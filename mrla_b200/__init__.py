"""mrla_b200 — B200-native (sm_100a) implementation of the MRLA layer-attention block tail.

Public surface mirrors the reference (joyfang1106/MRLA) module API for the hot path:

    mrla_b200.modules.mrla_light_layer            resnet/models/modules/mrla_light_module.py
    mrla_b200.modules.mrla_base_layer             resnet/models/modules/mrla_base_module.py
    mrla_b200.resnet_mrla_light.{mrla_module, MRLA_Bottleneck, ResNet_mrlal, resnet50_mrlal, resnet101_mrlal}
    mrla_b200.resnet_mrla_base.{mrla_module, MRLA_Bottleneck, ResNet_mrlab, resnet50_mrlab, resnet101_mrlab}
    mrla_b200.deit_mrla_light.{mrlal_layer, mrlal_module}, mrla_b200.deit_mrla_base.{mrlab_layer, mrlab_module}

All arithmetic of the tail runs in hand-written CUDA kernels reached through the C ABI in
include/mrla_b200.h (libmrla_b200.so).  There is no CPU / eager fallback: using the modules
without the built library or on non-CUDA tensors raises.
"""
from . import _lib  # noqa: F401
from .ops import LightCfg, light_tail  # noqa: F401

__version__ = "0.1.0"

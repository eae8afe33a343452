"""Import the UNMODIFIED reference classes (dev container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so nothing
that runs there may call this module; it is used by ``tests/golden/make_golden.py``
(to freeze golden vectors) and by ``tests/test_oracle_vs_reference.py`` (skipped when
the reference tree is absent).

Two work-arounds are needed to import the reference as shipped:

* ``import models`` raises ``AttributeError`` because ``__all__`` names symbols that
  are never defined (resnet/models/resnet_mrla_light.py:14-18 vs :242-250), so a stub
  ``models`` package is registered and the sub-modules are imported individually
  (they use absolute ``from models.modules...`` imports, resnet_mrla_light.py:8-11).
* ``timm`` is not installed; the DeiT files only need a handful of names from it
  (deit/deit_mrla_light.py:16-22), which are stubbed in ``sys.modules``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("MRLA_REF", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "resnet", "models"))


def _stub_models_pkg():
    if "models" in sys.modules and getattr(sys.modules["models"], "__mrla_stub__", False):
        return
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF_ROOT, "resnet", "models")]
    pkg.__mrla_stub__ = True
    sys.modules["models"] = pkg


def resnet_light():
    """-> module object of resnet/models/resnet_mrla_light.py"""
    _stub_models_pkg()
    return importlib.import_module("models.resnet_mrla_light")


def resnet_base():
    """-> module object of resnet/models/resnet_mrla_base.py"""
    _stub_models_pkg()
    return importlib.import_module("models.resnet_mrla_base")


def light_layer_mod():
    _stub_models_pkg()
    return importlib.import_module("models.modules.mrla_light_module")


def base_layer_mod():
    _stub_models_pkg()
    return importlib.import_module("models.modules.mrla_base_module")


def drop_mod():
    _stub_models_pkg()
    return importlib.import_module("models.utils.drop")


def _stub_timm():
    if "timm" in sys.modules:
        return
    import torch
    import torch.nn as nn

    def _mk(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    timm = _mk("timm")
    tm = _mk("timm.models")
    vt = _mk("timm.models.vision_transformer")
    reg = _mk("timm.models.registry")
    lay = _mk("timm.models.layers")
    hlp = _mk("timm.models.layers.helpers")
    timm.models = tm
    tm.vision_transformer, tm.registry, tm.layers = vt, reg, lay
    lay.helpers = hlp
    vt.default_cfgs = {}
    vt._cfg = lambda **kw: dict(kw)
    reg.register_model = lambda fn: fn
    lay.trunc_normal_ = lambda t, std=1.0, **kw: nn.init.trunc_normal_(t, std=std)
    hlp.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    class DropPath(nn.Module):  # timm.models.layers.DropPath semantics (per-sample)
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    lay.DropPath = DropPath


def _import_deit(name):
    _stub_timm()
    d = os.path.join(REF_ROOT, "deit")
    if d not in sys.path:
        sys.path.insert(0, d)
    return importlib.import_module(name)


def deit_light():
    """-> module object of deit/deit_mrla_light.py"""
    return _import_deit("deit_mrla_light")


def deit_base():
    """-> module object of deit/deit_mrla_base.py"""
    return _import_deit("deit_mrla_base")

"""Drop-in module aliases so the reference's model files import unmodified.

models/cross.py:9,14,16,17 and models/vmamba_Fusion_efficross.py:12,14-16 import, at module import time,
``selective_scan_cuda``, ``mamba_ssm.Mamba``, ``mamba_ssm.ops.selective_scan_interface`` and
``timm.models.layers``.  ``install()`` registers those names in ``sys.modules`` (the same modules also
exist as files under ``compat/`` for the PYTHONPATH route described in INTEGRATION.md).
"""
from __future__ import annotations

import importlib
import sys
import types


def _timm_stub() -> None:
    """Minimal ``timm.models.layers`` (DropPath, to_2tuple, trunc_normal_) -- registered only if timm is absent."""
    import torch
    from torch import nn

    class DropPath(nn.Module):
        """Stochastic depth per sample (same semantics as timm.layers.DropPath)."""

        def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
            super().__init__()
            self.drop_prob = float(drop_prob)
            self.scale_by_keep = scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                mask.div_(keep)
            return x * mask

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
    timm.models, models.layers = models, layers
    timm.__fm_stub__ = True
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def install(force_timm_stub: bool = False) -> None:
    from . import interface, scan_cuda

    ssc = types.ModuleType("selective_scan_cuda")
    ssc.__doc__ = "fusionmamba_b200 drop-in for the reference pybind module (selective_scan/selective_scan.cpp:494-497)"
    ssc.fwd, ssc.bwd = scan_cuda.fwd, scan_cuda.bwd
    sys.modules["selective_scan_cuda"] = ssc

    class Mamba:  # models/cross.py:9 imports the name; FusionMamba never instantiates it
        def __init__(self, *a, **k):
            raise NotImplementedError("the 1-D Mamba block is outside fusionmamba_b200's scope (SURVEY.md section 2 #18)")

    ms = types.ModuleType("mamba_ssm")
    ops = types.ModuleType("mamba_ssm.ops")
    iface = types.ModuleType("mamba_ssm.ops.selective_scan_interface")
    for name in ("SelectiveScanFn", "selective_scan_fn", "selective_scan_ref"):
        setattr(iface, name, getattr(interface, name))
    iface.selective_scan_cuda = ssc
    ms.Mamba, ms.ops, ops.selective_scan_interface = Mamba, ops, iface
    ms.selective_scan_fn = interface.selective_scan_fn
    sys.modules.update({"mamba_ssm": ms, "mamba_ssm.ops": ops, "mamba_ssm.ops.selective_scan_interface": iface})

    need_stub = force_timm_stub
    if not need_stub:
        try:
            importlib.import_module("timm.models.layers")
        except Exception:
            need_stub = True
    if need_stub:
        _timm_stub()

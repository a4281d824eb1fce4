"""fusionmamba_b200 -- B200 (sm_100a) implementation of FusionMamba's SS2D hot path.

Public surface (mirrors the reference's operator API for this path and nothing else):
  * ``scan_cuda.fwd`` / ``scan_cuda.bwd``            == ``selective_scan_cuda.fwd`` / ``.bwd``
  * ``selective_scan_fn`` / ``selective_scan_ref``    == mamba_ssm.ops.selective_scan_interface
  * ``compat.install()``                              makes ``import selective_scan_cuda``, ``from mamba_ssm import Mamba``,
                                                      ``mamba_ssm.ops.selective_scan_interface`` and (only if absent)
                                                      ``timm.models.layers`` resolve, so reference model files run unmodified
The compute lives in ``libfm_scan.so`` (C ABI: include/fm_scan.h); there is no CPU / PyTorch fallback.
"""
from .interface import SelectiveScanFn, selective_scan_fn, selective_scan_ref  # noqa: F401
from . import scan_cuda  # noqa: F401

__version__ = "0.1.0"

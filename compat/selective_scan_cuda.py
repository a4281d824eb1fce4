"""Drop-in for the reference's pybind extension module ``selective_scan_cuda``
(selective_scan/selective_scan.cpp:494-497): put ``<repo>/compat`` and ``<repo>`` on PYTHONPATH and the
reference's ``import selective_scan_cuda`` binds to the sm_100a kernels."""
from fusionmamba_b200.scan_cuda import bwd, fwd  # noqa: F401

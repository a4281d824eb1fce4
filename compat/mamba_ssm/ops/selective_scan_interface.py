"""``mamba_ssm.ops.selective_scan_interface`` with the reference's names (selective_scan_interface.py:20-158)."""
import selective_scan_cuda  # noqa: F401  (same import the reference performs at :16)
from fusionmamba_b200.interface import SelectiveScanFn, selective_scan_fn, selective_scan_ref  # noqa: F401

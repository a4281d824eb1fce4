"""Import shim for ``from mamba_ssm import Mamba`` (models/cross.py:9). Only the selective-scan operator is provided."""
from fusionmamba_b200.interface import selective_scan_fn  # noqa: F401


class Mamba:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("the 1-D Mamba block is outside fusionmamba_b200's scope (SURVEY.md section 2 #18)")

from fusionmamba_b200 import compat as _c

_c._timm_stub()
import sys as _s

DropPath = _s.modules["timm.models.layers"].DropPath
to_2tuple = _s.modules["timm.models.layers"].to_2tuple
trunc_normal_ = _s.modules["timm.models.layers"].trunc_normal_

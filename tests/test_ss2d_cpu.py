"""CPU-side checks of the SS2D mirror: state_dict key/shape compatibility with the reference's modules (keys and shapes
come from fixtures written by the reference's own classes), constructor defaults, scan lengths, the opt-in patcher, and
that the CUDA-only path refuses CPU tensors loudly instead of falling back."""
import ast
import os
import types

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,cls_name", [("mod_v2", "SS2D"), ("mod_v2_odd", "SS2D"), ("mod_v0", "SS2D"), ("mod_cross", "SS2D_cross_new")])
def test_state_dict_keys_and_shapes_match_reference(name, cls_name):
    from fusionmamba_b200 import ss2d
    g = np.load(os.path.join(GOLD, f"ss2d_{name}.npz"))
    m = getattr(ss2d, cls_name)(**ast.literal_eval(str(g["kwargs"])))
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {k[3:]: tuple(g[k].shape) for k in g.files if k.startswith("sd/")}
    assert ours == ref
    m.load_state_dict({k: torch.from_numpy(g["sd/" + k]) for k in ref}, strict=True)


def test_stage0_shapes_are_the_published_ones():
    # stage-0 parameter shapes of the published checkpoints
    from fusionmamba_b200 import ss2d
    sd = ss2d.SS2D(d_model=96).state_dict()
    assert tuple(sd["x_proj_weight"].shape) == (4, 38, 192)
    assert tuple(sd["dt_projs_weight"].shape) == (4, 192, 6)
    assert tuple(sd["dt_projs_bias"].shape) == (4, 192)
    assert tuple(sd["A_logs"].shape) == (768, 16) and tuple(sd["Ds"].shape) == (768,)
    assert tuple(sd["in_proj.weight"].shape) == (384, 96) and tuple(sd["out_proj.weight"].shape) == (96, 192)
    assert tuple(sd["conv2d.weight"].shape) == (192, 1, 3, 3) and tuple(sd["out_norm.weight"].shape) == (192,)
    assert torch.allclose(sd["A_logs"][5], torch.log(torch.arange(1, 17, dtype=torch.float32)))
    c = ss2d.SS2D_cross_new(d_model=96).state_dict()
    assert "in_proj1.weight" in c and "in_proj2.weight" in c and "in_proj.weight" not in c


def test_scan_len_and_modes():
    from fusionmamba_b200 import ss2d
    assert ss2d.scan_len(64, 64, ss2d.MAP_V0) == 4096 and ss2d.scan_len(64, 64, ss2d.MAP_V2) == 1024
    assert ss2d.scan_len(7, 5, ss2d.MAP_V2) == 12 and ss2d.scan_len(1, 1, ss2d.MAP_V2) == 1


def test_patch_reference_rebinds_module_level_functions():
    from fusionmamba_b200 import ss2d
    fake = types.SimpleNamespace(cross_selective_scan=None, cross_selective_scan_cross=None)
    ss2d.patch_reference(fake)
    assert fake.cross_selective_scan is ss2d.cross_selective_scan
    assert fake.cross_selective_scan_cross is ss2d.cross_selective_scan_cross


def test_no_cpu_fallback():
    from fusionmamba_b200 import ss2d
    with pytest.raises(RuntimeError, match="CUDA"):
        ss2d.scan_unfold(torch.randn(1, 2, 4, 4))
    with pytest.raises(NotImplementedError):
        ss2d.cross_selective_scan(torch.randn(1, 2, 4, 4), step_size=3)

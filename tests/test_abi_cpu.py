"""CPU-tier checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/fm_scan.h declares,
rejects bad arguments without touching a GPU, and the ctypes records match the C structs byte for byte."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "fm_scan.h")


@pytest.fixture(scope="module")
def lib():
    from fusionmamba_b200 import _lib, build
    build.build()
    return _lib.lib()


def test_exports_match_header(lib):
    from fusionmamba_b200 import _lib
    src = open(HDR).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char \*)\s*\*?\s*(fm_\w+)\s*\(", src, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fm_abi_version() == _lib.ABI_VERSION and lib.fm_target_sm() == 100


def test_struct_layout_matches_c(tmp_path):
    """sizeof / offsetof from a C translation unit including the header == the ctypes mirrors."""
    from fusionmamba_b200 import _lib
    prog = tmp_path / "lay.c"
    fields = {
        "FmScanFwdParams": ["abi_version", "seqlen", "hck_len", "u_batch_stride", "C_dstate_stride", "u", "D", "x", "hck", "workspace",
                            "workspace_bytes"],
        "FmScanBwdParams": ["f", "dout_batch_stride", "dC_dstate_stride", "dout", "dA", "ddelta_bias"],
        "FmPermuteParams": ["abi_version", "map", "w", "src", "dst"],
        "FmNormBwdParams": ["abi_version", "rows", "eps", "x", "dbias", "workspace", "workspace_bytes"],
        "FmBlockGatesParams": ["abi_version", "reduce_dim", "eps", "x", "eca_weight", "b2", "se_gate", "workspace_bytes"],
        "FmBlockScaleParams": ["abi_version", "dim", "x", "gate", "y"],
        "FmBlockCombineParams": ["abi_version", "dim", "eps", "input_dtype", "input", "gate_conv", "ln_bias", "y_out"],
        "FmConvUnfoldBwdParams": ["abi_version", "w", "dsrc_channel_offset", "src_channel_stride", "src", "dxs", "dsrc", "dbias"],
    }
    body = "".join(
        f'printf("{s} %zu\\n", sizeof({s}));' + "".join(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));' for f in fs)
        for s, fs in fields.items())
    prog.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{HDR}"\nint main(void){{{body}return 0;}}')
    exe = tmp_path / "lay"
    subprocess.run(["/usr/bin/gcc", str(prog), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for s, fs in fields.items():
        ct = getattr(_lib, s)
        assert int(out[s]) == C.sizeof(ct), s
        for f in fs:
            assert int(out[f"{s}.{f}"]) == getattr(ct, f).offset, f"{s}.{f}"


def test_invalid_arguments_fail_loudly_without_gpu(lib):
    from fusionmamba_b200 import _lib
    p = _lib.FmScanFwdParams()
    assert lib.fm_selective_scan_fwd(C.byref(p), None) == 1          # abi_version 0
    assert b"abi_version" in lib.fm_last_error()
    p.abi_version = _lib.ABI_VERSION
    p.batch, p.dim, p.seqlen, p.dstate, p.n_groups = 1, 6, 10, 300, 4
    p.chunk_len, p.n_chunks = 2048, 1
    assert lib.fm_selective_scan_fwd(C.byref(p), None) == 1
    assert b"state dimension <= 256" in lib.fm_last_error()          # same message as selective_scan.cpp:264
    p.dstate = 16
    assert lib.fm_selective_scan_fwd(C.byref(p), None) == 1 and b"multiple of n_groups" in lib.fm_last_error()
    p.n_groups = 2
    assert lib.fm_selective_scan_fwd(C.byref(p), None) == 1 and b"non-null" in lib.fm_last_error()
    assert lib.fm_selective_scan_bwd(None, None) == 1
    q = _lib.FmPermuteParams()
    assert lib.fm_scan_unfold(C.byref(q), None) == 1


def test_python_shim_surface():
    """Names, signatures and import paths the reference's model files rely on."""
    import inspect
    import fusionmamba_b200 as fm
    from fusionmamba_b200 import compat
    compat.install()
    import selective_scan_cuda
    from mamba_ssm import Mamba  # noqa: F401
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn, selective_scan_ref
    from timm.models.layers import DropPath, to_2tuple, trunc_normal_  # noqa: F401
    want = ["u", "delta", "A", "B", "C", "D", "z", "delta_bias", "delta_softplus", "return_last_state"]
    assert list(inspect.signature(selective_scan_fn).parameters) == want
    assert list(inspect.signature(selective_scan_ref).parameters) == want
    assert len(inspect.signature(selective_scan_cuda.fwd).parameters) == 9       # selective_scan.cpp:226-232
    assert len(inspect.signature(selective_scan_cuda.bwd).parameters) == 14      # selective_scan.cpp:338-349
    assert selective_scan_fn is fm.selective_scan_fn


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing on the host."""
    import torch
    import fusionmamba_b200 as fm
    u = torch.randn(1, 4, 8); A = -torch.rand(4, 2); B = torch.randn(1, 1, 2, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        fm.selective_scan_fn(u, u.clone(), A, B, B.clone())


def test_selective_scan_ref_matches_oracle():
    """API-parity definition of the op (pure PyTorch, CPU) agrees with the oracle on a golden case."""
    import numpy as np
    import torch
    from fusionmamba_b200 import selective_scan_ref
    d = dict(np.load(os.path.join(ROOT, "tests", "golden", "scan_f32_z.npz")))
    t = lambda k: torch.from_numpy(d[k]) if k in d else None
    out, last = selective_scan_ref(t("u"), t("delta"), t("A"), t("B"), t("C"), t("D"), t("z"), t("delta_bias"), True, True)
    np.testing.assert_allclose(out.numpy(), d["out"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(last.numpy(), d["last_state"], rtol=2e-4, atol=2e-5)

"""Host-only checks of the launch planning exported by the C ABI (no CUDA call is made)."""
import ctypes as C

import pytest


def _params(batch, dim, L, groups=4, dstate=16):
    from fusionmamba_b200 import _lib
    p = _lib.FmScanFwdParams()
    p.abi_version, p.dtype, p.out_dtype = _lib.ABI_VERSION, _lib.FM_F32, _lib.FM_F32
    p.batch, p.dim, p.seqlen, p.dstate, p.n_groups = batch, dim, L, dstate, groups
    p.chunk_len, p.n_chunks = 2048, (L + 2047) // 2048
    return p


def test_time_split_workspace_plan():
    """fm_scan_fwd_workspace_bytes: non-zero only for few-row, long-sequence forwards (BASELINE configs[4]: one 1024x1024 pair);
    every batched shape of the model and the configs[1] benchmark keep the single pass (0 bytes)."""
    from fusionmamba_b200 import _lib
    lib = _lib.lib()
    q = lambda *a, **k: int(lib.fm_scan_fwd_workspace_bytes(C.byref(_params(*a, **k))))
    long_ = q(1, 768, 65536)
    assert long_ > 0 and long_ % (768 * 36 * 4) == 0 and 2 <= long_ // (768 * 36 * 4) <= 16      # rows x segments x 36 floats
    assert q(1, 768, 16384) > 0 and q(1, 1536, 4096) > 0          # stage 0 / 1 of one 1024^2 pair
    assert q(8, 768, 4096) == 0                                   # configs[1]
    for dim, L in ((768, 1024), (1536, 256), (3072, 64), (6144, 16)):
        assert q(32, dim, L) == 0                                 # batch-32 stage shapes
    assert q(1, 768, 512) == 0                                    # too short to split
    assert q(1, 96, 65536) == 0                                   # 24 channels per group: not a multiple of the 16-row tile
    assert q(1, 768, 65536, dstate=8) == 0                        # only the dstate-16 kernel splits
    bad = _params(1, 768, 65536)
    bad.abi_version = 0
    assert int(lib.fm_scan_fwd_workspace_bytes(C.byref(bad))) == 0
    assert int(lib.fm_scan_fwd_workspace_bytes(None)) == 0


def test_checkpoint_spacing_selects_backward_kernel():
    from fusionmamba_b200 import scan_cuda
    assert scan_cuda._hck_len(16, 16) == 8 and scan_cuda._hck_len(16, 512) == 8          # lane-serial backward
    assert scan_cuda._hck_len(16, 513) == 64 and scan_cuda._hck_len(16, 4096) == 64      # row-pair backward (few rows)
    assert scan_cuda._hck_len(16, 4096, 8 * 768) == 8 and scan_cuda._hck_len(16, 65536, 768) == 64   # lane-serial once 8-row warps fill the GPU
    assert scan_cuda._n_hck(4096, 16, 8 * 768) == 511
    assert scan_cuda._hck_len(8, 64) == 64 and scan_cuda._hck_len(128, 64) == 16
    assert scan_cuda._n_hck(4096, 16) == 63 and scan_cuda._n_hck(256, 16) == 31 and scan_cuda._n_hck(8, 16) == 0

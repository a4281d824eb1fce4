"""ctypes binding of libfm_scan.so (C ABI declared in include/fm_scan.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(the product path never routes through the CPU oracle or a PyTorch re-implementation).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfm_scan.so")
ABI_VERSION = 3

FM_F32, FM_F16, FM_BF16 = 0, 1, 2
FM_MAP_LINEAR, FM_MAP_CROSS_V0, FM_MAP_EFFICIENT_V2, FM_MAP_EFFICIENT_V2_CL = 0, 1, 2, 3

_i32, _i64, _vp = C.c_int32, C.c_int64, C.c_void_p


class FmScanFwdParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dtype", _i32),
        ("batch", _i32), ("dim", _i32), ("seqlen", _i32), ("dstate", _i32), ("n_groups", _i32),
        ("n_chunks", _i32), ("chunk_len", _i32), ("delta_softplus", _i32),
        ("u_map", _i32), ("out_map", _i32), ("map_h", _i32), ("map_w", _i32),
        ("hck_len", _i32), ("n_hck", _i32), ("out_dtype", _i32), ("reserved0", _i32),
        ("u_batch_stride", _i64), ("u_d_stride", _i64),
        ("delta_batch_stride", _i64), ("delta_d_stride", _i64),
        ("z_batch_stride", _i64), ("z_d_stride", _i64),
        ("out_batch_stride", _i64), ("out_d_stride", _i64),
        ("out_z_batch_stride", _i64), ("out_z_d_stride", _i64),
        ("A_d_stride", _i64), ("A_dstate_stride", _i64),
        ("B_batch_stride", _i64), ("B_group_stride", _i64), ("B_dstate_stride", _i64),
        ("C_batch_stride", _i64), ("C_group_stride", _i64), ("C_dstate_stride", _i64),
        ("u", _vp), ("delta", _vp), ("A", _vp), ("B", _vp), ("C", _vp),
        ("D", _vp), ("z", _vp), ("delta_bias", _vp),
        ("out", _vp), ("out_z", _vp), ("x", _vp), ("hck", _vp),
        ("workspace", _vp), ("workspace_bytes", _i64),
    ]


class FmScanBwdParams(C.Structure):
    _fields_ = [
        ("f", FmScanFwdParams),
        ("dout_batch_stride", _i64), ("dout_d_stride", _i64),
        ("du_batch_stride", _i64), ("du_d_stride", _i64),
        ("ddelta_batch_stride", _i64), ("ddelta_d_stride", _i64),
        ("dz_batch_stride", _i64), ("dz_d_stride", _i64),
        ("dB_batch_stride", _i64), ("dB_group_stride", _i64), ("dB_dstate_stride", _i64),
        ("dC_batch_stride", _i64), ("dC_group_stride", _i64), ("dC_dstate_stride", _i64),
        ("dout", _vp), ("du", _vp), ("ddelta", _vp), ("dz", _vp),
        ("dA", _vp), ("dB", _vp), ("dC", _vp), ("dD", _vp), ("ddelta_bias", _vp),
    ]


class FmPermuteParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dtype", _i32), ("map", _i32),
        ("batch", _i32), ("dim", _i32), ("h", _i32), ("w", _i32),
        ("src", _vp), ("dst", _vp),
    ]


class FmNormParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("out_dtype", _i32),
        ("batch", _i32), ("dim", _i32), ("positions", _i32), ("eps", C.c_float),
        ("src", _vp), ("weight", _vp), ("bias", _vp), ("dst", _vp),
        ("gate", _vp), ("gate_channel_stride", _i64), ("gate_channel_offset", _i32), ("src_channels_last", _i32),
    ]


class FmNormBwdParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dim", _i32), ("rows", _i64), ("eps", C.c_float), ("reserved0", _i32),
        ("x", _vp), ("dy", _vp), ("weight", _vp), ("dx", _vp), ("dweight", _vp), ("dbias", _vp),
        ("workspace", _vp), ("workspace_bytes", _i64),
    ]


class FmBlockGatesParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dtype", _i32), ("batch", _i32), ("positions", _i32), ("dim", _i32), ("reduce_dim", _i32),
        ("eps", C.c_float), ("reserved0", _i32),
        ("x", _vp), ("ln_weight", _vp), ("ln_bias", _vp), ("eca_weight", _vp),
        ("w1", _vp), ("b1", _vp), ("w2", _vp), ("b2", _vp),
        ("eca_scale", _vp), ("se_gate", _vp), ("workspace", _vp), ("workspace_bytes", _i64),
    ]


class FmBlockScaleParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dtype", _i32), ("batch", _i32), ("positions", _i32), ("dim", _i32), ("reserved0", _i32),
        ("x", _vp), ("gate", _vp), ("y", _vp),
    ]


class FmBlockCombineParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dtype", _i32), ("batch", _i32), ("positions", _i32), ("dim", _i32), ("eps", C.c_float),
        ("input_dtype", _i32), ("reserved0", _i32),
        ("input", _vp), ("x_ssm", _vp), ("x_conv", _vp), ("gate_ssm", _vp), ("gate_conv", _vp),
        ("ln_weight", _vp), ("ln_bias", _vp), ("x_out", _vp), ("y_out", _vp),
    ]


class FmConvUnfoldParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dtype", _i32),
        ("batch", _i32), ("dim", _i32), ("h", _i32), ("w", _i32),
        ("src_channel_offset", _i32), ("reserved0", _i32), ("src_channel_stride", _i64),
        ("src", _vp), ("weight", _vp), ("bias", _vp), ("dst", _vp),
    ]


class FmConvUnfoldBwdParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dtype", _i32),
        ("batch", _i32), ("dim", _i32), ("h", _i32), ("w", _i32),
        ("src_channel_offset", _i32), ("dsrc_channel_offset", _i32),
        ("src_channel_stride", _i64), ("dsrc_channel_stride", _i64),
        ("src", _vp), ("weight", _vp), ("bias", _vp), ("dxs", _vp),
        ("dsrc", _vp), ("dweight", _vp), ("dbias", _vp),
    ]


class FmDtProjParams(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("dtype", _i32), ("weight_dtype", _i32),
        ("batch", _i32), ("n_groups", _i32), ("dim", _i32), ("rank", _i32), ("seqlen", _i32),
        ("src_batch_stride", _i64), ("src_group_stride", _i64), ("src_rank_stride", _i64),
        ("src", _vp), ("weight", _vp), ("dst", _vp),
    ]


EXPORTS = (
    "fm_selective_scan_fwd", "fm_selective_scan_bwd", "fm_scan_unfold", "fm_scan_merge", "fm_merge_norm", "fm_conv_unfold", "fm_dt_proj",
    "fm_last_error", "fm_abi_version", "fm_target_sm", "fm_launch_count", "fm_scan_fwd_workspace_bytes",
    "fm_layer_norm_bwd", "fm_layer_norm_bwd_workspace_bytes",
    "fm_block_gates", "fm_block_gates_workspace_bytes", "fm_block_scale", "fm_block_combine_norm", "fm_conv_unfold_bwd",
)

_lib = None


def lib() -> C.CDLL:
    """Load libfm_scan.so (once). Raises RuntimeError if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the sm_100a extension first (python -m fusionmamba_b200.build or "
            "__graft_entry__.build()); fusionmamba_b200 has no CPU or PyTorch fallback path")
    L = C.CDLL(LIB_PATH)
    L.fm_selective_scan_fwd.argtypes = [C.POINTER(FmScanFwdParams), _vp]
    L.fm_selective_scan_fwd.restype = C.c_int
    L.fm_selective_scan_bwd.argtypes = [C.POINTER(FmScanBwdParams), _vp]
    L.fm_selective_scan_bwd.restype = C.c_int
    L.fm_scan_unfold.argtypes = [C.POINTER(FmPermuteParams), _vp]
    L.fm_scan_unfold.restype = C.c_int
    L.fm_scan_merge.argtypes = [C.POINTER(FmPermuteParams), _vp]
    L.fm_scan_merge.restype = C.c_int
    L.fm_merge_norm.argtypes = [C.POINTER(FmNormParams), _vp]
    L.fm_merge_norm.restype = C.c_int
    L.fm_conv_unfold.argtypes = [C.POINTER(FmConvUnfoldParams), _vp]
    L.fm_conv_unfold.restype = C.c_int
    L.fm_conv_unfold_bwd.argtypes = [C.POINTER(FmConvUnfoldBwdParams), _vp]
    L.fm_conv_unfold_bwd.restype = C.c_int
    L.fm_dt_proj.argtypes = [C.POINTER(FmDtProjParams), _vp]
    L.fm_dt_proj.restype = C.c_int
    L.fm_last_error.restype = C.c_char_p
    L.fm_abi_version.restype = C.c_int
    L.fm_target_sm.restype = C.c_int
    L.fm_launch_count.restype = C.c_int64
    L.fm_layer_norm_bwd.argtypes = [C.POINTER(FmNormBwdParams), _vp]
    L.fm_layer_norm_bwd.restype = C.c_int
    L.fm_layer_norm_bwd_workspace_bytes.argtypes = [_i32, _i64]
    L.fm_layer_norm_bwd_workspace_bytes.restype = C.c_int64
    L.fm_block_gates.argtypes = [C.POINTER(FmBlockGatesParams), _vp]
    L.fm_block_gates.restype = C.c_int
    L.fm_block_gates_workspace_bytes.argtypes = [_i32, _i32, _i32]
    L.fm_block_gates_workspace_bytes.restype = C.c_int64
    L.fm_block_scale.argtypes = [C.POINTER(FmBlockScaleParams), _vp]
    L.fm_block_scale.restype = C.c_int
    L.fm_block_combine_norm.argtypes = [C.POINTER(FmBlockCombineParams), _vp]
    L.fm_block_combine_norm.restype = C.c_int
    L.fm_scan_fwd_workspace_bytes.argtypes = [C.POINTER(FmScanFwdParams)]
    L.fm_scan_fwd_workspace_bytes.restype = C.c_int64
    if L.fm_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libfm_scan.so ABI {L.fm_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().fm_last_error().decode("utf-8", "replace")
        raise RuntimeError(msg or f"{what} failed with status {rc}")


def launch_count() -> int:
    return int(lib().fm_launch_count())

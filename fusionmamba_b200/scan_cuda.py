"""Host-side mirror of the reference's pybind module ``selective_scan_cuda`` (``fwd`` / ``bwd``).

Same positional signatures, argument meaning, output lists and error behaviour as
selective_scan/selective_scan.cpp:226-336 (``fwd``) and :338-492 (``bwd``); the work is done by the
hand-written sm_100a kernels behind the C ABI in include/fm_scan.h.  PyTorch is used only for device
memory and the current stream.  ``compat/selective_scan_cuda.py`` re-exports these two functions under the
reference's module name so that ``import selective_scan_cuda`` in unmodified model code binds to them.

Scope notes: complex ``A`` and constant (2-D) ``B``/``C`` are never
exercised by FusionMamba and are rejected with a RuntimeError instead of being silently mis-computed.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib

CHUNK_LEN = 2048  # checkpoint spacing of x, as in the reference (selective_scan.cpp:307)
HCK_LEN = 64      # spacing of the dense state checkpoints the backward kernel starts its chunks from


HCK_LEN_16 = 8    # dstate == 16: the lane-serial backward kernels rebuild 8 steps at a time in registers
LS_MAX_SEQLEN = int(__import__("os").environ.get("FM_SCAN_BWD_LS_MAXL", "512"))
# longer sequences: lane-serial only when its (8 rows x whole sequence) warps fill the machine -- 5 per SM of a B200
LS_MIN_UNITS = int(__import__("os").environ.get("FM_SCAN_BWD_LS_MINUNITS", "740"))


def _hck_len(dstate: int, seqlen: int = 0, rows: int = 0) -> int:
    """Spacing of the dense state checkpoints, which also selects the backward kernel (the C ABI dispatches on hck_len):
    8 -> lane-serial kernels (fm_scan_bwd_ls2.cuh / fm_scan_bwd_ls.cuh): sequences up to ~512 steps (the short-L stages of the
    model: 1.2-4x, profiles/r02_bwd_ls_ab.jsonl) and longer ones whose ``rows`` (batch * dim) give at least LS_MIN_UNITS
    8-row warps (profiles/r02_ls2_ab.jsonl); 64 -> row-pair kernel (fm_scan_bwd_rp.cuh), time-parallel, for few long rows, with
    8x less checkpoint traffic.  Very wide states (dstate > 64, never used by FusionMamba) use 16-step chunks."""
    if dstate == 16 and seqlen > 0 and (seqlen <= LS_MAX_SEQLEN or rows // 8 >= LS_MIN_UNITS):
        return HCK_LEN_16
    return HCK_LEN if dstate <= 64 else 16

# Dense checkpoints ride in the SAME storage as ``x``, after its (batch, dim, n_chunks, 2*dstate) payload:
# ``x`` stays a contiguous tensor of exactly the reference's shape (so it can be saved / passed around like
# the reference's x), and ``bwd`` recognises its own buffers by the exact storage size.  A foreign or
# copied ``x`` simply lacks the tail; ``bwd`` then recomputes the checkpoints with one extra forward launch.
stats = {"hck_fast": 0, "hck_recomputed": 0}


def _n_hck(seqlen: int, dstate: int, rows: int = 0) -> int:
    hl = _hck_len(dstate, seqlen, rows)
    return (seqlen + hl - 1) // hl - 1


def _alloc_x(batch, dim, seqlen, dstate, device, with_hck):
    n_chunks = (seqlen + CHUNK_LEN - 1) // CHUNK_LEN
    nx = batch * dim * n_chunks * 2 * dstate
    n_hck = _n_hck(seqlen, dstate, batch * dim)
    nh = batch * dim * n_hck * dstate if with_hck else 0
    buf = torch.empty(nx + nh, device=device, dtype=torch.float32)
    x = buf[:nx].view(batch, dim, n_chunks, 2 * dstate)
    hck = buf[nx:].view(batch, dim, n_hck, dstate) if nh > 0 else None
    return x, hck


def _hck_of(x: torch.Tensor, batch, dim, seqlen, dstate):
    """Recover the hidden checkpoint tail of an ``x`` produced by :func:`fwd` (None if it is not one of ours)."""
    n_chunks = (seqlen + CHUNK_LEN - 1) // CHUNK_LEN
    nx = batch * dim * n_chunks * 2 * dstate
    n_hck = _n_hck(seqlen, dstate, batch * dim)
    nh = batch * dim * n_hck * dstate
    if nh == 0 or x.storage_offset() != 0 or not x.is_contiguous():
        return None
    if x.untyped_storage().nbytes() != 4 * (nx + nh):
        return None
    return torch.as_strided(x, (batch, dim, n_hck, dstate), (n_hck * dstate * dim, n_hck * dstate, dstate, 1), nx)

_DT = {torch.float32: _lib.FM_F32, torch.float16: _lib.FM_F16, torch.bfloat16: _lib.FM_BF16}


def _check(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _validate(u, delta, A, B, C_, D_, z_, delta_bias_, who):
    _check(u.dtype in _DT, f"{who}: input dtype must be float32, float16 or bfloat16")
    _check(not A.is_complex(), f"{who}: complex A is not supported by fusionmamba_b200 (never used by FusionMamba)")
    _check(A.dtype == torch.float32, f"{who}: A must be float32")
    _check(B.dim() >= 3 and C_.dim() >= 3,
           f"{who}: constant (dim, dstate) B/C are not supported by fusionmamba_b200; pass (batch, groups, dstate, L)")
    _check(B.dim() == 4 and C_.dim() == 4, f"{who}: B and C must be 4-D (batch, n_groups, dstate, seqlen)")
    _check(delta.dtype == u.dtype and B.dtype == u.dtype and C_.dtype == u.dtype,
           f"{who}: delta, B, C must have the dtype of u")
    for name, t in (("u", u), ("delta", delta), ("A", A), ("B", B), ("C", C_)):
        _check(t.is_cuda, f"{who}: {name} must be a CUDA tensor")
    _check(u.dim() == 3, f"{who}: u must be (batch, dim, seqlen)")
    batch, dim, seqlen = u.shape
    dstate = A.shape[1]
    n_groups = B.shape[1]
    _check(u.stride(-1) == 1 or seqlen == 1, f"{who}: u must be contiguous in the last dimension")
    _check(delta.stride(-1) == 1 or seqlen == 1, f"{who}: delta must be contiguous in the last dimension")
    _check(dstate <= 256, "selective_scan only supports state dimension <= 256")
    _check(tuple(delta.shape) == (batch, dim, seqlen), f"{who}: delta must have shape {(batch, dim, seqlen)}")
    _check(tuple(A.shape) == (dim, dstate), f"{who}: A must have shape {(dim, dstate)}")
    _check(tuple(B.shape) == (batch, n_groups, dstate, seqlen),
           f"{who}: B must have shape {(batch, n_groups, dstate, seqlen)}")
    _check(tuple(C_.shape) == (batch, n_groups, dstate, seqlen),
           f"{who}: C must have shape {(batch, n_groups, dstate, seqlen)}")
    _check(B.stride(-1) == 1 or seqlen == 1, f"{who}: B must be contiguous in the last dimension")
    _check(C_.stride(-1) == 1 or seqlen == 1, f"{who}: C must be contiguous in the last dimension")
    _check(dim % n_groups == 0, f"{who}: dim must be a multiple of n_groups")
    for name, t in (("D", D_), ("delta_bias", delta_bias_)):
        if t is not None:
            _check(t.dtype == torch.float32, f"{who}: {name} must be float32")
            _check(t.is_cuda, f"{who}: {name} must be a CUDA tensor")
            _check(tuple(t.shape) == (dim,), f"{who}: {name} must have shape {(dim,)}")
            _check(t.stride(-1) == 1 or dim == 1, f"{who}: {name} must be contiguous")
    if z_ is not None:
        _check(z_.dtype == u.dtype and z_.is_cuda, f"{who}: z must be a CUDA tensor with the dtype of u")
        _check(tuple(z_.shape) == (batch, dim, seqlen), f"{who}: z must have shape {(batch, dim, seqlen)}")
        _check(z_.stride(-1) == 1 or seqlen == 1, f"{who}: z must be contiguous in the last dimension")
    return batch, dim, seqlen, dstate, n_groups


def _fill_fwd(p, u, delta, A, B, C_, D_, z_, delta_bias_, out, out_z, x, delta_softplus,
              batch, dim, seqlen, dstate, n_groups):
    p.abi_version = _lib.ABI_VERSION
    p.dtype = _DT[u.dtype]
    p.out_dtype = _DT[out.dtype] if out is not None else _DT[u.dtype]
    p.reserved0 = 0
    p.batch, p.dim, p.seqlen, p.dstate, p.n_groups = batch, dim, seqlen, dstate, n_groups
    p.chunk_len = CHUNK_LEN
    p.n_chunks = (seqlen + CHUNK_LEN - 1) // CHUNK_LEN
    p.delta_softplus = int(bool(delta_softplus))
    p.u_map = p.out_map = _lib.FM_MAP_LINEAR
    p.map_h = p.map_w = 0
    p.u_batch_stride, p.u_d_stride = u.stride(0), u.stride(1)
    p.delta_batch_stride, p.delta_d_stride = delta.stride(0), delta.stride(1)
    if z_ is not None:
        p.z_batch_stride, p.z_d_stride = z_.stride(0), z_.stride(1)
    if out is not None:
        p.out_batch_stride, p.out_d_stride = out.stride(0), out.stride(1)
    if out_z is not None:
        p.out_z_batch_stride, p.out_z_d_stride = out_z.stride(0), out_z.stride(1)
    p.A_d_stride, p.A_dstate_stride = A.stride(0), A.stride(1)
    p.B_batch_stride, p.B_group_stride, p.B_dstate_stride = B.stride(0), B.stride(1), B.stride(2)
    p.C_batch_stride, p.C_group_stride, p.C_dstate_stride = C_.stride(0), C_.stride(1), C_.stride(2)
    p.u, p.delta, p.A, p.B, p.C = _ptr(u), _ptr(delta), _ptr(A), _ptr(B), _ptr(C_)
    p.D, p.z, p.delta_bias = _ptr(D_), _ptr(z_), _ptr(delta_bias_)
    p.out, p.out_z, p.x = _ptr(out), _ptr(out_z), _ptr(x)
    p.hck, p.hck_len, p.n_hck = None, 0, 0


def _set_hck(p, hck, seqlen, dstate):
    if hck is not None:                                          # hck: (batch, dim, n_hck, dstate)
        rows = hck.shape[0] * hck.shape[1]
        p.hck, p.hck_len, p.n_hck = _ptr(hck), _hck_len(dstate, seqlen, rows), _n_hck(seqlen, dstate, rows)


def fwd(u: torch.Tensor, delta: torch.Tensor, A: torch.Tensor, B: torch.Tensor, C: torch.Tensor,
        D_: Optional[torch.Tensor], z_: Optional[torch.Tensor], delta_bias_: Optional[torch.Tensor],
        delta_softplus: bool) -> List[torch.Tensor]:
    """selective_scan_cuda.fwd -> [out, x] or [out, x, out_z]   (selective_scan.cpp:226-336)."""
    batch, dim, seqlen, dstate, n_groups = _validate(u, delta, A, B, C, D_, z_, delta_bias_, "selective_scan_fwd")
    p, outs = prepare_fwd(u, delta, A, B, C, D_, z_, delta_bias_, delta_softplus,
                          dims=(batch, dim, seqlen, dstate, n_groups))
    launch_fwd(p, u.device)
    return outs


def _set_workspace(p, device):
    """Scratch for the time-split forward (few rows, long sequence; FmScanFwdParams.workspace): allocated here with torch --
    the library itself never allocates -- when the C side says the shape can use it.  Returns the tensor to keep alive."""
    need = int(_lib.lib().fm_scan_fwd_workspace_bytes(C.byref(p)))
    if need <= 0:
        return None
    ws = torch.empty((need + 3) // 4, device=device, dtype=torch.float32)
    p.workspace, p.workspace_bytes = C.c_void_p(ws.data_ptr()), need
    return ws


def prepare_fwd(u, delta, A, B, C_, D_, z_, delta_bias_, delta_softplus, dims=None, with_hck=None):
    """Allocate the outputs and fill the C-ABI record for one forward launch -> (params, [out, x, (out_z)])."""
    if dims is None:
        dims = _validate(u, delta, A, B, C_, D_, z_, delta_bias_, "selective_scan_fwd")
    batch, dim, seqlen, dstate, n_groups = dims
    out = torch.empty_like(delta)                      # inherits delta's layout, selective_scan.cpp:311
    if with_hck is None:
        # dense checkpoints are only worth writing when a backward can follow (activations require grad)
        with_hck = any(t is not None and t.requires_grad for t in (u, delta, A, B, C_, z_))
    x, hck = _alloc_x(batch, dim, seqlen, dstate, u.device, with_hck)
    out_z = torch.empty_like(z_) if z_ is not None else None
    p = _lib.FmScanFwdParams()
    _fill_fwd(p, u, delta, A, B, C_, D_, z_, delta_bias_, out, out_z, x, delta_softplus,
              batch, dim, seqlen, dstate, n_groups)
    _set_hck(p, hck, seqlen, dstate)
    ws = _set_workspace(p, u.device)
    p._keep = (u, delta, A, B, C_, D_, z_, delta_bias_, out, out_z, x, ws, hck)   # keep device memory alive with the record
    return p, ([out, x] if z_ is None else [out, x, out_z])


def fwd_merge_v2(u, delta, A, B, C_, D_, delta_bias_, delta_softplus: bool, H: int, W: int, out_dtype=None,
                 channels_last: bool = False) -> torch.Tensor:
    """Inference-only forward with EfficientMerge fused into the store (FmScanFwdParams.out_map = EFFICIENT_V2):
    u, delta (batch, 4*D, L) with L = ceil(H/2)*ceil(W/2) -> y (batch, D, H*W); the (batch, 4, D, L) scan output of
    models/cross.py:323-328 is never materialised.  ``out_dtype=torch.float32`` with 16-bit inputs reproduces the reference's
    "upcast, scan in fp32, keep y in fp32" exactly without the cast copies.  No checkpoints are written, so it cannot be
    followed by ``bwd``."""
    # same fix-ups as the autograd wrapper (selective_scan_interface.py:25-36): only the last dimension must be contiguous
    u, delta, B, C_ = (t if t.stride(-1) == 1 else t.contiguous() for t in (u, delta, B, C_))
    batch, dim, seqlen, dstate, n_groups = _validate(u, delta, A, B, C_, D_, None, delta_bias_, "selective_scan_fwd")
    _check(n_groups == 4 and dim % 4 == 0, "selective_scan_fwd: fused merge needs 4 scan directions")
    # channels_last: y (batch, H*W, D) -- contiguous runs per pixel for the store, and LayerNorm needs no transpose
    y = torch.empty((batch, H * W, dim // 4) if channels_last else (batch, dim // 4, H * W), device=u.device,
                    dtype=out_dtype or u.dtype)
    x, _ = _alloc_x(batch, dim, seqlen, dstate, u.device, False)
    p = _lib.FmScanFwdParams()
    _fill_fwd(p, u, delta, A, B, C_, D_, None, delta_bias_, y, None, x, delta_softplus, batch, dim, seqlen, dstate, n_groups)
    p.out_map, p.map_h, p.map_w = (_lib.FM_MAP_EFFICIENT_V2_CL if channels_last else _lib.FM_MAP_EFFICIENT_V2), H, W
    p._keep = (u, delta, A, B, C_, D_, delta_bias_, y, x, _set_workspace(p, u.device))
    launch_fwd(p, u.device)
    return y


def launch_fwd(p, device) -> None:
    """One asynchronous forward launch on the current stream of ``device`` (C ABI: fm_selective_scan_fwd)."""
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().fm_selective_scan_fwd(C.byref(p), C.c_void_p(stream)), "fm_selective_scan_fwd")


def bwd(u: torch.Tensor, delta: torch.Tensor, A: torch.Tensor, B: torch.Tensor, C: torch.Tensor,
        D_: Optional[torch.Tensor], z_: Optional[torch.Tensor], delta_bias_: Optional[torch.Tensor],
        dout: torch.Tensor, x_: Optional[torch.Tensor], out_: Optional[torch.Tensor],
        dz_: Optional[torch.Tensor], delta_softplus: bool, recompute_out_z: bool) -> List[Optional[torch.Tensor]]:
    """selective_scan_cuda.bwd -> [du, ddelta, dA, dB, dC, dD, ddelta_bias, (dz), (out_z)]  (selective_scan.cpp:338-492)."""
    p, r = prepare_bwd(u, delta, A, B, C, D_, z_, delta_bias_, dout, x_, out_, dz_, delta_softplus, recompute_out_z)
    launch_bwd(p, u.device)
    result = [r["du"], r["ddelta"], r["dA"], r["dB"].to(B.dtype), r["dC"].to(C.dtype), r["dD"], r["ddelta_bias"]]
    if z_ is not None:
        result.append(r["dz"])
    if recompute_out_z:
        result.append(r["out_z"])
    return result


def prepare_bwd(u, delta, A, B, C, D_, z_, delta_bias_, dout, x_, out_, dz_, delta_softplus, recompute_out_z):
    """Validate, allocate (zero-initialised accumulators) and fill the C-ABI record for one backward launch."""
    who = "selective_scan_bwd"
    batch, dim, seqlen, dstate, n_groups = _validate(u, delta, A, B, C, D_, z_, delta_bias_, who)
    _check(dout.dtype == u.dtype and dout.is_cuda, f"{who}: dout must be a CUDA tensor with the dtype of u")
    _check(tuple(dout.shape) == (batch, dim, seqlen), f"{who}: dout must have shape {(batch, dim, seqlen)}")
    _check(dout.stride(-1) == 1 or seqlen == 1, f"{who}: dout must be contiguous in the last dimension")
    n_chunks = (seqlen + CHUNK_LEN - 1) // CHUNK_LEN
    out = out_z = dz = None
    if z_ is not None:
        _check(out_ is not None, f"{who}: out must be given when z is given")
        out = out_
        _check(out.dtype == u.dtype and out.is_cuda and tuple(out.shape) == (batch, dim, seqlen),
               f"{who}: out must be a CUDA tensor of shape {(batch, dim, seqlen)} with the dtype of u")
        _check(out.stride(-1) == 1 or seqlen == 1, f"{who}: out must be contiguous in the last dimension")
        if dz_ is not None:
            dz = dz_
            _check(dz.dtype == u.dtype and dz.is_cuda and tuple(dz.shape) == (batch, dim, seqlen),
                   f"{who}: dz must be a CUDA tensor of shape {(batch, dim, seqlen)} with the dtype of u")
            _check(dz.stride(-1) == 1 or seqlen == 1, f"{who}: dz must be contiguous in the last dimension")
        else:
            dz = torch.empty_like(z_)
        if recompute_out_z:
            out_z = torch.empty_like(out)
    if n_chunks > 1:
        _check(x_ is not None, f"{who}: x must be given when seqlen > {CHUNK_LEN}")
    if x_ is not None:
        _check(x_.dtype == torch.float32 and x_.is_cuda and x_.is_contiguous(), f"{who}: x must be contiguous float32 CUDA")
        _check(tuple(x_.shape) == (batch, dim, n_chunks, 2 * dstate),
               f"{who}: x must have shape {(batch, dim, n_chunks, 2 * dstate)}")
    hck = None
    if _n_hck(seqlen, dstate, batch * dim) > 0:
        hck = _hck_of(x_, batch, dim, seqlen, dstate) if x_ is not None else None
        if hck is not None:
            stats["hck_fast"] += 1
        else:
            # x did not come from our fwd (or was copied): rebuild the checkpoints with one forward launch
            stats["hck_recomputed"] += 1
            pf, _ = prepare_fwd(u, delta, A, B, C, D_, None, delta_bias_, delta_softplus,
                                dims=(batch, dim, seqlen, dstate, n_groups), with_hck=True)
            launch_fwd(pf, u.device)
            hck = pf._keep[-1]
    du = torch.empty_like(u)
    ddelta = torch.empty_like(delta)
    dA = torch.zeros_like(A, memory_format=torch.contiguous_format)
    dB = torch.zeros(B.shape, device=B.device, dtype=torch.float32)   # fp32 accumulators, selective_scan.cpp:461-462
    dC = torch.zeros(C.shape, device=C.device, dtype=torch.float32)
    dD = torch.zeros_like(D_) if D_ is not None else None
    ddelta_bias = torch.zeros_like(delta_bias_) if delta_bias_ is not None else None

    p = _lib.FmScanBwdParams()
    _fill_fwd(p.f, u, delta, A, B, C, D_, z_, delta_bias_, out, out_z, x_, delta_softplus,
              batch, dim, seqlen, dstate, n_groups)
    if out is None:
        p.f.out = None
    _set_hck(p.f, hck, seqlen, dstate)
    p.dout_batch_stride, p.dout_d_stride = dout.stride(0), dout.stride(1)
    p.du_batch_stride, p.du_d_stride = du.stride(0), du.stride(1)
    p.ddelta_batch_stride, p.ddelta_d_stride = ddelta.stride(0), ddelta.stride(1)
    if dz is not None:
        p.dz_batch_stride, p.dz_d_stride = dz.stride(0), dz.stride(1)
    p.dB_batch_stride, p.dB_group_stride, p.dB_dstate_stride = dB.stride(0), dB.stride(1), dB.stride(2)
    p.dC_batch_stride, p.dC_group_stride, p.dC_dstate_stride = dC.stride(0), dC.stride(1), dC.stride(2)
    p.dout, p.du, p.ddelta, p.dz = _ptr(dout), _ptr(du), _ptr(ddelta), _ptr(dz)
    p.dA, p.dB, p.dC, p.dD, p.ddelta_bias = _ptr(dA), _ptr(dB), _ptr(dC), _ptr(dD), _ptr(ddelta_bias)
    p._keep = (u, delta, A, B, C, D_, z_, delta_bias_, dout, x_, out, out_z, hck, du, ddelta, dA, dB, dC, dD, ddelta_bias, dz)
    return p, dict(du=du, ddelta=ddelta, dA=dA, dB=dB, dC=dC, dD=dD, ddelta_bias=ddelta_bias, dz=dz, out_z=out_z)


def launch_bwd(p, device) -> None:
    """One asynchronous backward launch on the current stream of ``device`` (C ABI: fm_selective_scan_bwd).
    dA, dB, dC, dD, ddelta_bias accumulate: the caller zeroes them before the launch."""
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().fm_selective_scan_bwd(C.byref(p), C.c_void_p(stream)), "fm_selective_scan_bwd")

"""ctypes wrapper of oracle/scan_oracle.c (TEST INFRASTRUCTURE ONLY; see the header of that file).

Used by tests/ (parity at sizes the numpy oracle would take minutes for), ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  Never imported by fusionmamba_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libscan_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "scan_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"] + (["-B"] if force else []), check=True)
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.fm_oracle_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().fm_oracle_num_threads())


def set_threads(n: int) -> None:
    """Override OMP_NUM_THREADS (torchrun exports 1) for the timed CPU baseline."""
    lib().fm_oracle_set_threads(C.c_int(int(n)))


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _norm(u, delta, A, B, Cm, D, z, bias):
    u, delta, A, B, Cm, D, z, bias = map(_f32, (u, delta, A, B, Cm, D, z, bias))
    sq = B.ndim == 3
    if sq:
        B, Cm = B[:, None], Cm[:, None]
    B, Cm = np.ascontiguousarray(B), np.ascontiguousarray(Cm)
    batch, dim, L = u.shape
    return u, delta, A, B, Cm, D, z, bias, batch, dim, L, A.shape[1], B.shape[1], sq


def scan_fwd(u, delta, A, B, Cm, D=None, z=None, delta_bias=None, delta_softplus=False):
    """-> (out float64 (batch, dim, L), y_pre float64, last_state float64 (batch, dim, N))."""
    u, delta, A, B, Cm, D, z, bias, batch, dim, L, N, G, _ = _norm(u, delta, A, B, Cm, D, z, delta_bias)
    out = np.empty((batch, dim, L)); ypre = np.empty((batch, dim, L)); last = np.empty((batch, dim, N))
    rc = lib().fm_oracle_scan_fwd(batch, dim, L, N, G, _p(u), _p(delta), _p(A), _p(B), _p(Cm), _p(D), _p(z), _p(bias),
                                  int(bool(delta_softplus)), _p(out, C.c_double), _p(ypre, C.c_double),
                                  _p(last, C.c_double))
    if rc:
        raise ValueError("fm_oracle_scan_fwd: bad arguments")
    return out, ypre, last


def scan_bwd(u, delta, A, B, Cm, D, z, delta_bias, dout, delta_softplus=False):
    u, delta, A, B, Cm, D, z, bias, batch, dim, L, N, G, sq = _norm(u, delta, A, B, Cm, D, z, delta_bias)
    dout = _f32(dout)
    du = np.empty((batch, dim, L)); dd = np.empty((batch, dim, L)); dA = np.empty((dim, N))
    dB = np.empty((batch, G, N, L)); dC = np.empty((batch, G, N, L))
    dD = np.empty(dim) if D is not None else None
    db = np.empty(dim) if bias is not None else None
    dz = np.empty((batch, dim, L)) if z is not None else None
    d = C.c_double
    rc = lib().fm_oracle_scan_bwd(batch, dim, L, N, G, _p(u), _p(delta), _p(A), _p(B), _p(Cm), _p(D), _p(z), _p(bias),
                                  int(bool(delta_softplus)), _p(dout), _p(du, d), _p(dd, d), _p(dA, d), _p(dB, d),
                                  _p(dC, d), _p(dD, d), _p(db, d), _p(dz, d))
    if rc:
        raise ValueError("fm_oracle_scan_bwd: bad arguments or out of memory")
    if sq:
        dB, dC = dB[:, 0], dC[:, 0]
    return dict(du=du, ddelta=dd, dA=dA, dB=dB, dC=dC, dD=dD, ddelta_bias=db, dz=dz)

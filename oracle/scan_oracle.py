"""CPU oracle for the SS2D hot path (TEST INFRASTRUCTURE ONLY -- never a product path).

This file restates, in numpy (float64 accumulation), the algorithm of the reference's
pure-PyTorch selective scan and of its scan-unfold / scan-merge permutations.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker (or the timed CPU baseline) -- the product path
(``fusionmamba_b200``) must never route through it.

Reference anchors (paths relative to the upstream repository root):
  * ``selective_scan_ref``      mamba_ssm/ops/selective_scan_interface.py:92-158
  * gradients                   autograd of the same function; closed form in SURVEY.md section 3.5,
                                kernel restatement selective_scan/selective_scan_bwd_kernel.cuh:146-477
  * ``EfficientScan``           models/cross.py:139-190      (stride-2 "v2" unfold)
  * ``EfficientMerge``          models/cross.py:34-88        (stride-2 "v2" merge, pure permutation)
  * CrossScan / CrossMerge      models/cross.py:610-612, 639-642 and
                                models/vmamba_Fusion_efficross.py:398-400, 425-429  ("v0", 4-way sum)

Parity pinning: ``tests/golden/*.npz`` were produced by importing the reference's own
``selective_scan_ref`` (forward + autograd backward) and ``EfficientScan``/``EfficientMerge``
in the build container (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks
this oracle against every one of those fixtures.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "softplus",
    "selective_scan_fwd",
    "selective_scan_bwd",
    "efficient_scan",
    "efficient_merge",
    "cross_scan_v0",
    "cross_merge_v0",
    "efficient_scan_index",
    "cross_scan_v0_index",
]


def softplus(x: np.ndarray) -> np.ndarray:
    """F.softplus(beta=1, threshold=20): x if x > 20 else log1p(exp(x)).

    selective_scan_interface.py:113 (F.softplus) and the CUDA kernel's
    ``delta <= 20 ? log1pf(expf(delta)) : delta`` (selective_scan_fwd_kernel.cuh:153-156).
    """
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _prep(u, delta, A, B, C, D, z, delta_bias, delta_softplus):
    u = np.asarray(u, dtype=np.float64)
    delta = np.asarray(delta, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    batch, dim, L = u.shape
    N = A.shape[1]
    squeeze_B = B.ndim == 3
    squeeze_C = C.ndim == 3
    if squeeze_B:
        B = B[:, None]
    if squeeze_C:
        C = C[:, None]
    G = B.shape[1]
    assert dim % G == 0 and C.shape[1] == G
    x = delta + (np.asarray(delta_bias, dtype=np.float64)[None, :, None] if delta_bias is not None else 0.0)
    dt = softplus(x) if delta_softplus else x
    return u, delta, A, B, C, x, dt, batch, dim, L, N, G, squeeze_B, squeeze_C


def selective_scan_fwd(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                       return_last_state=False, return_y=False):
    """Restatement of ``selective_scan_ref`` (selective_scan_interface.py:92-158), real A only.

    u, delta, z: (batch, dim, L); A: (dim, N); B, C: (batch, G, N, L) or (batch, N, L);
    D, delta_bias: (dim,).  Returns float64 ``out`` (caller casts to the I/O dtype) and optionally
    ``last_state`` (batch, dim, N) and the pre-gate ``y`` (what the CUDA ``fwd`` returns as ``out``
    when ``z`` is given, selective_scan.cpp:334-336).
    """
    u, delta, A, B, C, x, dt, batch, dim, L, N, G, _, _ = _prep(u, delta, A, B, C, D, z, delta_bias, delta_softplus)
    H = dim // G
    h = np.zeros((batch, dim, N))
    y = np.empty((batch, dim, L))
    for t in range(L):
        a = np.exp(dt[:, :, t, None] * A[None])                                   # :127
        Bt = np.repeat(B[:, :, :, t], H, axis=1)                                  # :134  (B G N -> B (G H) N)
        Ct = np.repeat(C[:, :, :, t], H, axis=1)                                  # :137
        h = a * h + (dt[:, :, t] * u[:, :, t])[:, :, None] * Bt                   # :135, :140
        y[:, :, t] = np.einsum("bdn,bdn->bd", h, Ct)                              # :147
    if D is not None:
        y = y + u * np.asarray(D, dtype=np.float64)[None, :, None]                # :154
    out = y
    if z is not None:
        zz = np.asarray(z, dtype=np.float64)
        out = y * zz * _sigmoid(zz)                                               # :155-156
    res = [out]
    if return_last_state:
        res.append(h)
    if return_y:
        res.append(y)
    return res[0] if len(res) == 1 else tuple(res)


def selective_scan_bwd(u, delta, A, B, C, D, z, delta_bias, dout, delta_softplus=False):
    """Closed-form gradients of ``selective_scan_ref`` (what autograd computes for it).

    Returns dict with du, ddelta, dA, dB, dC, dD, ddelta_bias, dz (entries None when the input
    was None).  Formulas: SURVEY.md section 3.5; kernel statement selective_scan_bwd_kernel.cuh:171-207
    (z path), :277-296 (per-state products), :439-452 (softplus derivative).
    """
    u, delta, A, B, C, x, dt, batch, dim, L, N, G, squeeze_B, squeeze_C = _prep(
        u, delta, A, B, C, D, z, delta_bias, delta_softplus)
    H = dim // G
    g = np.asarray(dout, dtype=np.float64)
    Dv = np.asarray(D, dtype=np.float64) if D is not None else None
    # forward states (kept: small-case oracle)
    hs = np.empty((L, batch, dim, N))
    h = np.zeros((batch, dim, N))
    a_all = np.empty((L, batch, dim, N))
    for t in range(L):
        a = np.exp(dt[:, :, t, None] * A[None])
        Bt = np.repeat(B[:, :, :, t], H, axis=1)
        h = a * h + (dt[:, :, t] * u[:, :, t])[:, :, None] * Bt
        hs[t] = h
        a_all[t] = a
    y = np.einsum("tbdn,bdnt->bdt", hs, np.repeat(C, H, axis=1))
    if Dv is not None:
        y = y + u * Dv[None, :, None]
    dz = None
    if z is not None:
        zz = np.asarray(z, dtype=np.float64)
        sg = _sigmoid(zz)
        dz = g * y * sg * (1.0 + zz * (1.0 - sg))
        dy = g * zz * sg
    else:
        dy = g
    du = np.zeros_like(u)
    ddt = np.zeros_like(u)
    dA = np.zeros_like(A)
    dB = np.zeros((batch, G, N, L))
    dC = np.zeros((batch, G, N, L))
    dh = np.zeros((batch, dim, N))
    for t in range(L - 1, -1, -1):
        Bt = np.repeat(B[:, :, :, t], H, axis=1)
        Ct = np.repeat(C[:, :, :, t], H, axis=1)
        a_next = a_all[t + 1] if t + 1 < L else 0.0
        dh = Ct * dy[:, :, t, None] + a_next * dh
        hprev = hs[t - 1] if t > 0 else 0.0
        gk = a_all[t] * hprev                      # h_t - b_t
        s1 = np.einsum("bdn,bdn->bd", dh, Bt)
        du[:, :, t] = dt[:, :, t] * s1
        w = dh * gk
        ddt[:, :, t] = u[:, :, t] * s1 + np.einsum("bdn,dn->bd", w, A)
        dA += np.einsum("bdn,bd->dn", w, dt[:, :, t])
        dB[:, :, :, t] = (dh * (dt[:, :, t] * u[:, :, t])[:, :, None]).reshape(batch, G, H, N).sum(axis=2)
        dC[:, :, :, t] = (hs[t] * dy[:, :, t, None]).reshape(batch, G, H, N).sum(axis=2)
    dD = None
    if Dv is not None:
        du += dy * Dv[None, :, None]
        dD = np.einsum("bdt,bdt->d", dy, u)
    if delta_softplus:
        ddelta = np.where(x > 20.0, ddt, ddt * _sigmoid(x))
    else:
        ddelta = ddt
    dbias = ddelta.sum(axis=(0, 2)) if delta_bias is not None else None
    if squeeze_B:
        dB = dB[:, 0]
    if squeeze_C:
        dC = dC[:, 0]
    return dict(du=du, ddelta=ddelta, dA=dA, dB=dB, dC=dC, dD=dD, ddelta_bias=dbias, dz=dz)


# ------------------------------------------------------------------------------------------------
# scan unfold / merge permutations (bit-exact integer index work)
# ------------------------------------------------------------------------------------------------

def efficient_scan_index(H: int, W: int, step: int = 2):
    """Index map of EfficientScan.forward (models/cross.py:139-169).

    Returns ``idx`` int64 (4, Lp) with Lp = ceil(H/2)*ceil(W/2): xs[b,k,d,l] = x[b,d].flat[idx[k,l]]
    when idx >= 0, and 0 (the F.pad value, :147-153) when idx == -1.
    """
    assert step == 2, "the reference hard-codes 4 sub-grids (step_size=2) in its K=4 layout"
    Hp, Wp = math.ceil(H / 2), math.ceil(W / 2)
    idx = np.full((4, Hp * Wp), -1, dtype=np.int64)
    i = np.arange(Hp)[:, None]
    j = np.arange(Wp)[None, :]

    def put(k, l, h, w):
        ok = (h < H) & (w < W)
        flat = np.where(ok, h * W + w, -1)
        idx[k, l.ravel()] = np.broadcast_to(flat, l.shape).ravel()

    put(0, i * Wp + j, 2 * i, 2 * j)              # x[:, :, ::2, ::2]                         :163
    put(1, j * Hp + i, 2 * i + 1, 2 * j)          # x.transpose(2,3)[:, :, ::2, 1::2]         :164
    put(2, i * Wp + j, 2 * i, 2 * j + 1)          # x[:, :, ::2, 1::2]                        :165
    put(3, j * Hp + i, 2 * i + 1, 2 * j + 1)      # x.transpose(2,3)[:, :, 1::2, 1::2]        :166
    return idx


def efficient_scan(x: np.ndarray, step: int = 2) -> np.ndarray:
    """EfficientScan.forward: (B, C, H, W) -> (B, 4, C, ceil(H/2)*ceil(W/2)); zero pad for odd sizes."""
    Bn, Cn, H, W = x.shape
    idx = efficient_scan_index(H, W, step)
    flat = np.concatenate([x.reshape(Bn, Cn, H * W), np.zeros((Bn, Cn, 1), dtype=x.dtype)], axis=2)
    return np.stack([flat[:, :, idx[k]] for k in range(4)], axis=1)


def efficient_merge(ys: np.ndarray, H: int, W: int, step: int = 2) -> np.ndarray:
    """EfficientMerge.forward (models/cross.py:34-58): (B, 4, C, Lp) -> (B, C, H*W); inverse permutation, crops pad."""
    Bn, K, Cn, Lp = ys.shape
    idx = efficient_scan_index(H, W, step)
    y = np.zeros((Bn, Cn, H * W + 1), dtype=ys.dtype)
    for k in range(4):
        y[:, :, idx[k]] = ys[:, k]           # idx == -1 lands in the scratch slot
    return y[:, :, : H * W]


def cross_scan_v0_index(H: int, W: int):
    """Index map of the classic CrossScan (models/cross.py:610-612): idx (4, L), L = H*W."""
    L = H * W
    l = np.arange(L)
    idx = np.empty((4, L), dtype=np.int64)
    idx[0] = l                                   # x.view(B, -1, L)
    idx[1] = (l % H) * W + (l // H)              # transpose(2,3).contiguous().view  -> l = w*H + h
    idx[2] = idx[0][::-1]                        # flip of the stacked pair
    idx[3] = idx[1][::-1]
    return idx


def cross_scan_v0(x: np.ndarray) -> np.ndarray:
    Bn, Cn, H, W = x.shape
    idx = cross_scan_v0_index(H, W)
    flat = x.reshape(Bn, Cn, H * W)
    return np.stack([flat[:, :, idx[k]] for k in range(4)], axis=1)


def cross_merge_v0(out_y: np.ndarray, H: int, W: int) -> np.ndarray:
    """CrossMerge (models/cross.py:639-642): y = out_y[:,0] + inv_y[:,0] + wh_y + invwh_y in THAT order."""
    Bn, K, Cn, L = out_y.shape
    idx = cross_scan_v0_index(H, W)
    inv = [np.empty(L, dtype=np.int64) for _ in range(4)]
    for k in range(4):
        inv[k][idx[k]] = np.arange(L)            # position l of pixel p in direction k
    y = out_y[:, 0][:, :, inv[0]]
    y = y + out_y[:, 2][:, :, inv[2]]
    y = y + out_y[:, 1][:, :, inv[1]]
    y = y + out_y[:, 3][:, :, inv[3]]
    return y

"""Python-level operator API of the selective scan -- same names, signatures and autograd contract as the
reference's ``mamba_ssm.ops.selective_scan_interface`` (selective_scan_interface.py:20-158):

    selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False)
    selective_scan_ref(... same signature ...)      # pure-PyTorch definition of the op (API parity only)

``selective_scan_fn`` runs the sm_100a kernels through ``fusionmamba_b200.scan_cuda`` (C ABI); it never falls
back to ``selective_scan_ref``.  Autograd contract kept from the reference (:25-80): inputs whose last
stride is not 1 are made contiguous, 3-D B/C are treated as one group and their gradients squeezed back,
dB/dC come back in the input dtype, dD / ddelta_bias are None when the input was None, and the gradient
of ``last_state`` is ignored.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import scan_cuda


def _last_contig(t):
    return t if (t is None or t.stride(-1) == 1) else t.contiguous()


class SelectiveScanFn(torch.autograd.Function):
    """Autograd wrapper over scan_cuda.fwd / scan_cuda.bwd (reference: selective_scan_interface.py:20-80)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                return_last_state=False):
        u, delta, B, C, z = map(_last_contig, (u, delta, B, C, z))
        if D is not None:
            D = D.contiguous()
        ctx.squeeze_B = B.dim() == 3
        ctx.squeeze_C = C.dim() == 3
        if ctx.squeeze_B:
            B = B.unsqueeze(1)
        if ctx.squeeze_C:
            C = C.unsqueeze(1)
        res = scan_cuda.fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus)
        out, x = res[0], res[1]
        ctx.delta_softplus = delta_softplus
        ctx.has_z = z is not None
        last_state = x[:, :, -1, 1::2]  # (batch, dim, dstate)
        if ctx.has_z:
            ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, x, out)
            result = res[2]
        else:
            ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
            result = out
        return (result, last_state) if return_last_state else result

    @staticmethod
    def backward(ctx, dout, *ignored):
        if ctx.has_z:
            u, delta, A, B, C, D, z, delta_bias, x, out = ctx.saved_tensors
        else:
            u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
            z = out = None
        dout = _last_contig(dout)
        res = scan_cuda.bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out, None, ctx.delta_softplus, False)
        du, ddelta, dA, dB, dC, dD, ddelta_bias = res[:7]
        dz = res[7] if ctx.has_z else None
        if ctx.squeeze_B:
            dB = dB.squeeze(1)
        if ctx.squeeze_C:
            dC = dC.squeeze(1)
        return (du, ddelta, dA, dB, dC,
                dD if D is not None else None,
                dz,
                ddelta_bias if delta_bias is not None else None,
                None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """If ``return_last_state`` the result is ``(out, last_state)`` with last_state (batch, dim, dstate) fp32;
    its gradient is not propagated (reference docstring, selective_scan_interface.py:85-88)."""
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                       return_last_state=False):
    """Pure-PyTorch statement of the op with the reference's signature (selective_scan_interface.py:92-158).

    Kept for API parity (``from mamba_ssm.ops.selective_scan_interface import selective_scan_ref`` in
    models/cross.py:16) and for users who want a differentiable definition on any device.  Real ``A`` and
    time-varying ``B``/``C`` of shape (batch, L-last) only.  Not used by any fusionmamba_b200 code path.
    """
    in_dtype = u.dtype
    u32, dt = u.float(), delta.float()
    if delta_bias is not None:
        dt = dt + delta_bias.float().unsqueeze(-1)
    if delta_softplus:
        dt = F.softplus(dt)
    if A.is_complex() or B.dim() < 3 or C.dim() < 3:
        raise NotImplementedError("selective_scan_ref: only real A with time-varying B and C is provided")
    batch, dim, L = u32.shape
    N = A.shape[1]
    Bf = B.float() if B.dim() == 4 else B.float().unsqueeze(1)
    Cf = C.float() if C.dim() == 4 else C.float().unsqueeze(1)
    reps = dim // Bf.shape[1]
    Bf = Bf.repeat_interleave(reps, dim=1)              # (batch, dim, N, L)
    Cf = Cf.repeat_interleave(reps, dim=1)
    decay = torch.exp(dt.unsqueeze(-1) * A.float()[None, :, None, :])                  # (batch, dim, L, N)
    drive = (dt * u32).unsqueeze(-1) * Bf.transpose(2, 3)               # (batch, dim, L, N)
    state = u32.new_zeros(batch, dim, N)
    ys = []
    for t in range(L):
        state = decay[:, :, t] * state + drive[:, :, t]
        ys.append((state * Cf[:, :, :, t]).sum(-1))
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + u32 * D.float().unsqueeze(-1)
    if z is not None:
        y = y * F.silu(z.float())
    y = y.to(in_dtype)
    return (y, state) if return_last_state else y

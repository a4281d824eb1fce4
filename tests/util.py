"""Shared helpers for the parity tests (inputs follow mamba_ssm/ops/test_selective_scan.py:406-441, 474)."""
import numpy as np
import torch

# Tolerances (BASELINE.json north_star): rtol 1e-4 for fp32, 2e-2 for bf16 on outputs, last_state and every
# gradient; fp16 uses the reference test's 3e-3 (test_selective_scan.py:398).  atol is stated relative to the
# largest reference magnitude of the tensor being compared: |a - r| <= atol_rel * max|r| + rtol * |r|.
TOL = {
    torch.float32: dict(rtol=1e-4, atol_rel=2e-5),
    torch.float16: dict(rtol=3e-3, atol_rel=2e-3),
    torch.bfloat16: dict(rtol=2e-2, atol_rel=1e-2),
}


def make_inputs(batch, dim, L, N, G, itype, has_D=True, has_z=False, has_bias=True, seed=0, device="cuda",
                model_init=False, squeeze=False):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=gen)
    rn = lambda *s: torch.randn(*s, generator=gen)
    if model_init:  # models/cross.py:556-595
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(dim, 1)
        dt = torch.exp(r(dim) * (np.log(0.1) - np.log(1e-3)) + np.log(1e-3)).clamp(min=1e-4)
        bias = dt + torch.log(-torch.expm1(-dt))
        D = torch.ones(dim)
    else:
        A = -0.5 * r(dim, N)
        bias = 0.5 * r(dim)
        D = rn(dim)
    shp = (batch, N, L) if squeeze else (batch, G, N, L)
    d = dict(
        u=rn(batch, dim, L).to(itype), delta=(0.5 * r(batch, dim, L)).to(itype), A=A.float(),
        B=rn(*shp).to(itype), C=rn(*shp).to(itype),
        D=D.float() if has_D else None, z=rn(batch, dim, L).to(itype) if has_z else None,
        delta_bias=bias.float() if has_bias else None, g=rn(batch, dim, L).to(itype))
    return {k: (v.to(device) if v is not None else None) for k, v in d.items()}


def to_np(t):
    return None if t is None else t.detach().float().cpu().numpy()


def assert_close(actual, ref, dtype, name="", rtol_mul=1.0, atol_mul=1.0):
    """|a - r| <= atol_rel*max|r| + rtol*|r| elementwise; returns the worst normalised error for reporting."""
    tol = TOL[dtype]
    a = np.asarray(to_np(actual) if torch.is_tensor(actual) else actual, dtype=np.float64)
    r = np.asarray(ref, dtype=np.float64)
    assert a.shape == r.shape, f"{name}: shape {a.shape} vs {r.shape}"
    assert np.isfinite(a).all(), f"{name}: non-finite values"
    scale = max(float(np.abs(r).max()), 1e-30)
    bound = tol["atol_rel"] * atol_mul * scale + tol["rtol"] * rtol_mul * np.abs(r)
    err = np.abs(a - r)
    worst = float((err / bound).max())
    assert worst <= 1.0, (f"{name}: max |err| {err.max():.3e} (ref scale {scale:.3e}); worst err/bound {worst:.2f} "
                          f"with rtol {tol['rtol'] * rtol_mul:g}, atol {tol['atol_rel'] * atol_mul:g}*max|ref|")
    return worst

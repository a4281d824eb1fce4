"""Generate SS2D-level golden fixtures by running the REFERENCE's own model code on CPU (build container only).

    python tests/golden/make_golden_ss2d.py

The reference's ``SS2D`` (forward_type v2 and v0) and ``SS2D_cross_new`` modules (models/cross.py:417-742, 890-1230) and its
``cross_selective_scan`` are instantiated unmodified; their ``selective_scan_cuda.fwd/bwd`` boundary (models/cross.py:119,
130-133) and ``selective_scan_fn`` are served by the reference's pure-PyTorch ``selective_scan_ref`` + autograd
(mamba_ssm/ops/selective_scan_interface.py:92-158), i.e. everything in the fixture is the reference's own arithmetic.
Each fixture holds the module's state_dict, the inputs, the upstream gradient, the output and the gradients of the inputs and of
every parameter.  tests/test_ss2d_gpu.py loads the state_dict into fusionmamba_b200.ss2d modules and compares on the B200.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _load_reference  # noqa: E402


def _install_scan_stub(iface):
    """selective_scan_cuda.fwd / .bwd computed by the reference's selective_scan_ref (+ autograd for bwd)."""
    ssc = sys.modules["selective_scan_cuda"]

    def fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus):
        with torch.no_grad():
            out, last = iface.selective_scan_ref(u, delta, A, B, C, D, z=z, delta_bias=delta_bias,
                                                 delta_softplus=delta_softplus, return_last_state=True)
        x = torch.zeros(u.shape[0], u.shape[1], 1, 2 * A.shape[1])
        x[:, :, 0, 1::2] = last
        return [out, x]

    def bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out, dz, delta_softplus, recompute_out_z):
        ins = [t.detach().clone().requires_grad_() if t is not None else None for t in (u, delta, A, B, C, D, delta_bias)]
        with torch.enable_grad():
            o = iface.selective_scan_ref(ins[0], ins[1], ins[2], ins[3], ins[4], ins[5], z=None, delta_bias=ins[6],
                                         delta_softplus=delta_softplus)
            grads = torch.autograd.grad(o, [t for t in ins if t is not None], dout)
        it = iter(grads)
        return [next(it) if t is not None else None for t in ins]

    ssc.fwd, ssc.bwd = fwd, bwd

    def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
        return iface.selective_scan_ref(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)
    return selective_scan_fn


def _np(t):
    return t.detach().float().numpy()


def module_case(cross, name, cls, kwargs, shapes, seed, randomize=True, core_only=False):
    """core_only: call ``forward_corev0(x, channel_first=True)`` on x (B, D, H, W) instead of ``forward`` -- the reference's
    SS2D.forward passes ``step_size=`` to forward_corev0, which does not take it (models/cross.py:739 vs :598), so the
    v0 core is only reachable directly."""
    torch.manual_seed(seed)
    m = cls(**kwargs)
    if randomize:  # move A_logs / Ds off their constant init so that per-channel values matter
        with torch.no_grad():
            m.A_logs.add_(0.3 * torch.randn_like(m.A_logs))
            m.Ds.add_(0.5 * torch.randn_like(m.Ds))
            m.out_norm.weight.add_(0.2 * torch.randn_like(m.out_norm.weight))
            m.out_norm.bias.add_(0.2 * torch.randn_like(m.out_norm.bias))
    xs = [torch.randn(*s, requires_grad=True) for s in shapes]
    out = m.forward_corev0(xs[0], channel_first=True) if core_only else m(*xs)
    g = torch.randn_like(out)
    out.backward(g)
    d = {"out": _np(out), "g": _np(g)}
    for i, x in enumerate(xs):
        d[f"x{i}"] = _np(x)
        d[f"dx{i}"] = _np(x.grad)
    for k, v in m.state_dict().items():
        d["sd/" + k] = _np(v)
    for k, p in m.named_parameters():
        if p.grad is not None:
            d["grad/" + k] = _np(p.grad)
    d["kwargs"] = np.array(repr(kwargs))
    np.savez_compressed(os.path.join(HERE, f"ss2d_{name}.npz"), **d)
    print("wrote", name, tuple(out.shape), "params", sum(p.numel() for p in m.parameters()))


def core_case(cross, name, B, D, H, W, N, R, seed):
    """cross_selective_scan alone (no module): models/cross.py:266-337, bf16-free fp32 case with odd sizes."""
    torch.manual_seed(seed)
    x = torch.randn(B, D, H, W, requires_grad=True)
    xw = (0.3 * torch.randn(4, R + 2 * N, D)).requires_grad_()
    dw = (0.3 * torch.randn(4, D, R)).requires_grad_()
    db = (0.5 * torch.rand(4, D)).requires_grad_()
    Al = torch.log(0.5 + torch.rand(4 * D, N)).requires_grad_()
    Ds = torch.randn(4 * D, requires_grad=True)
    norm = torch.nn.LayerNorm(D)
    y = cross.cross_selective_scan(x, xw, None, dw, db, Al, Ds, norm, nrows=1, delta_softplus=True, step_size=2)
    g = torch.randn_like(y)
    y.backward(g)
    d = dict(x=_np(x), x_proj_weight=_np(xw), dt_projs_weight=_np(dw), dt_projs_bias=_np(db), A_logs=_np(Al), Ds=_np(Ds),
             norm_weight=_np(norm.weight), norm_bias=_np(norm.bias), y=_np(y), g=_np(g), dx=_np(x.grad),
             dx_proj_weight=_np(xw.grad), ddt_projs_weight=_np(dw.grad), ddt_projs_bias=_np(db.grad), dA_logs=_np(Al.grad),
             dDs=_np(Ds.grad), dnorm_weight=_np(norm.weight.grad), dnorm_bias=_np(norm.bias.grad))
    np.savez_compressed(os.path.join(HERE, f"ss2d_{name}.npz"), **d)
    print("wrote", name, tuple(y.shape))


def main():
    iface, cross = _load_reference()
    cross.selective_scan_fn = _install_scan_stub(iface)   # forward_corev0 binds the module-level name (models/cross.py:602)
    core_case(cross, "core_v2_odd", B=2, D=8, H=7, W=5, N=16, R=2, seed=11)
    core_case(cross, "core_v2_even", B=1, D=12, H=8, W=12, N=16, R=3, seed=12)
    module_case(cross, "mod_v2", cross.SS2D, dict(d_model=16, d_state=16, ssm_ratio=2.0, dt_rank="auto", d_conv=3, forward_type="v2"),
                [(2, 8, 8, 16)], seed=13)
    module_case(cross, "mod_v2_odd", cross.SS2D, dict(d_model=24, d_state=16, ssm_ratio=2.0, dt_rank="auto", d_conv=3, forward_type="v2"),
                [(1, 9, 7, 24)], seed=14)
    module_case(cross, "mod_v0", cross.SS2D, dict(d_model=16, d_state=16, ssm_ratio=2.0, dt_rank="auto", d_conv=3, forward_type="v0"),
                [(2, 32, 6, 5)], seed=15, core_only=True)
    module_case(cross, "mod_cross", cross.SS2D_cross_new, dict(d_model=16, d_state=16, ssm_ratio=2.0, dt_rank="auto", d_conv=3, forward_type="v2"),
                [(2, 8, 6, 16), (2, 8, 6, 16)], seed=16)


if __name__ == "__main__":
    main()

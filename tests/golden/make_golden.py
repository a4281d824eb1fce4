"""Generate golden fixtures by running the REFERENCE's own Python code (build container only).

Run once in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Writes small ``.npz`` fixtures next to this file.  They pin ``oracle/`` (tests/test_oracle_golden.py)
and the CUDA path (tests/test_scan_gpu.py) to the reference's results:
  * ``selective_scan_ref`` forward + autograd backward   mamba_ssm/ops/selective_scan_interface.py:92-158
  * ``EfficientScan`` / ``EfficientMerge`` forward+backward   models/cross.py:34-88, 139-190
  * classic CrossScan / CrossMerge expressions              models/cross.py:610-612, 639-642
Inputs follow the reference test's distributions (mamba_ssm/ops/test_selective_scan.py:406-441, 474).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("FM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _load_reference():
    # The reference imports its CUDA extension unconditionally (selective_scan_interface.py:16,
    # models/cross.py:17); a dummy module satisfies the import -- nothing here calls into it.
    sys.modules.setdefault("selective_scan_cuda", types.ModuleType("selective_scan_cuda"))
    spec = importlib.util.spec_from_file_location(
        "ref_selective_scan_interface", os.path.join(REF, "mamba_ssm/ops/selective_scan_interface.py"))
    iface = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(iface)

    # stubs for the model file's third-party imports (timm absent; mamba_ssm package import is broken
    # under transformers 5.x -- see SURVEY.md section 8c)
    timm = types.ModuleType("timm"); timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__(); self.drop_prob = p

        def forward(self, x):
            return x
    timm_layers.DropPath = DropPath
    timm_layers.to_2tuple = lambda v: (v, v)
    timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})
    ms = types.ModuleType("mamba_ssm"); ms.Mamba = object
    ms_ops = types.ModuleType("mamba_ssm.ops")
    sys.modules.update({"mamba_ssm": ms, "mamba_ssm.ops": ms_ops,
                        "mamba_ssm.ops.selective_scan_interface": iface})
    spec = importlib.util.spec_from_file_location("ref_models_cross", os.path.join(REF, "models/cross.py"))
    cross = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cross)
    return iface, cross


def scan_case(iface, name, batch, dim, L, N, G, itype, has_D, has_z, has_bias, softplus, seed, squeeze=False,
              model_init=False):
    torch.random.manual_seed(seed)
    if model_init:
        # models/cross.py:556-595 (A = -[1..N], D = 1, delta_bias = softplus^-1(exp(U(log 1e-3, log 1e-1))))
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(dim, 1)
        dt = torch.exp(torch.rand(dim) * (np.log(0.1) - np.log(1e-3)) + np.log(1e-3)).clamp(min=1e-4)
        bias0 = dt + torch.log(-torch.expm1(-dt))
    else:
        A = -0.5 * torch.rand(dim, N, dtype=torch.float32)
        bias0 = 0.5 * torch.rand(dim, dtype=torch.float32)
    A.requires_grad_()
    shp = (batch, N, L) if squeeze else (batch, G, N, L)
    B = torch.randn(*shp, dtype=itype, requires_grad=True)
    C = torch.randn(*shp, dtype=itype, requires_grad=True)
    D = (torch.ones(dim) if model_init else torch.randn(dim)).float().requires_grad_() if has_D else None
    z = torch.randn(batch, dim, L, dtype=itype, requires_grad=True) if has_z else None
    bias = bias0.clone().requires_grad_() if has_bias else None
    u = torch.randn(batch, dim, L, dtype=itype, requires_grad=True)
    delta = (0.5 * torch.rand(batch, dim, L, dtype=itype)).requires_grad_()
    out, last = iface.selective_scan_ref(u, delta, A, B, C, D, z=z, delta_bias=bias,
                                         delta_softplus=softplus, return_last_state=True)
    g = torch.randn_like(out)
    out.backward(g)
    f = lambda t: None if t is None else t.detach().float().numpy()
    d = dict(u=f(u), delta=f(delta), A=f(A), B=f(B), C=f(C), g=f(g), out=f(out), last_state=f(last),
             du=f(u.grad), ddelta=f(delta.grad), dA=f(A.grad), dB=f(B.grad), dC=f(C.grad),
             delta_softplus=np.array(softplus), itype=np.array(str(itype).replace("torch.", "")))
    if has_D:
        d.update(D=f(D), dD=f(D.grad))
    if has_z:
        d.update(z=f(z), dz=f(z.grad))
    if has_bias:
        d.update(delta_bias=f(bias), ddelta_bias=f(bias.grad))
    np.savez_compressed(os.path.join(HERE, f"scan_{name}.npz"), **d)
    print("wrote", name, {k: v.shape for k, v in d.items() if hasattr(v, "shape") and v.ndim})


def perm_cases(cross):
    out = {}
    for (H, W) in [(4, 4), (6, 8), (5, 7), (7, 4), (1, 1), (3, 1), (2, 9), (8, 8)]:
        torch.manual_seed(H * 100 + W)
        Bn, Cn = 2, 3
        x = torch.randn(Bn, Cn, H, W, requires_grad=True)
        xs = cross.EfficientScan.apply(x, 2)
        gxs = torch.randn_like(xs)
        xs.backward(gxs)
        ys = torch.randn(Bn, 4, Cn, xs.shape[-1], requires_grad=True)
        y = cross.EfficientMerge.apply(ys, H, W, 2)
        gy = torch.randn_like(y)
        y.backward(gy)
        # classic v0 expressions, verbatim structure of models/cross.py:610-612 and :639-642
        L = H * W
        xd = x.detach()
        x_hwwh = torch.stack([xd.view(Bn, -1, L), torch.transpose(xd, 2, 3).contiguous().view(Bn, -1, L)], dim=1)
        xs0 = torch.cat([x_hwwh, torch.flip(x_hwwh, dims=[-1])], dim=1)
        oy = torch.randn(Bn, 4, Cn, L)
        inv_y = torch.flip(oy[:, 2:4], dims=[-1]).view(Bn, 2, -1, L)
        wh_y = torch.transpose(oy[:, 1].view(Bn, -1, W, H), 2, 3).contiguous().view(Bn, -1, L)
        invwh_y = torch.transpose(inv_y[:, 1].view(Bn, -1, W, H), 2, 3).contiguous().view(Bn, -1, L)
        y0 = oy[:, 0] + inv_y[:, 0] + wh_y + invwh_y
        key = f"{H}x{W}"
        for nm, t in dict(x=x, xs=xs, gxs=gxs, gx=x.grad, ys=ys, y=y, gy=gy, gys=ys.grad,
                          xs0=xs0, oy=oy, y0=y0).items():
            out[f"{key}/{nm}"] = t.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "perm_cases.npz"), **out)
    print("wrote perm_cases", len(out))


def main():
    iface, cross = _load_reference()
    f32, bf16, f16 = torch.float32, torch.bfloat16, torch.float16
    #          name            b  dim  L    N   G  itype D     z      bias   sp     seed
    scan_case(iface, "f32_base",      2, 8, 64, 16, 2, f32, True, False, True, True, 0)
    scan_case(iface, "f32_z",         2, 8, 96, 16, 2, f32, True, True, True, True, 1)
    scan_case(iface, "f32_plain",     1, 4, 37, 16, 1, f32, False, False, False, False, 2)
    scan_case(iface, "f32_squeeze",   2, 6, 50, 8, 1, f32, True, False, True, True, 3, squeeze=True)
    scan_case(iface, "f32_n1",        2, 8, 130, 1, 2, f32, True, False, True, False, 4)
    scan_case(iface, "f32_long",      1, 8, 700, 16, 4, f32, True, False, True, True, 5)
    scan_case(iface, "f32_modelinit", 2, 8, 256, 16, 4, f32, True, False, True, True, 6, model_init=True)
    scan_case(iface, "bf16_base",     2, 8, 64, 16, 2, bf16, True, False, True, True, 7)
    scan_case(iface, "bf16_z",        2, 8, 72, 16, 2, bf16, True, True, True, True, 8)
    scan_case(iface, "f16_base",      2, 8, 64, 16, 2, f16, True, False, True, True, 9)
    perm_cases(cross)


if __name__ == "__main__":
    main()

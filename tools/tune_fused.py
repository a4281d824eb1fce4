"""Sweep the dstate-16 forward kernel's launch shape on the SS2D inference call (16-bit inputs, fp32 channels-last fused-merge
output) at a FusionMamba stage shape."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import scan_cuda  # noqa: E402
from tools.bench_vs_ref_cuda import timeit  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
itype = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float32
hw, dm = [(64, 96), (32, 192), (16, 384), (8, 768)][stage]
B, D, H, W, N = 32, 2 * dm, hw, hw, 16
L = (H // 2) * (W // 2)
torch.manual_seed(0)
u = torch.randn(B, 4 * D, L, device="cuda").to(itype)
delta = (0.5 * torch.rand(B, 4 * D, L, device="cuda")).to(itype)
A = -0.5 * torch.rand(4 * D, N, device="cuda")
x_dbl = torch.randn(B, 4, 6 + 2 * N, L, device="cuda").to(itype)
_, Bs, Cs = torch.split(x_dbl, [6, N, N], dim=2)
Dp, bias = torch.randn(4 * D, device="cuda"), 0.5 * torch.rand(4 * D, device="cuda")
for cfg in (sys.argv[3] if len(sys.argv) > 3 else "0x0x0,2x4x1,2x2x1,2x4x2,4x4x1,4x2x1,4x4x2,4x2x2,2x8x1").split(","):
    spl, nw, kt = cfg.split("x")
    os.environ["FM_SCAN_FWD16_SPL"], os.environ["FM_SCAN_FWD16_NW"], os.environ["FM_SCAN_FWD16_KT"] = spl, nw, kt
    for cl in (True, False):
        try:
            f = lambda: scan_cuda.fwd_merge_v2(u, delta, A, Bs, Cs, Dp, bias, True, H, W, out_dtype=torch.float32, channels_last=cl)
            with torch.no_grad():
                us = timeit(f, 20)
            print(json.dumps({"stage": stage, "in": str(itype)[6:], "cfg": cfg, "channels_last": cl, "us": round(us, 1)}), flush=True)
        except RuntimeError as e:
            print(json.dumps({"cfg": cfg, "error": str(e)[:80]}), flush=True)

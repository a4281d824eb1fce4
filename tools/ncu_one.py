"""One forward + one backward launch on a chosen shape (target for ncu captures on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import scan_cuda  # noqa: E402
from tools.bench_vs_ref_cuda import SHAPES  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "configs1"
dt = sys.argv[2] if len(sys.argv) > 2 else "f32"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
itype = torch.float32 if dt == "f32" else torch.bfloat16
Bn, dim, L, N, G = SHAPES[shape]
torch.manual_seed(0)
u = torch.randn(Bn, dim, L, device="cuda").to(itype).requires_grad_()
delta = (0.5 * torch.rand(Bn, dim, L, device="cuda")).to(itype)
A = -0.5 * torch.rand(dim, N, device="cuda")
Bm = torch.randn(Bn, G, N, L, device="cuda").to(itype); Cm = torch.randn(Bn, G, N, L, device="cuda").to(itype)
D = torch.randn(dim, device="cuda"); bias = 0.5 * torch.rand(dim, device="cuda")
g = torch.randn(Bn, dim, L, device="cuda").to(itype)
pf, (out, x) = scan_cuda.prepare_fwd(u, delta, A, Bm, Cm, D, None, bias, True)
pb, r = scan_cuda.prepare_bwd(u.detach(), delta, A, Bm, Cm, D, None, bias, g, x, None, None, True, False)
for _ in range(reps):
    scan_cuda.launch_fwd(pf, u.device)
    scan_cuda.launch_bwd(pb, u.device)
torch.cuda.synchronize()
print("ok")

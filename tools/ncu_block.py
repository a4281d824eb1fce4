"""A few SS2D inference forwards at one stage shape (target for ncu captures of the prologue / epilogue kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import ss2d  # noqa: E402

STAGES = [(64, 96), (32, 192), (16, 384), (8, 768)]
si = int(sys.argv[1]) if len(sys.argv) > 1 else 0
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
hw, dm = STAGES[si]
torch.manual_seed(0)
m = ss2d.SS2D(d_model=dm, d_state=16).cuda().eval()
x = torch.randn(batch, hw, hw, dm, device="cuda")
with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
    for _ in range(reps):
        y = m(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape))

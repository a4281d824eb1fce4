"""torch.profiler breakdown of one SS2D block forward (inference, bf16 autocast) at a FusionMamba stage shape."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import ss2d  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
hw, dm = [(64, 96), (32, 192), (16, 384), (8, 768)][stage]
torch.manual_seed(0)
m = ss2d.SS2D(d_model=dm, d_state=16).cuda().eval()
x = torch.randn(32, hw, hw, dm, device="cuda")


def run():
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        return m(x)


for _ in range(5):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))

"""torch.profiler breakdown of one SS2D block forward+backward (training, fp32) at a FusionMamba stage shape."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import ss2d  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
amp = len(sys.argv) > 2 and sys.argv[2] == "amp"
hw, dm = [(64, 96), (32, 192), (16, 384), (8, 768)][stage]
torch.manual_seed(0)
m = ss2d.SS2D(d_model=dm, d_state=16).cuda().train()
x = torch.randn(32, hw, hw, dm, device="cuda")


def run():
    xx = x.detach().requires_grad_()
    m.zero_grad(set_to_none=True)
    if amp:
        with torch.autocast("cuda", torch.bfloat16):
            y = m(xx)
    else:
        y = m(xx)
    y.float().sum().backward()


for _ in range(5):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))

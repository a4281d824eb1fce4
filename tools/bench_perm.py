import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import ss2d
from tools.bench_vs_ref_cuda import timeit
for dt in (torch.bfloat16, torch.float32):
    for (B, D, H, W) in [(32, 192, 64, 64), (32, 384, 32, 32), (32, 1536, 8, 8)]:
        x = torch.randn(B, D, H, W, device="cuda").to(dt)
        xs = ss2d.scan_unfold(x)
        tu = timeit(lambda: ss2d.scan_unfold(x), 20)
        tm = timeit(lambda: ss2d.scan_merge(xs, H, W), 20)
        print(dt, (B, D, H, W), "unfold us", round(tu, 1), "merge us", round(tm, 1), flush=True)

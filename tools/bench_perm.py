"""Unfold / merge permutation kernels vs their HBM roofline, on the GPU box.
One JSON line per (map, dtype, shape): algorithmic bytes = tensor read once + tensor written once."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import ss2d  # noqa: E402
from tools.bench_vs_ref_cuda import timeit  # noqa: E402

SHAPES = [(8, 192, 64, 64), (32, 192, 64, 64), (32, 384, 32, 32), (32, 1536, 8, 8)]
for mode, name in ((ss2d.MAP_V2, "efficient_v2"), (ss2d.MAP_V0, "cross_v0")):
    for dt in (torch.float32, torch.bfloat16):
        for (B, D, H, W) in SHAPES:
            x = torch.randn(B, D, H, W, device="cuda").to(dt)
            xs = ss2d.scan_unfold(x, mode)
            tu = timeit(lambda: ss2d.scan_unfold(x, mode), 20)
            tm = timeit(lambda: ss2d.scan_merge(xs, H, W, mode), 20)
            by = (x.numel() + xs.numel()) * x.element_size()
            print(json.dumps({"map": name, "dtype": str(dt).split(".")[-1], "shape": [B, D, H, W], "unfold_us": round(tu, 1),
                              "merge_us": round(tm, 1), "algorithmic_MB": round(by / 1e6, 1),
                              "unfold_GBs": round(by / tu / 1e3, 0), "merge_GBs": round(by / tm / 1e3, 0)}), flush=True)

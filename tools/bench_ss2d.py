"""SS2D-block throughput at the four FusionMamba stage shapes (256x256 input), on the GPU box.

    python tools/bench_ss2d.py [--batch 32] [--iters 20]

Per stage: forward under bf16 autocast + no_grad (inference; EfficientMerge fused into the scan's store) and
forward+backward in fp32 (training), CUDA-event timed; prints one JSON line per (stage, mode) with blocks/s, the share of the
time spent in this library's kernels' launches and the launch count per call.  Weights are random-init (no checkpoints here).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import _lib, ss2d  # noqa: E402
from fusionmamba_b200.graph import GraphedForward  # noqa: E402

STAGES = [(64, 96), (32, 192), (16, 384), (8, 768)]   # (tokens per side, d_model)


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    torch.manual_seed(0)
    for si, (hw, dm) in enumerate(STAGES):
        m = ss2d.SS2D(d_model=dm, d_state=16).cuda()
        x = torch.randn(a.batch, hw, hw, dm, device="cuda")

        def infer():
            with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
                return m(x)

        def train():
            xx = x.detach().requires_grad_()
            m.zero_grad(set_to_none=True)
            m(xx).sum().backward()

        graphed = GraphedForward(m, autocast_dtype=torch.bfloat16)

        def infer_graph():
            return graphed(x)

        for mode, fn in (("infer_bf16_autocast", infer), ("infer_bf16_autocast_cuda_graph", infer_graph), ("train_fp32_fwd_bwd", train)):
            n0 = _lib.launch_count()
            fn()
            torch.cuda.synchronize()
            launches = _lib.launch_count() - n0
            ms = timeit(fn, a.iters)
            print(json.dumps({"stage": si, "tokens": f"{hw}x{hw}", "d_model": dm, "d_inner": 2 * dm, "batch": a.batch, "mode": mode,
                              "ms": round(ms, 3), "blocks_per_s": round(1e3 / ms, 1), "pairs_per_s_per_block": round(a.batch * 1e3 / ms, 1),
                              "fm_kernel_launches": launches}), flush=True)


if __name__ == "__main__":
    main()

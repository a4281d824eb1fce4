"""Kernel-level comparison on the GPU box: our sm_100a kernels vs the reference's CUDA extension built unmodified for
sm_100a (oracle/_ref, see oracle/build_ref.py).  Measurement tooling only (uses oracle/ as comparator).

    python tools/bench_vs_ref_cuda.py [--shapes configs1,stage0,...] [--dtype f32|bf16] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import scan_cuda  # noqa: E402
from oracle import build_ref  # noqa: E402

SHAPES = {  # name: (batch, dim, L, N, groups)
    "configs1": (8, 768, 4096, 16, 4),
    "stage0": (32, 768, 1024, 16, 4),
    "stage1": (32, 1536, 256, 16, 4),
    "stage2": (32, 3072, 64, 16, 4),
    "stage3": (32, 6144, 16, 16, 4),
    "long": (1, 768, 65536, 16, 4),
    # the four stage shapes of ONE 1024x1024 pair through EfficientScan (BASELINE configs[4] in the model)
    "p1024_s0": (1, 768, 16384, 16, 4),
    "p1024_s1": (1, 1536, 4096, 16, 4),
    "p1024_s2": (1, 3072, 1024, 16, 4),
    "p1024_s3": (1, 6144, 256, 16, 4),
    # stage 0 of the BASELINE configs[3] training batch (8 pairs of 512x640 through EfficientScan)
    "train_s0": (8, 768, 5120, 16, 4),
}


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
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="configs1,stage0,stage1,stage2,stage3,long")
    ap.add_argument("--dtype", default="f32,bf16")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    ref = build_ref.load_ref()
    rows = []
    for dt in a.dtype.split(","):
        itype = torch.float32 if dt == "f32" else torch.bfloat16
        es = 4 if dt == "f32" else 2
        for name in a.shapes.split(","):
            Bn, dim, L, N, G = SHAPES[name]
            torch.manual_seed(0)
            u = torch.randn(Bn, dim, L, device="cuda").to(itype).requires_grad_()
            delta = (0.5 * torch.rand(Bn, dim, L, device="cuda")).to(itype)
            A = -0.5 * torch.rand(dim, N, device="cuda")
            Bm = torch.randn(Bn, G, N, L, device="cuda").to(itype)
            Cm = torch.randn(Bn, G, N, L, device="cuda").to(itype)
            D = torch.randn(dim, device="cuda"); bias = 0.5 * torch.rand(dim, device="cuda")
            g = torch.randn(Bn, dim, L, device="cuda").to(itype)
            E, Gg, P = Bn * dim * L, Bn * G * N * L, dim * (N + 2)
            fb, bb = (3 * E + 2 * Gg) * es + 4 * P, (5 * E + 4 * Gg) * es + 8 * P
            pf, (out, x) = scan_cuda.prepare_fwd(u, delta, A, Bm, Cm, D, None, bias, True)
            scan_cuda.launch_fwd(pf, u.device)
            pb, r = scan_cuda.prepare_bwd(u.detach(), delta, A, Bm, Cm, D, None, bias, g, x, None, None, True, False)
            row = {"shape": name, "dtype": dt,
                   "ours_fwd_us": timeit(lambda: scan_cuda.launch_fwd(pf, u.device), a.iters),
                   "ours_bwd_us": timeit(lambda: scan_cuda.launch_bwd(pb, u.device), a.iters)}
            row["ours_fwd_GBs"] = fb / row["ours_fwd_us"] / 1e3
            row["ours_bwd_GBs"] = bb / row["ours_bwd_us"] / 1e3
            if ref is not None:
                ud = u.detach()
                o2, x2 = ref.fwd(ud, delta, A, Bm, Cm, D, None, bias, True)
                row["ref_fwd_us"] = timeit(lambda: ref.fwd(ud, delta, A, Bm, Cm, D, None, bias, True), a.iters)
                # includes the reference's own allocation + zero-fill of dA/dB/dC, as its API does
                row["ref_bwd_us"] = timeit(lambda: ref.bwd(ud, delta, A, Bm, Cm, D, None, bias, g, x2, None, None, True, False), a.iters)
                row["ours_api_bwd_us"] = timeit(lambda: scan_cuda.bwd(ud, delta, A, Bm, Cm, D, None, bias, g, x, None, None, True, False), a.iters)
                row["speedup_fwd"] = row["ref_fwd_us"] / row["ours_fwd_us"]
                row["speedup_bwd"] = row["ref_bwd_us"] / row["ours_api_bwd_us"]
            rows.append(row)
            print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)
    return rows


if __name__ == "__main__":
    main()

"""A/B of the three backward kernels (row-pair on 64-step checkpoints; lane-serial and pipelined lane-serial on 8-step checkpoints)
and of the forward that writes those checkpoints, per shape and dtype (GPU box).  -> JSON lines.

    python tools/r02_ls2_ab.py [--shapes configs1,stage0,...] [--dtypes f32,bf16] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import scan_cuda  # noqa: E402
from tools.bench_vs_ref_cuda import SHAPES, timeit  # noqa: E402

SHAPES = dict(SHAPES, configs1_b16=(16, 768, 4096, 16, 4), **{f"b{n}": (n, 768, 4096, 16, 4) for n in (3, 4, 5, 6, 7)})
ARMS = {  # name: (LS_MAX_SEQLEN, LS_MIN_UNITS, env)
    "row_pair": (0, 1 << 30, {}),
    "lane_serial": (1 << 30, 0, {"FM_SCAN_BWD_LS2": "0"}),
    "lane_serial2": (1 << 30, 0, {}),
    "lane_serial2_nw2": (1 << 30, 0, {"FM_SCAN_BWD_LS2_NW": "2"}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="configs1,stage0,stage1,stage2,stage3,train_s0,p1024_s0,p1024_s1")
    ap.add_argument("--dtypes", default="f32,bf16")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    for shape in a.shapes.split(","):
        for dt in a.dtypes.split(","):
            itype = torch.float32 if dt == "f32" else torch.bfloat16
            Bn, dim, L, N, G = SHAPES[shape]
            torch.manual_seed(0)
            u = torch.randn(Bn, dim, L, device="cuda").to(itype)
            delta = (0.5 * torch.rand(Bn, dim, L, device="cuda")).to(itype)
            A = -0.5 * torch.rand(dim, N, device="cuda")
            Bm = torch.randn(Bn, G, N, L, device="cuda").to(itype)
            Cm = torch.randn(Bn, G, N, L, device="cuda").to(itype)
            D = torch.randn(dim, device="cuda"); bias = 0.5 * torch.rand(dim, device="cuda")
            g = torch.randn(Bn, dim, L, device="cuda").to(itype)
            rec = {"shape": shape, "dims": [Bn, dim, L, N, G], "dtype": dt}
            for arm, (maxl, minu, env) in ARMS.items():
                scan_cuda.LS_MAX_SEQLEN, scan_cuda.LS_MIN_UNITS = maxl, minu
                os.environ.update(env)
                try:
                    pf, (out, x) = scan_cuda.prepare_fwd(u, delta, A, Bm, Cm, D, None, bias, True, with_hck=True)
                    scan_cuda.launch_fwd(pf, u.device)
                    pb, r = scan_cuda.prepare_bwd(u, delta, A, Bm, Cm, D, None, bias, g, x, None, None, True, False)
                    rec[arm] = {"hck_len": int(pf.hck_len), "fwd_us": round(timeit(lambda: scan_cuda.launch_fwd(pf, u.device), a.iters), 1),
                                "bwd_us": round(timeit(lambda: scan_cuda.launch_bwd(pb, u.device), a.iters), 1)}
                    del pf, pb, r, out, x
                except RuntimeError as e:
                    rec[arm] = {"error": str(e)[:100]}
                for k in env:
                    os.environ.pop(k, None)
            print(json.dumps(rec), flush=True)
            del u, delta, Bm, Cm, g
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

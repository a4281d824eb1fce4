"""A/B the scan kernels under different environment knobs on several shapes (GPU box).

    python tools/ab.py --env "" --env "FM_SCAN_BWD_FIXN=0" --shapes configs1,stage0 --dtypes f32,bf16

Every --env is a comma-separated list of NAME=VALUE pairs applied for that arm (the launchers read them per launch).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusionmamba_b200 import scan_cuda  # noqa: E402
from tools.bench_vs_ref_cuda import SHAPES, timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", action="append", default=[])
    ap.add_argument("--shapes", default="configs1")
    ap.add_argument("--dtypes", default="f32")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--which", default="fwd,bwd")
    a = ap.parse_args()
    arms = a.env or [""]
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
            pf, (out, x) = scan_cuda.prepare_fwd(u, delta, A, Bm, Cm, D, None, bias, True)
            scan_cuda.launch_fwd(pf, u.device)
            pb, r = scan_cuda.prepare_bwd(u, delta, A, Bm, Cm, D, None, bias, g, x, None, None, True, False)
            for arm in arms:
                kv = dict(p.split("=", 1) for p in arm.split(",") if p)
                os.environ.update(kv)
                rec = {"shape": shape, "dtype": dt, "env": arm}
                try:
                    if "fwd" in a.which:
                        rec["fwd_us"] = round(timeit(lambda: scan_cuda.launch_fwd(pf, u.device), a.iters), 1)
                    if "bwd" in a.which:
                        rec["bwd_us"] = round(timeit(lambda: scan_cuda.launch_bwd(pb, u.device), a.iters), 1)
                except RuntimeError as e:
                    rec["error"] = str(e)[:120]
                for k in kv:
                    os.environ.pop(k, None)
                print(json.dumps(rec), flush=True)
            del u, delta, Bm, Cm, g, out, x, pf, pb, r
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

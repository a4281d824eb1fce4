"""Sweep the launch-shape knobs (lanes per row G, warps per CTA NW) of the scan kernels on one shape (GPU box)."""
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
    ap.add_argument("--shape", default="configs1")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--fwd", default="8x16x8x0,8x16x8x2,8x16x4x0,8x16x4x4,8x16x4x5,8x32x8x0,8x32x4x0,8x8x8x0,16x16x8x0")   # S x G x NW x MINB
    ap.add_argument("--bwd", default="16x8x1,16x4x1,32x8x1,32x4x1,8x8x1,8x4x1,16x8x0,16x4x0")        # G x NW x smemred
    ap.add_argument("--fwd16", default="2x4x1,2x2x1,2x8x1,2x1x1,2x4x2,2x2x2,4x2x2,4x4x2,4x1x2,4x2x1,4x4x1")   # SPL x NW x KT (dstate 16 kernel)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    itype = torch.float32 if a.dtype == "f32" else torch.bfloat16
    Bn, dim, L, N, G = SHAPES[a.shape]
    torch.manual_seed(0)
    u = torch.randn(Bn, dim, L, device="cuda").to(itype).requires_grad_()
    delta = (0.5 * torch.rand(Bn, dim, L, device="cuda")).to(itype)
    A = -0.5 * torch.rand(dim, N, device="cuda")
    Bm = torch.randn(Bn, G, N, L, device="cuda").to(itype); Cm = torch.randn(Bn, G, N, L, device="cuda").to(itype)
    D = torch.randn(dim, device="cuda"); bias = 0.5 * torch.rand(dim, device="cuda")
    g = torch.randn(Bn, dim, L, device="cuda").to(itype)
    pf, (out, x) = scan_cuda.prepare_fwd(u, delta, A, Bm, Cm, D, None, bias, True)
    scan_cuda.launch_fwd(pf, u.device)
    pb, r = scan_cuda.prepare_bwd(u.detach(), delta, A, Bm, Cm, D, None, bias, g, x, None, None, True, False)
    os.environ["FM_SCAN_FWD16"] = "1"
    for cfg in [c for c in a.fwd16.split(",") if c]:
        spl, nw, kt = cfg.split("x")
        os.environ["FM_SCAN_FWD16_SPL"], os.environ["FM_SCAN_FWD16_NW"], os.environ["FM_SCAN_FWD16_KT"] = spl, nw, kt
        try:
            f = timeit(lambda: scan_cuda.launch_fwd(pf, u.device), a.iters)
            print(json.dumps({"shape": a.shape, "dtype": a.dtype, "kernel": "fwd16", "SPL": int(spl), "NW": int(nw), "KT": int(kt), "us": round(f, 1)}), flush=True)
        except RuntimeError as e:
            print(json.dumps({"kernel": "fwd16", "cfg": cfg, "error": str(e)[:100]}), flush=True)
    for k in ("SPL", "NW", "KT"):
        os.environ.pop(f"FM_SCAN_FWD16_{k}", None)
    os.environ["FM_SCAN_FWD16"] = "0"
    for cfg in [c for c in a.fwd.split(",") if c]:
        ss, gg, nw, mb = cfg.split("x")
        os.environ["FM_SCAN_FWD_S"], os.environ["FM_SCAN_FWD_G"], os.environ["FM_SCAN_FWD_NW"] = ss, gg, nw
        os.environ["FM_SCAN_FWD_MINB"] = mb
        try:
            f = timeit(lambda: scan_cuda.launch_fwd(pf, u.device), a.iters)
            print(json.dumps({"shape": a.shape, "dtype": a.dtype, "kernel": "fwd", "S": int(ss), "G": int(gg), "NW": int(nw), "MINB": int(mb), "us": round(f, 1)}), flush=True)
        except RuntimeError as e:
            print(json.dumps({"kernel": "fwd", "cfg": cfg, "error": str(e)[:100]}), flush=True)
    for k in ("S", "G", "NW", "MINB"):
        os.environ.pop(f"FM_SCAN_FWD_{k}", None)
    os.environ.pop("FM_SCAN_FWD16", None)
    for cfg in [c for c in a.bwd.split(",") if c]:
        gg, nw, sr = cfg.split("x")
        os.environ["FM_SCAN_BWD_G"], os.environ["FM_SCAN_BWD_NW"], os.environ["FM_SCAN_BWD_SMEMRED"] = gg, nw, sr
        try:
            b = timeit(lambda: scan_cuda.launch_bwd(pb, u.device), a.iters)
            print(json.dumps({"shape": a.shape, "dtype": a.dtype, "kernel": "bwd", "G": int(gg), "NW": int(nw), "smemred": int(sr), "us": round(b, 1)}), flush=True)
        except RuntimeError as e:
            print(json.dumps({"kernel": "bwd", "cfg": cfg, "error": str(e)[:100]}), flush=True)


if __name__ == "__main__":
    main()

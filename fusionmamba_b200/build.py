"""Build libfm_scan.so in-tree with plain nvcc for sm_100a (no torch headers -> seconds per file).

    python -m fusionmamba_b200.build [--force]

The shared library lands next to this file (fusionmamba_b200/libfm_scan.so); it is git-ignored but travels
to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libfm_scan.so")
SOURCES = ["fm_api.cu", "fm_scan_fwd.cu", "fm_scan_fwd_f32.cu", "fm_scan_fwd_f16.cu", "fm_scan_fwd_bf16.cu",
           "fm_scan_bwd.cu", "fm_scan_bwd_f32.cu", "fm_scan_bwd_f16.cu", "fm_scan_bwd_bf16.cu", "fm_permute.cu", "fm_norm.cu", "fm_norm_bwd.cu", "fm_block.cu", "fm_conv_unfold.cu", "fm_conv_unfold_bwd.cu", "fm_dt_proj.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _digest() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/fm_scan.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p) and p.endswith((".cu", ".cuh", ".h")):
            h.update(n.encode()); h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB

    def cc(src):
        out = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(out + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log[-6000:]}")
        if verbose:
            print(log)
        return out

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build the REFERENCE's own CUDA selective-scan extension, unmodified, for sm_100a (checker only).

Recipe for ``oracle/_ref/selective_scan_cuda_ref.so``: compiles the reference sources *where they lie*
under ``/root/reference/selective_scan`` (nothing is copied into this repository; outputs go only to
``oracle/_ref/``, which is git-ignored but travels to the GPU box).  It is the on-box GPU comparator
("the kernel to beat") and a second parity witness next to ``selective_scan_ref``.

TEST INFRASTRUCTURE ONLY: nothing under ``fusionmamba_b200/`` imports this module or its output.

Usage (build container; needs /root/reference):   python oracle/build_ref.py
On the GPU box the prebuilt .so is loaded with :func:`load_ref` (returns None when absent).
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "selective_scan_cuda_ref"
REF = os.environ.get("FM_REFERENCE", "/root/reference")


def build(verbose: bool = True, force: bool = False) -> str:
    so = os.path.join(OUT, NAME + ".so")
    if os.path.exists(so) and not force:
        return so
    import torch  # noqa: F401
    from torch.utils import cpp_extension

    src_dir = os.path.join(REF, "selective_scan")
    if not os.path.isdir(src_dir):
        raise FileNotFoundError(f"{src_dir} not found: the reference is only mounted in the build container")
    sources = [os.path.join(src_dir, "selective_scan.cpp")] + sorted(glob.glob(os.path.join(src_dir, "*.cu")))
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    cpp_extension.load(
        name=NAME,
        sources=sources,
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=[
            "-O3", "-std=c++17", "--use_fast_math", "--expt-relaxed-constexpr", "--expt-extended-lambda",
            "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
            "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
            "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
            "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-w",
        ],
        extra_include_paths=[src_dir],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,   # no import here: the build container has no GPU driver
    )
    for f in glob.glob(os.path.join(OUT, "*.o")):   # objects need not travel to the GPU box
        os.remove(f)
    return so


def load_ref():
    """Import the prebuilt reference extension (module exposing ``fwd`` / ``bwd``), or None if absent."""
    path = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[NAME] = mod
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

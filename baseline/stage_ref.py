"""Stage the UNMODIFIED reference application (model shell, loss, pure-PyTorch scan) for the GPU box.

    python baseline/stage_ref.py            # build container only: needs /root/reference

The reference is Python; its "build" is byte-compilation.  Each source file is compiled WHERE IT LIES under
/root/reference (``py_compile``, nothing is copied or edited) and only the resulting sourceless byte-code (``.bytecode``: the gpurun snapshot drops ``*.pyc``) lands in
``baseline/_ref/`` -- git-ignored (``baseline/_ref/``), so no reference source enters this repository's
history, but NOT gpurun-ignored, so it travels to the GPU box next to our own built ``.so`` files.  The GPU box has no
/root/reference; there the prebuilt files are imported as they are (same image, same CPython, same bytecode magic).

What is staged and who uses it:
  models/cross.bytecode, models/vmamba_Fusion_efficross.bytecode   the model shell (VSSM_Fusion, VSSBlock_new, LDC, BiAttn, ...):
        the application that runs "unmodified" on top of our selective_scan_cuda / SS2D path (tools/model_harness.py,
        tests/test_model_gpu.py, bench.py's model record)
  loss.bytecode, pytorch_msssim/__init__.bytecode                Fusionloss for the training-step record (train.py:157)
  refscan/selective_scan_interface.bytecode                 the reference's own CPU path ``selective_scan_ref``
        (mamba_ssm/ops/selective_scan_interface.py:92-158): oracle for the model-level parity test and the
        ``--impl reference`` arm of bench.py (kind "reference")
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("FM_REFERENCE", "/root/reference")

# (source under /root/reference, destination under baseline/_ref)
FILES = [
    ("models/cross.py", "models/cross.bytecode"),
    ("models/vmamba_Fusion_efficross.py", "models/vmamba_Fusion_efficross.bytecode"),
    ("loss.py", "loss.bytecode"),
    ("pytorch_msssim/__init__.py", "pytorch_msssim/__init__.bytecode"),
    ("mamba_ssm/ops/selective_scan_interface.py", "refscan/selective_scan_interface.bytecode"),
]


def manifest_path() -> str:
    return os.path.join(OUT, "MANIFEST.json")


def staged() -> bool:
    return all(os.path.exists(os.path.join(OUT, dst)) for _, dst in FILES)


def build(force: bool = False) -> str:
    if not os.path.isdir(REF):
        if staged():
            return OUT
        raise FileNotFoundError(f"{REF} not found and baseline/_ref is not staged (run this in the build container)")
    man = {"python": sys.version.split()[0], "magic": __import__("importlib.util").util.MAGIC_NUMBER.hex(), "files": {}}
    for src, dst in FILES:
        s, d = os.path.join(REF, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        digest = hashlib.sha256(open(s, "rb").read()).hexdigest()
        man["files"][dst] = {"source": src, "sha256": digest}
        if force or not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s):
            # dfile: the path shown in tracebacks (there is no source on the GPU box to show lines from)
            py_compile.compile(s, cfile=d, dfile=f"<reference>/{src}", doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(manifest_path(), "w") as f:
        json.dump(man, f, indent=1)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

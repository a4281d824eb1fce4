"""Generate the model-level golden fixture (BASELINE configs[0]) by running the REFERENCE's own model on CPU.

    python tests/golden/make_golden_model.py          (build container: needs /root/reference, ~15 s)

The unmodified tiny ``VSSM_Fusion(depths=[2,2,1,2], depths_decoder=[2,1,2,2])`` (models/vmamba_Fusion_efficross.py:753-929,
README.md:54) is built with ``torch.manual_seed(0)`` and run in fp32 on one synthetic 1x1x256x256 pair with every scan
served by the reference's pure-PyTorch ``selective_scan_ref`` (mamba_ssm/ops/selective_scan_interface.py:92-158) -- i.e.
everything in the fixture is the reference's own arithmetic on CPU.  Stored: the fused image (full), a strided sample
and the abs-mean / abs-max of every SS2D output in call order (25 calls), and a fingerprint of the initial weights
(141.9 M parameters cannot be committed; the same seed rebuilds them on any host and the fingerprint proves it).
tests/test_model_gpu.py rebuilds the model on the GPU box and compares the CUDA path with this file.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tools import model_harness as mh  # noqa: E402

SAMPLE = 2048   # values kept per SS2D output


def sample_idx(numel: int) -> np.ndarray:
    step = max(1, numel // SAMPLE)
    return np.arange(0, numel, step, dtype=np.int64)[:SAMPLE]


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    model = mh.build_model("tiny", device="cpu", seed=0).eval()
    fp = mh.weights_fingerprint(model)
    mh.set_backend("cpu_oracle")
    x1, x2 = mh.make_pair(1, 256, 256, seed=0)
    outs = []
    with torch.no_grad(), mh.capture_ss2d_outputs(model, outs):
        y = model(x1, x2)
    d = {"image": y.numpy().astype(np.float32), "n_calls": np.int64(len(outs)),
         "fp_n": np.int64(fp["n_params"]), "fp_sum": np.float64(fp["sum"]), "fp_sum_abs": np.float64(fp["sum_abs"]),
         "x_sum": np.float64(float(x1.double().sum() + 2 * x2.double().sum())),
         "names": np.array([n for n, _ in outs])}
    for i, (_, o) in enumerate(outs):
        flat = o.reshape(-1).numpy()
        d[f"s{i}"] = flat[sample_idx(flat.size)].astype(np.float32)
        d[f"m{i}"] = np.array([np.abs(flat).mean(), np.abs(flat).max(), flat.size], dtype=np.float64)
    out = os.path.join(HERE, "model_tiny_fwd.npz")
    np.savez_compressed(out, **d)
    print(out, os.path.getsize(out), "bytes;", len(outs), "SS2D calls; image abs-mean", float(np.abs(d["image"]).mean()))


if __name__ == "__main__":
    main()

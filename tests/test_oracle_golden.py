"""Pin oracle/ against fixtures produced by the REFERENCE's own code (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import scan_oracle as so

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SCAN_FILES = sorted(glob.glob(os.path.join(GOLD, "scan_*.npz")))


def _tol(itype):
    # fixture values are the reference's fp32 (or half) results; oracle is fp64.
    return dict(float32=(2e-4, 2e-5), bfloat16=(2e-2, 2e-2), float16=(3e-3, 3e-3))[itype]


@pytest.mark.parametrize("path", SCAN_FILES, ids=[os.path.basename(p)[5:-4] for p in SCAN_FILES])
def test_scan_oracle_matches_reference(path):
    d = dict(np.load(path))
    itype = str(d["itype"])
    rtol, atol = _tol(itype)
    sp = bool(d["delta_softplus"])
    out, last = so.selective_scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d.get("D"), d.get("z"),
                                      d.get("delta_bias"), sp, return_last_state=True)
    scale = max(1.0, float(np.abs(d["out"]).max()))
    np.testing.assert_allclose(out, d["out"], rtol=rtol, atol=atol * scale)
    np.testing.assert_allclose(last, d["last_state"], rtol=2e-4, atol=2e-5 * max(1.0, np.abs(last).max()))
    gr = so.selective_scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d.get("D"), d.get("z"),
                               d.get("delta_bias"), d["g"], sp)
    for k in ("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias", "dz"):
        if k not in d:
            assert gr[k] is None or k in ("dD", "ddelta_bias", "dz")
            continue
        ref = d[k]
        sc = max(1.0, float(np.abs(ref).max()))
        np.testing.assert_allclose(gr[k], ref, rtol=rtol, atol=atol * sc, err_msg=k)


def _perm():
    z = np.load(os.path.join(GOLD, "perm_cases.npz"))
    keys = sorted({k.split("/")[0] for k in z.files})
    return z, keys


@pytest.mark.parametrize("key", _perm()[1])
def test_permutations_bit_exact(key):
    z, _ = _perm()
    g = lambda n: z[f"{key}/{n}"]
    H, W = map(int, key.split("x"))
    x = g("x")
    # EfficientScan forward / backward (backward == EfficientMerge-style inverse scatter + crop)
    assert np.array_equal(so.efficient_scan(x), g("xs"))
    assert np.array_equal(so.efficient_merge(g("gxs"), H, W), g("gx").reshape(x.shape[0], x.shape[1], -1))
    # EfficientMerge forward / backward (backward == EfficientScan of the padded gradient)
    assert np.array_equal(so.efficient_merge(g("ys"), H, W), g("y"))
    assert np.array_equal(so.efficient_scan(g("gy").reshape(x.shape)), g("gys"))
    # classic CrossScan / CrossMerge
    assert np.array_equal(so.cross_scan_v0(x), g("xs0"))
    assert np.array_equal(so.cross_merge_v0(g("oy"), H, W), g("y0"))  # same fp32 add order -> bit exact


def test_index_maps_are_permutations():
    for (H, W) in [(2, 2), (64, 64), (5, 3), (128, 160)]:
        idx = so.efficient_scan_index(H, W)
        real = idx[idx >= 0]
        assert np.array_equal(np.sort(real), np.arange(H * W))
        i0 = so.cross_scan_v0_index(H, W)
        for k in range(4):
            assert np.array_equal(np.sort(i0[k]), np.arange(H * W))


@pytest.mark.parametrize("path", SCAN_FILES, ids=[os.path.basename(p)[5:-4] for p in SCAN_FILES])
def test_c_oracle_matches_reference_and_numpy(path):
    """The plain-C oracle (used at full sizes and as the CPU baseline) agrees with the fixtures and the numpy oracle."""
    from oracle import c_oracle
    d = dict(np.load(path))
    itype = str(d["itype"])
    rtol, atol = _tol(itype)
    sp = bool(d["delta_softplus"])
    args = (d["u"], d["delta"], d["A"], d["B"], d["C"], d.get("D"), d.get("z"), d.get("delta_bias"))
    out, ypre, last = c_oracle.scan_fwd(*args, sp)
    out_np, last_np = so.selective_scan_fwd(*args, sp, return_last_state=True)
    np.testing.assert_allclose(out, out_np, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(last, last_np, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(out, d["out"], rtol=rtol, atol=atol * max(1.0, float(np.abs(d["out"]).max())))
    gr = c_oracle.scan_bwd(*args, d["g"], sp)
    gr_np = so.selective_scan_bwd(*args, d["g"], sp)
    for k, v in gr.items():
        if v is None:
            assert gr_np[k] is None
            continue
        np.testing.assert_allclose(v, gr_np[k], rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(gr_np[k]).max())), err_msg=k)
        if k in d:
            np.testing.assert_allclose(v, d[k], rtol=rtol, atol=atol * max(1.0, float(np.abs(d[k]).max())), err_msg=k)

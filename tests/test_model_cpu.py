"""CPU tier of the model-level harness: the staged reference application imports, the tiny VSSM_Fusion rebuilds the
fixture's weights from the seed, the reference's own CPU path (selective_scan_ref) reproduces the committed fixture,
and this library's SS2D modules adopt the reference modules' state_dicts (strict=True).  No CUDA call is made."""
import os

import numpy as np
import pytest
import torch

from tools import model_harness as mh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "model_tiny_fwd.npz")
needs_ref = pytest.mark.skipif(not (mh.available() or os.path.isdir("/root/reference")),
                               reason="reference application not staged (baseline/stage_ref.py)")


@pytest.fixture(scope="module")
def tiny_cpu():
    return mh.build_model("tiny", device="cpu", seed=0).eval()


@needs_ref
def test_tiny_weights_rebuild_from_seed(tiny_cpu):
    gold = np.load(GOLD)
    fp = mh.weights_fingerprint(tiny_cpu)
    assert fp["n_params"] == int(gold["fp_n"]) == 141934037
    assert abs(fp["sum"] - float(gold["fp_sum"])) <= 1e-9 * abs(float(gold["fp_sum"]))
    assert abs(fp["sum_abs"] - float(gold["fp_sum_abs"])) <= 1e-9 * float(gold["fp_sum_abs"])


@needs_ref
def test_reference_cpu_path_reproduces_fixture(tiny_cpu):
    """The oracle of configs[0] (reference model + selective_scan_ref on CPU) against the committed fixture."""
    gold = np.load(GOLD)
    mh.set_backend("cpu_oracle")
    x1, x2 = mh.make_pair(1, 256, 256, seed=0)
    outs = []
    with torch.no_grad(), mh.capture_ss2d_outputs(tiny_cpu, outs):
        y = tiny_cpu(x1, x2)
    mh.set_backend("ours")
    assert len(outs) == int(gold["n_calls"]) == 25
    np.testing.assert_allclose(y.numpy(), gold["image"], rtol=1e-5, atol=1e-5)   # thread-count dependent summation order only
    for i, (name, o) in enumerate(outs):
        assert name == str(gold["names"][i])
        assert abs(float(o.abs().mean()) - gold[f"m{i}"][0]) <= 1e-4 * gold[f"m{i}"][0]


@needs_ref
def test_adopt_reference_modules_state_dict(tiny_cpu):
    import copy
    from fusionmamba_b200 import ss2d
    m = copy.deepcopy(tiny_cpu)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    n = ss2d.adopt_reference_modules(m)
    assert n == 18                                   # 7 shared encoder blocks + 4 cross blocks + 7 decoder blocks
    after = m.state_dict()
    assert list(before) == list(after)               # same keys in the same order: a reference checkpoint still loads
    for k in before:
        assert torch.equal(before[k], after[k]), k
    kinds = {type(mod).__module__ for _, mod in mh.ss2d_modules(m)}
    assert kinds == {"fusionmamba_b200.ss2d"}


def test_ss2d_rejects_unsupported_constructor_arguments():
    from fusionmamba_b200 import ss2d
    with pytest.raises(NotImplementedError):
        ss2d.SS2D(d_model=32, ssm_ratio=2.0, ssm_rank_ratio=1.0)
    with pytest.raises(NotImplementedError):
        ss2d.SS2D(d_model=32, simple_init=True)
    with pytest.warns(UserWarning):
        ss2d.SS2D(d_model=32, not_a_reference_argument=1)

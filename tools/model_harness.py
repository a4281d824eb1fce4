"""Run the UNMODIFIED reference FusionMamba model (VSSM_Fusion) on top of this library.

The reference application is imported from ``baseline/_ref`` (sourceless byte-code staged by baseline/stage_ref.py from
the sources under /root/reference; nothing of it lives in this repository).  Its imports of ``selective_scan_cuda``,
``mamba_ssm`` and ``timm`` (models/cross.py:9-17, models/vmamba_Fusion_efficross.py:12-16) are satisfied by
``fusionmamba_b200.compat.install()``.  The harness then selects who serves the scan boundary:

  backend "ours"        selective_scan_cuda.fwd/bwd = fusionmamba_b200.scan_cuda (the drop-in; model code untouched)
  backend "ref_cuda"    ... = the reference's own CUDA extension rebuilt for sm_100a (oracle/_ref, comparator)
  backend "cpu_oracle"  ... = the reference's pure-PyTorch selective_scan_ref (refscan/, CPU; the oracle of configs[0])

and, on top of backend "ours", the opt-in faster routes of the same path:

  fuse "patch"   ss2d.patch_reference: models.cross.cross_selective_scan(_cross) rebound to fusionmamba_b200.ss2d
  fuse "swap"    every reference SS2D / SS2D_cross_new module replaced by fusionmamba_b200.ss2d's module of the same
                 name, loaded from the reference module's own state_dict (strict=True)

This file is harness code (tests/, bench.py's model record); the library never imports it.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "baseline", "_ref")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TINY = dict(depths=[2, 2, 1, 2], depths_decoder=[2, 1, 2, 2])          # BASELINE configs[0]   (README.md:54)
FULL = dict(depths=[2, 2, 9, 2], depths_decoder=[2, 9, 2, 2])          # BASELINE configs[2..4]


def available() -> bool:
    return os.path.exists(os.path.join(REFDIR, "models", "vmamba_Fusion_efficross.bytecode"))


def _load_pyc(name: str, relpath: str):
    path = os.path.join(REFDIR, relpath)
    loader = importlib.machinery.SourcelessFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod                      # registered before execution, like a regular import
    try:
        loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    return mod


_state: dict = {}


def _ensure_stage():
    if not available():
        from baseline import stage_ref
        stage_ref.build()


def ref_scan_interface():
    """The reference's mamba_ssm/ops/selective_scan_interface.py (selective_scan_ref is pure PyTorch; its selective_scan_fn
    is bound to whatever ``selective_scan_cuda`` is registered when this is first called)."""
    if "iface" not in _state:
        _ensure_stage()
        if "selective_scan_cuda" not in sys.modules:
            sys.modules["selective_scan_cuda"] = types.ModuleType("selective_scan_cuda")
        _state["iface"] = _load_pyc("ref_selective_scan_interface", "refscan/selective_scan_interface.bytecode")
    return _state["iface"]


def load_reference():
    """Import the reference application (models.cross, models.vmamba_Fusion_efficross, loss) once; returns a namespace."""
    if "ref" in _state:
        return _state["ref"]
    _ensure_stage()
    from fusionmamba_b200 import compat
    compat.install()
    # the byte-code files carry a neutral extension (the gpurun snapshot drops *.pyc), so the import system cannot find them by
    # itself: register the reference's package layout by hand -- ``models`` (a namespace package in the reference: no
    # __init__.py), ``models.cross``, ``models.vmamba_Fusion_efficross`` -- in the order the reference imports them
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REFDIR, "models")]
    sys.modules["models"] = pkg
    with _cuda_noop_if_no_gpu():                 # loss.Sobelxy / LDC call .cuda() in __init__ (models/cross.py:798-800)
        cross = _load_pyc("models.cross", "models/cross.bytecode")
        pkg.cross = cross
        vm = _load_pyc("models.vmamba_Fusion_efficross", "models/vmamba_Fusion_efficross.bytecode")
        pkg.vmamba_Fusion_efficross = vm
    ns = types.SimpleNamespace(cross=cross, vmamba=vm, VSSM_Fusion=vm.VSSM_Fusion,
                               ssc=sys.modules["selective_scan_cuda"])
    ns.ours_fwd, ns.ours_bwd = ns.ssc.fwd, ns.ssc.bwd
    ns.orig_core = (cross.cross_selective_scan, cross.cross_selective_scan_cross)
    _state["ref"] = ns
    return ns


def load_loss():
    load_reference()
    if "loss" not in sys.modules or not hasattr(sys.modules["loss"], "Fusionloss"):
        with _cuda_noop_if_no_gpu():
            _load_pyc("pytorch_msssim", "pytorch_msssim/__init__.bytecode")
            _load_pyc("loss", "loss.bytecode")
    return sys.modules["loss"]


@contextlib.contextmanager
def _cuda_noop_if_no_gpu(force: bool = False):
    """the reference cannot be constructed on a CPU-only host unmodified (``.cuda()`` on a constant
    in LDC.__init__).  At harness level ``Tensor.cuda`` becomes the identity while constructing on such a host."""
    if torch.cuda.is_available() and not force:
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def set_backend(name: str) -> None:
    """Who serves ``selective_scan_cuda.fwd/bwd`` for the reference's SelectiveScan Function (models/cross.py:119, 130-133)."""
    ref = load_reference()
    if name == "ours":
        ref.ssc.fwd, ref.ssc.bwd = ref.ours_fwd, ref.ours_bwd
    elif name == "ref_cuda":
        from oracle import build_ref
        ext = build_ref.load_ref()
        if ext is None:
            raise RuntimeError("oracle/_ref/selective_scan_cuda_ref.so is not built")
        ref.ssc.fwd, ref.ssc.bwd = ext.fwd, ext.bwd
    elif name == "cpu_oracle":
        iface = ref_scan_interface()

        def fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus):
            with torch.no_grad():
                out, last = iface.selective_scan_ref(u, delta, A, B, C, D, z=z, delta_bias=delta_bias,
                                                     delta_softplus=delta_softplus, return_last_state=True)
            x = torch.zeros(u.shape[0], u.shape[1], 1, 2 * A.shape[1], device=u.device)
            x[:, :, 0, 1::2] = last
            return [out, x]

        def bwd(*a, **k):
            raise RuntimeError("cpu_oracle backend: forward only (selective_scan_ref's autograd is O(L^2), SURVEY.md 3.4)")

        ref.ssc.fwd, ref.ssc.bwd = fwd, bwd
    else:
        raise ValueError(name)
    _state["backend"] = name


def set_fuse(mode: str | None) -> None:
    """None: model code exactly as shipped.  "patch": rebind the SS2D core functions (ss2d.patch_reference)."""
    ref = load_reference()
    ref.cross.cross_selective_scan, ref.cross.cross_selective_scan_cross = ref.orig_core
    if mode == "patch":
        from fusionmamba_b200 import ss2d
        ss2d.patch_reference(ref.cross)
    elif mode not in (None, "none", "swap"):
        raise ValueError(mode)


def fix_device_attrs(model: torch.nn.Module, device) -> torch.nn.Module:
    """``center_mask`` of LDC / Enhancement_texture_LDC is a plain tensor attribute (not a buffer), so ``model.to(device)``
    leaves it behind; move it at harness level."""
    for m in model.modules():
        cm = m.__dict__.get("center_mask")
        if isinstance(cm, torch.Tensor):
            m.center_mask = cm.to(device)
    return model


def build_model(kind: str = "tiny", device="cpu", seed: int = 0, **kw) -> torch.nn.Module:
    """VSSM_Fusion with the reference's constructor defaults; parameters are always initialised on the CPU generator
    (``torch.manual_seed(seed)``), so the same seed gives the same weights on every host."""
    ref = load_reference()
    cfg = dict(TINY if kind == "tiny" else FULL)
    cfg.update(kw)
    torch.manual_seed(seed)
    with _cuda_noop_if_no_gpu(force=True):       # construct on CPU everywhere: identical init stream with and without a GPU
        model = ref.VSSM_Fusion(**cfg)
    model = model.to(device)
    return fix_device_attrs(model, device)


def swap_ss2d(model: torch.nn.Module) -> int:
    """Replace every reference SS2D / SS2D_cross_new by this library's module of the same name (same state_dict)."""
    from fusionmamba_b200 import ss2d
    return ss2d.adopt_reference_modules(model)


def swap_layer_norms(model: torch.nn.Module) -> int:
    """nn.LayerNorm -> fusionmamba_b200.blocks.FastLayerNorm (same parameters), inference path only."""
    from fusionmamba_b200 import blocks
    return blocks.adopt_layer_norms(model)


def swap_vss_blocks(model: torch.nn.Module) -> int:
    """VSSBlock_new.forward -> fusionmamba_b200.blocks' fused inference tail (ECA, LDC weight cache, BiAttn, adds, norm2)."""
    from fusionmamba_b200 import blocks
    return blocks.adopt_vss_blocks(model)


def ss2d_modules(model: torch.nn.Module):
    """(name, module) of every SS2D-like block in forward-definition order."""
    out = []
    for name, m in model.named_modules():
        if type(m).__name__ in ("SS2D", "SS2D_cross_new"):
            out.append((name, m))
    return out


@contextlib.contextmanager
def capture_ss2d_outputs(model: torch.nn.Module, store: list):
    """Forward hooks recording every SS2D output in call order: (module name, tensor)."""
    hooks = []
    for name, m in ss2d_modules(model):
        hooks.append(m.register_forward_hook(lambda mod, inp, out, _n=name: store.append((_n, out.detach()))))
    try:
        yield store
    finally:
        for h in hooks:
            h.remove()


def weights_fingerprint(model: torch.nn.Module) -> dict:
    """Cheap, order-sensitive fingerprint of the initialisation (float64 sums), to prove two hosts built the same weights."""
    tot, tot_abs, n = 0.0, 0.0, 0
    for p in model.parameters():
        q = p.detach().double().cpu()
        tot += float(q.sum()); tot_abs += float(q.abs().sum()); n += q.numel()
    return {"n_params": n, "sum": tot, "sum_abs": tot_abs}


def make_pair(batch: int, H: int = 256, W: int = 256, seed: int = 0, device="cpu"):
    """Synthetic image pair, values in [0, 1] like TaskFusion_dataset.py:259-261."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x1 = torch.rand(batch, 1, H, W, generator=g)
    x2 = torch.rand(batch, 1, H, W, generator=g)
    return x1.to(device), x2.to(device)

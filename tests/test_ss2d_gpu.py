"""GPU parity tests of the SS2D core and modules against fixtures produced by the REFERENCE's own model code
(tests/golden/make_golden_ss2d.py: models/cross.py SS2D / SS2D_cross_new / cross_selective_scan with the scan served by
selective_scan_ref).  Everything below runs on the sm_100a kernels through the C ABI: fm_scan_unfold, fm_selective_scan_fwd,
fm_scan_merge and their backward counterparts.  Tolerance: rtol 1e-4 (fp32, north_star) with an absolute floor relative to the
largest reference magnitude (tests/util.py); parameter gradients, which sum thousands of terms, use 5x that."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
F32 = torch.float32


@pytest.fixture(autouse=True)
def _exact_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _cuda(a):
    return torch.from_numpy(np.asarray(a)).cuda()


@pytest.mark.parametrize("name", ["core_v2_odd", "core_v2_even"])
def test_cross_selective_scan_matches_reference(name):
    from fusionmamba_b200 import _lib, ss2d
    g = np.load(os.path.join(GOLD, f"ss2d_{name}.npz"))
    leaves = {k: _cuda(g[k]).requires_grad_() for k in ("x", "x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds")}
    norm = torch.nn.LayerNorm(leaves["x"].shape[1]).cuda()
    with torch.no_grad():
        norm.weight.copy_(_cuda(g["norm_weight"])); norm.bias.copy_(_cuda(g["norm_bias"]))
    n0 = _lib.launch_count()
    y = ss2d.cross_selective_scan(leaves["x"], leaves["x_proj_weight"], None, leaves["dt_projs_weight"], leaves["dt_projs_bias"],
                                  leaves["A_logs"], leaves["Ds"], norm, nrows=1, delta_softplus=True, step_size=2)
    y.backward(_cuda(g["g"]))
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 == 6, "unfold, scan fwd, merge + their three backward kernels"
    assert_close(y, g["y"], F32, name + " y")
    assert_close(leaves["x"].grad, g["dx"], F32, name + " dx", rtol_mul=5, atol_mul=5)
    for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds"):
        assert_close(leaves[k].grad, g["d" + k], F32, f"{name} d{k}", rtol_mul=5, atol_mul=5)
    assert_close(norm.weight.grad, g["dnorm_weight"], F32, name + " dnorm_weight", rtol_mul=5, atol_mul=5)
    assert_close(norm.bias.grad, g["dnorm_bias"], F32, name + " dnorm_bias", rtol_mul=5, atol_mul=5)


def _load_module(g, cls):
    kwargs = ast.literal_eval(str(g["kwargs"]))
    m = cls(**kwargs)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    m.load_state_dict(sd, strict=True)            # the reference's state_dict keys and shapes, verbatim
    return m.cuda()


@pytest.mark.parametrize("name,cls_name,core_only", [("mod_v2", "SS2D", False), ("mod_v2_odd", "SS2D", False),
                                                      ("mod_v0", "SS2D", True), ("mod_cross", "SS2D_cross_new", False)])
def test_module_matches_reference(name, cls_name, core_only):
    from fusionmamba_b200 import ss2d
    g = np.load(os.path.join(GOLD, f"ss2d_{name}.npz"))
    m = _load_module(g, getattr(ss2d, cls_name))
    xs = [_cuda(g[k]).requires_grad_() for k in sorted(f for f in g.files if f.startswith("x") and f[1:].isdigit())]
    out = m.forward_core(xs[0], channel_first=True) if core_only else m(*xs)
    out.backward(_cuda(g["g"]))
    torch.cuda.synchronize()
    assert_close(out, g["out"], F32, name + " out")
    for i, x in enumerate(xs):
        assert_close(x.grad, g[f"dx{i}"], F32, f"{name} dx{i}", rtol_mul=5, atol_mul=5)
    seen = 0
    for k, p in m.named_parameters():
        if "grad/" + k in g.files:
            assert p.grad is not None, k
            assert_close(p.grad, g["grad/" + k], F32, f"{name} grad {k}", rtol_mul=5, atol_mul=5)
            seen += 1
    assert seen >= (7 if core_only else 10)       # the v0 core alone does not touch in_proj / conv2d / out_proj


@pytest.mark.parametrize("name,cls_name", [("mod_v2", "SS2D"), ("mod_v2_odd", "SS2D"), ("mod_cross", "SS2D_cross_new")])
def test_inference_fast_path_matches_reference_fixture(name, cls_name):
    """The no-grad route (fused conv+SiLU+unfold prologue, merge fused into the scan's store, LayerNorm(+gate) kernel; for the
    cross module also the fused input x1*x2 + x1 + x2 on the unfolded tensors) against the reference module's own output in the
    fixture -- directly, not through the training path; fp32 tolerance, and the library's kernels must actually have run."""
    from fusionmamba_b200 import _lib, ss2d
    g = np.load(os.path.join(GOLD, f"ss2d_{name}.npz"))
    m = _load_module(g, getattr(ss2d, cls_name)).eval()
    xs = [_cuda(g[k]) for k in sorted(f for f in g.files if f.startswith("x") and f[1:].isdigit())]
    n0 = _lib.launch_count()
    with torch.no_grad():
        out = m(*xs)
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 >= (4 if cls_name == "SS2D_cross_new" else 3), "the fused inference kernels did not run"
    assert_close(out, g["out"], F32, name + " inference out")


@pytest.mark.parametrize("mode_name", ["v2", "v0"])
@pytest.mark.parametrize("shape", [(2, 3, 5, 7), (1, 4, 8, 8), (2, 2, 1, 1), (1, 3, 64, 64)])
def test_unfold_merge_bit_exact_and_inverse(mode_name, shape):
    """Permutations are bit-exact against the CPU oracle's index maps; V2 merge(unfold(x)) == x; both backward kernels are the
    adjoint permutations (checked with autograd on random cotangents)."""
    from fusionmamba_b200 import ss2d
    from oracle import scan_oracle as so
    mode = ss2d.MAP_V2 if mode_name == "v2" else ss2d.MAP_V0
    B, D, H, W = shape
    torch.manual_seed(H * 131 + W)
    x = torch.randn(B, D, H, W, device="cuda", requires_grad=True)
    xs = ss2d.scan_unfold(x, mode)
    ref_xs = so.efficient_scan(x.detach().cpu().numpy()) if mode_name == "v2" else so.cross_scan_v0(x.detach().cpu().numpy())
    assert np.array_equal(xs.detach().cpu().numpy(), ref_xs)
    ys = torch.randn_like(xs).requires_grad_()
    y = ss2d.scan_merge(ys, H, W, mode)
    ref_y = so.efficient_merge(ys.detach().cpu().numpy(), H, W) if mode_name == "v2" else so.cross_merge_v0(ys.detach().cpu().numpy(), H, W)
    assert np.array_equal(y.detach().cpu().numpy(), ref_y)
    if mode_name == "v2":
        assert torch.equal(ss2d.scan_merge(xs.detach(), H, W, mode).view(B, D, H, W), x.detach())
    # adjointness: <unfold(x), ys> == <x, unfold^T(ys)> and the same for merge
    gx, = torch.autograd.grad(xs, x, ys.detach())
    lhs = (xs.detach().double() * ys.detach().double()).sum()
    rhs = (x.detach().double() * gx.double()).sum()
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))
    gy = torch.randn_like(y)
    gys, = torch.autograd.grad(y, ys, gy)
    lhs = (y.detach().double() * gy.double()).sum()
    rhs = (ys.detach().double() * gys.double()).sum()
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))


@pytest.mark.parametrize("mode_name", ["v2", "v0"])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(1, 5, 16, 12), (1, 2, 40, 48), (1, 3, 12, 40), (2, 2, 6, 16), (1, 70, 8, 8), (1, 2, 33, 64),
                                   (1, 1, 64, 36), (1, 3, 100, 8)])
def test_unfold_merge_tiled_vector_paths(mode_name, dt, shape):
    """Every tile size (8 / 16 / 32), whole and partial tiles, widths / heights that are and are not multiples of the 16-byte
    vector (mixed vector + element paths inside one launch), more channels than one CTA holds, fp32 and 16-bit data --
    bit-exact against the oracle's index maps (V0 merge: the reference's add order with its intermediate roundings)."""
    from fusionmamba_b200 import ss2d
    from oracle import scan_oracle as so
    mode = ss2d.MAP_V2 if mode_name == "v2" else ss2d.MAP_V0
    B, D, H, W = shape
    torch.manual_seed(H * 7 + W)
    x = torch.randn(B, D, H, W, device="cuda").to(dt)
    xs = ss2d.scan_unfold(x, mode)
    xn = x.float().cpu().numpy()
    ref_xs = so.efficient_scan(xn) if mode_name == "v2" else so.cross_scan_v0(xn)
    assert np.array_equal(xs.float().cpu().numpy(), ref_xs)
    ys = torch.randn(*xs.shape, device="cuda").to(dt)
    y = ss2d.scan_merge(ys, H, W, mode)
    if mode_name == "v2":
        assert np.array_equal(y.float().cpu().numpy(), so.efficient_merge(ys.float().cpu().numpy(), H, W))
        assert torch.equal(ss2d.scan_merge(xs, H, W, mode).view(B, D, H, W), x)
    else:
        L = H * W                                  # ((o0 + flip o2) + wh o1) + wh flip o3, rounded to dt after every add like torch
        o = ys.view(B, 4, D, L)
        wh = lambda t: t.view(B, D, W, H).transpose(2, 3).reshape(B, D, L)
        ref = ((o[:, 0] + o[:, 2].flip(-1)) + wh(o[:, 1])) + wh(o[:, 3].flip(-1))
        assert torch.equal(y, ref)


def test_bf16_autocast_runs_scan_in_fp32():
    """Under autocast the projections run in bf16 but the scan sees fp32 tensors, like the reference (models/cross.py:312-318)."""
    from fusionmamba_b200 import ss2d
    torch.manual_seed(0)
    m = ss2d.SS2D(d_model=32, d_state=16).cuda().eval()
    x = torch.randn(2, 16, 16, 32, device="cuda")
    with torch.no_grad():
        ref = m(x)
        with torch.autocast("cuda", torch.bfloat16):
            out = m(x)
    assert out.dtype == torch.bfloat16
    err = (out.float() - ref).abs().max().item()
    assert err <= 5e-2 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (1, 12, 7, 5), (2, 16, 64, 64), (1, 4, 9, 16)])
def test_fused_merge_store_equals_unfused(shape):
    """Inference path: EfficientMerge fused into the scan's store (out_map = EFFICIENT_V2) is bit-identical to the unfused
    scan + fm_scan_merge (same arithmetic, only the store address changes) and launches no merge kernel."""
    from fusionmamba_b200 import _lib, scan_cuda, ss2d, selective_scan_fn
    B, D, H, W = shape
    N, L = 16, ss2d.scan_len(H, W, ss2d.MAP_V2)
    torch.manual_seed(H * 7 + W)
    u = torch.randn(B, 4 * D, L, device="cuda")
    delta = 0.5 * torch.rand(B, 4 * D, L, device="cuda")
    A = -0.5 * torch.rand(4 * D, N, device="cuda")
    x_dbl = torch.randn(B, 4, 3 + 2 * N, L, device="cuda")
    _, Bs, Cs = torch.split(x_dbl, [3, N, N], dim=2)                      # strided views, like the SS2D core passes them
    Dp, bias = torch.randn(4 * D, device="cuda"), 0.5 * torch.rand(4 * D, device="cuda")
    with torch.no_grad():
        ys = selective_scan_fn(u, delta, A, Bs, Cs, Dp, None, bias, True).view(B, 4, D, L)
        ref = ss2d.scan_merge(ys, H, W, ss2d.MAP_V2)
        n0 = _lib.launch_count()
        y = scan_cuda.fwd_merge_v2(u, delta, A, Bs, Cs, Dp, bias, True, H, W)
        torch.cuda.synchronize()
        assert _lib.launch_count() - n0 == 1
    assert torch.equal(y, ref)


def test_ss2d_inference_is_cuda_graph_capturable():
    """The library never synchronises or allocates behind torch's back, so a whole SS2D inference forward (unfold, scan with
    the fused merge, cuBLAS projections, LayerNorm) can be captured once and replayed -- the launch-bound short-L stages of
    the model (L = 16 ... 256) are the ones that need it."""
    from fusionmamba_b200 import _lib, ss2d
    torch.manual_seed(1)
    m = ss2d.SS2D(d_model=64, d_state=16).cuda().eval()
    x = torch.randn(4, 8, 8, 64, device="cuda")
    with torch.no_grad():
        ref = m(x).clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                m(x)                                   # warm-up on the side stream (cuBLAS workspaces, func attributes)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        static_x = x.clone()
        with torch.cuda.graph(g):
            static_y = m(static_x)
        x2 = torch.randn_like(x)
        static_x.copy_(x2)
        n0 = _lib.launch_count()
        g.replay()
        torch.cuda.synchronize()
        assert _lib.launch_count() == n0               # replay goes through the graph, not through the C ABI
        assert torch.allclose(static_y, m(x2), rtol=1e-5, atol=1e-6)
        static_x.copy_(x)
        g.replay()
        torch.cuda.synchronize()
        assert torch.allclose(static_y, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 4, 6, 192, 1024), (1, 4, 12, 40, 256), (3, 2, 1, 5, 17), (1, 4, 7, 33, 100)])
def test_dt_proj_matches_einsum(dt, shape):
    """fm_dt_proj == einsum("b k r l, k d r -> b k d l") on a strided view of x_dbl (models/cross.py:305-310): fp32 to 1e-6
    relative, 16-bit within one ulp of the rounded fp32-accumulated product; ragged L and channel counts take the element paths."""
    from fusionmamba_b200 import ss2d
    B, K, R, D, L = shape
    torch.manual_seed(R * 100 + L)
    x_dbl = torch.randn(B, K, R + 32, L, device="cuda").to(dt)
    dts = x_dbl[:, :, :R]                                   # strided view, last dim contiguous
    w = torch.randn(K, D, R, device="cuda")
    out = ss2d.dt_proj(dts, w.to(dt))
    ref = torch.einsum("bkrl,kdr->bkdl", dts.float(), w.to(dt).float())
    assert out.shape == (B, K, D, L) and out.dtype == dt and out.is_contiguous()
    if dt == torch.float32:
        assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    else:
        assert torch.allclose(out.float(), ref.to(dt).float(), rtol=1e-2 if dt == torch.bfloat16 else 2e-3, atol=1e-2)
    out32 = ss2d.dt_proj(dts, w)                            # fp32 weights with 16-bit activations
    assert out32.dtype == dt
    with pytest.raises(RuntimeError):
        ss2d.dt_proj(torch.randn(1, 4, 13, 64, device="cuda"), torch.randn(4, 8, 13, device="cuda"))   # rank > 12


def test_inference_weight_cache_follows_parameter_updates():
    """SS2D keeps autocast-dtype copies of its projection weights and -exp(A_logs) between no-grad forwards; an in-place
    parameter update (optimizer step, load_state_dict) must invalidate them, and the cached path must equal the uncached one."""
    from fusionmamba_b200 import ss2d
    torch.manual_seed(4)
    m = ss2d.SS2D(d_model=32, d_state=16).cuda().eval()
    x = torch.randn(2, 8, 8, 32, device="cuda")

    def run():
        with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
            return m(x).clone()

    y0 = run()
    assert torch.equal(run(), y0) and len(m._icache) >= 5            # second call served from the cache
    with torch.no_grad():
        m.in_proj.weight.mul_(1.5); m.A_logs.add_(0.3); m.x_proj_weight.mul_(0.5); m.out_proj.weight.mul_(2.0)
    y1 = run()
    m.clear_inference_cache()
    y2 = run()
    assert torch.equal(y1, y2) and not torch.equal(y1, y0)
    sd = m.state_dict()
    assert not any("icache" in k for k in sd)
    m.train()
    assert "_icache" not in m.__dict__


def test_graphed_forward_replays_ss2d_per_shape():
    """fusionmamba_b200.graph.GraphedForward: one capture per input signature, replays bypass the C ABI, results equal the
    eager forward (the forward has no atomics, so bit for bit), and a second shape gets its own graph."""
    from fusionmamba_b200 import _lib, ss2d
    from fusionmamba_b200.graph import GraphedForward
    torch.manual_seed(2)
    m = ss2d.SS2D(d_model=64, d_state=16).cuda().eval()
    fast = GraphedForward(m, autocast_dtype=torch.bfloat16)
    for shape in [(4, 8, 8, 64), (2, 16, 12, 64), (4, 8, 8, 64)]:
        x = torch.randn(*shape, device="cuda")
        with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
            ref = m(x).clone()
        y = fast(x)
        assert torch.equal(y, ref)
        x2 = torch.randn(*shape, device="cuda")
        with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
            ref2 = m(x2).clone()
        n0 = _lib.launch_count()
        y2 = fast(x2)
        torch.cuda.synchronize()
        assert _lib.launch_count() == n0
        assert torch.equal(y2, ref2)
    assert len(fast._graphs) == 2
    with pytest.raises(RuntimeError):
        fast(torch.randn(1, 8, 8, 64))           # CPU tensor: no fallback
    # an in-place weight update invalidates the captures (they hold cached low-precision copies of the weights)
    x = torch.randn(4, 8, 8, 64, device="cuda")
    with torch.no_grad():
        m.out_proj.weight.mul_(0.5); m.x_proj_weight.add_(0.01)
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        ref = m(x).clone()
    assert torch.equal(fast(x), ref) and len(fast._graphs) == 1


def test_graphed_forward_survives_cache_clear():
    """A captured graph bakes pointers to the module's derived inference tensors (-exp(A_logs), bf16 weight copies).  eval() /
    train() / clear_inference_cache() drop the module's references without changing any parameter version: the graph must keep
    that memory alive, so a replay after the cache was cleared -- and after the allocator had every chance to hand the freed
    blocks to someone else -- still equals the eager forward."""
    from fusionmamba_b200 import ss2d
    from fusionmamba_b200.graph import GraphedForward
    torch.manual_seed(5)
    m = ss2d.SS2D(d_model=64, d_state=16).cuda().eval()
    fast = GraphedForward(m, autocast_dtype=torch.bfloat16)
    x = torch.randn(4, 8, 8, 64, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        ref = m(x).clone()
    assert torch.equal(fast(x), ref)
    m.eval()                                   # clears the inference cache; no parameter version moves
    m.clear_inference_cache()
    torch.cuda.synchronize()
    junk = [torch.full((n,), float("nan"), device="cuda") for n in (64, 256, 1024, 4096, 16384, 65536) for _ in range(8)]
    torch.cuda.synchronize()
    assert torch.equal(fast(x), ref)           # replay of the SAME capture (no recapture: versions unchanged)
    assert len(fast._graphs) == 1
    del junk
    # LRU: more signatures than max_graphs evicts the oldest instead of failing
    small = GraphedForward(m, autocast_dtype=torch.bfloat16, max_graphs=2)
    for hw in (4, 6, 8):
        small(torch.randn(1, hw, hw, 64, device="cuda"))
    assert len(small._graphs) == 2


@pytest.mark.parametrize("itype", [torch.bfloat16, torch.float16])
def test_fp32_output_from_16bit_inputs_is_bit_identical_to_upcasting(itype):
    """out_dtype = fp32 with 16-bit u/delta/B/C (FmScanFwdParams.out_dtype) equals upcasting the same tensors to fp32 first --
    what the reference does (models/cross.py:312-318) -- bit for bit, including ragged sizes that take the scalar paths."""
    from fusionmamba_b200 import scan_cuda, ss2d
    for (B, D, H, W) in [(2, 16, 16, 16), (1, 8, 7, 9)]:
        N, L = 16, ss2d.scan_len(H, W, ss2d.MAP_V2)
        torch.manual_seed(3)
        u = torch.randn(B, 4 * D, L, device="cuda").to(itype)
        delta = (0.5 * torch.rand(B, 4 * D, L, device="cuda")).to(itype)
        A = -0.5 * torch.rand(4 * D, N, device="cuda")
        Bm, Cm = (torch.randn(B, 4, N, L, device="cuda").to(itype) for _ in range(2))
        Dp, bias = torch.randn(4 * D, device="cuda"), 0.5 * torch.rand(4 * D, device="cuda")
        with torch.no_grad():
            ref = scan_cuda.fwd_merge_v2(u.float(), delta.float(), A, Bm.float(), Cm.float(), Dp, bias, True, H, W)
            y = scan_cuda.fwd_merge_v2(u, delta, A, Bm, Cm, Dp, bias, True, H, W, out_dtype=torch.float32)
        assert y.dtype == torch.float32 and torch.equal(y, ref)
    with pytest.raises(RuntimeError, match="out_dtype"):
        scan_cuda.fwd_merge_v2(u.float(), delta.float(), A, Bm.float(), Cm.float(), Dp, bias, True, H, W, out_dtype=torch.bfloat16)


@pytest.mark.parametrize("shape", [(2, 192, 64 * 64), (3, 48, 35), (1, 1536, 64), (2, 33, 1), (1, 768, 250)])
@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_merge_norm_matches_torch_layer_norm(shape, odt):
    """fm_merge_norm == y.transpose(1,2).contiguous() -> nn.LayerNorm -> cast (models/cross.py:334-337); fp32 results within
    1e-5 of torch's (different summation order), bf16 results within one bf16 ulp of the rounded fp32 reference.  Includes a
    large-mean input (the shifted sums must not cancel)."""
    from fusionmamba_b200 import ss2d
    B, D, P = shape
    torch.manual_seed(D + P)
    y = torch.randn(B, D, P, device="cuda") * 3.0 + 50.0 * torch.randn(B, 1, P, device="cuda")
    norm = torch.nn.LayerNorm(D).cuda()
    with torch.no_grad():
        norm.weight.copy_(1.0 + 0.3 * torch.randn(D, device="cuda")); norm.bias.copy_(0.2 * torch.randn(D, device="cuda"))
        ref = norm(y.transpose(1, 2).contiguous())
        out = ss2d.merge_norm(y, norm, odt)
    assert out.shape == (B, P, D) and out.dtype == odt
    if odt == torch.float32:
        assert torch.allclose(out, ref, rtol=1e-5, atol=2e-5), (out - ref).abs().max().item()
    else:
        assert torch.allclose(out.float(), ref.to(odt).float(), rtol=1.6e-2, atol=1e-2)


def test_inference_core_equals_training_core():
    """The inference path (16-bit direct reads, fused merge store, fused transpose+LayerNorm) and the autograd path (separate
    kernels, torch LayerNorm) give the same SS2D output."""
    from fusionmamba_b200 import ss2d
    torch.manual_seed(5)
    m = ss2d.SS2D(d_model=48, d_state=16).cuda()
    x = torch.randn(3, 10, 14, 48, device="cuda")
    with torch.no_grad():
        fast = m(x)
    slow = m(x.clone().requires_grad_())
    assert torch.allclose(fast, slow.detach(), rtol=2e-5, atol=2e-5), (fast - slow).abs().max().item()


@pytest.mark.parametrize("shape", [(2, 64, 64, 48), (1, 7, 9, 20), (2, 33, 40, 16), (1, 1, 1, 8), (1, 5, 64, 35)])
@pytest.mark.parametrize("itype", [torch.float32, torch.bfloat16])
def test_conv_silu_unfold_matches_torch(shape, itype):
    """fm_conv_unfold == permute + depthwise conv2d(3x3, pad 1) + SiLU + EfficientScan (models/cross.py:727-731, 297), incl.
    odd sizes (zero padding of the unfold), image borders (zero padding of the conv), channel counts that do not fill a CTA
    tile and a channel offset into a wider channels-last tensor."""
    from fusionmamba_b200 import ss2d
    B, H, W, D = shape
    torch.manual_seed(H * 3 + W)
    xz = torch.randn(B, H, W, 2 * D + 3, device="cuda").to(itype)
    conv = torch.nn.Conv2d(D, D, 3, padding=1, groups=D).cuda()
    off = 2
    with torch.no_grad():
        x = xz[..., off:off + D].float().permute(0, 3, 1, 2).contiguous()
        ref = ss2d.scan_unfold(torch.nn.functional.silu(conv(x)), ss2d.MAP_V2)
        out = ss2d.conv_silu_unfold(xz, conv, D, off)
    assert out.shape == ref.shape and out.dtype == itype
    if itype == torch.float32:
        assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5), (out - ref).abs().max().item()
    else:
        assert torch.allclose(out.float(), ref, rtol=1e-2, atol=1e-2), (out.float() - ref).abs().max().item()


@pytest.mark.parametrize("shape", [(2, 64, 64, 48), (1, 7, 9, 20), (2, 33, 40, 16), (1, 1, 1, 8), (1, 5, 64, 35), (2, 16, 16, 192), (1, 17, 31, 32),
                                   (1, 128, 160, 192)])      # last: stage 0 of a 512x640 pair (BASELINE configs[3])
@pytest.mark.parametrize("itype", [torch.float32, torch.bfloat16])
def test_conv_silu_unfold_autograd_matches_torch(shape, itype):
    """ConvSiluUnfold (fm_conv_unfold forward + fm_conv_unfold_bwd: dx, dweight, dbias in one kernel) == torch autograd through
    permute + depthwise conv2d + SiLU + EfficientScan (models/cross.py:727-731, 297, 171-190) on a strided channel slice of a wider
    channels-last tensor: tile borders (16-pixel tiles), image borders, odd sizes, channel counts that do not fill a CTA tile."""
    from fusionmamba_b200 import ss2d
    B, H, W, D = shape
    torch.manual_seed(H * 5 + W)
    xz = torch.randn(B, H, W, 2 * D + 3, device="cuda").to(itype)
    conv = torch.nn.Conv2d(D, D, 3, padding=1, groups=D).cuda()
    off = 4 if D % 4 == 0 else 2
    g = torch.randn(B, 4, D, ss2d.scan_len(H, W, ss2d.MAP_V2), device="cuda").to(itype)
    # reference: fp32 torch ops on the same (rounded) inputs
    xr = xz[..., off:off + D].float().detach().requires_grad_()
    ref = ss2d.scan_unfold(torch.nn.functional.silu(conv(xr.permute(0, 3, 1, 2).contiguous())), ss2d.MAP_V2)
    ref.backward(g.float())
    rdx, rdw, rdb = xr.grad, conv.weight.grad.clone(), conv.bias.grad.clone()
    conv.weight.grad = conv.bias.grad = None
    xo = xz.detach().requires_grad_()
    out = ss2d.ConvSiluUnfold.apply(xo[..., off:off + D], conv.weight, conv.bias)
    out.backward(g)
    dx = xo.grad[..., off:off + D].float()
    assert xo.grad[..., :off].abs().max().item() == 0 and xo.grad[..., off + D:].abs().max().item() == 0
    tol = dict(rtol=1e-4, atol=1e-4) if itype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    assert torch.allclose(out.float(), ref, **tol)
    assert torch.allclose(dx, rdx, **tol), (dx - rdx).abs().max().item()
    # weight / bias gradients are sums over batch * H * W terms: bound relative to the tensor's scale
    for got, want, nm in ((conv.weight.grad, rdw, "dweight"), (conv.bias.grad, rdb, "dbias")):
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs().max().item()
        assert err <= (2e-5 if itype == torch.float32 else 2e-2) * scale + 1e-5, (nm, err, scale)


@pytest.mark.parametrize("cls_name", ["SS2D", "SS2D_cross_new"])
def test_training_prologue_is_the_fused_op(cls_name):
    """Under autograd the modules run conv + SiLU + unfold as ConvSiluUnfold (one launch of this library for the prologue and one
    for its backward, in place of the unfold launches plus torch's conv / SiLU / permute kernels), and the gradients equal those
    of the unfused path."""
    from fusionmamba_b200 import ss2d, _lib
    torch.manual_seed(3)
    m = getattr(ss2d, cls_name)(d_model=32, d_state=16, ssm_ratio=2.0, dt_rank="auto").cuda()
    xs = [torch.randn(2, 12, 10, 32, device="cuda", requires_grad=True) for _ in range(2 if cls_name == "SS2D_cross_new" else 1)]

    def run(fused):
        for p_ in m.parameters():
            p_.grad = None
        for x in xs:
            x.grad = None
        orig = ss2d.SS2D._fused_prologue_ok
        if not fused:
            ss2d.SS2D._fused_prologue_ok = lambda self, xz, training_too=False: False
        try:
            n0 = _lib.lib().fm_launch_count()
            y = m(*xs)
            y.square().mean().backward()
            n = _lib.lib().fm_launch_count() - n0
        finally:
            ss2d.SS2D._fused_prologue_ok = orig
        return y.detach(), [x.grad.clone() for x in xs], {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}, n

    y1, gx1, gp1, n1 = run(True)
    y0, gx0, gp0, n0 = run(False)
    assert torch.allclose(y1, y0, rtol=1e-4, atol=1e-5)
    for a_, b_ in zip(gx1, gx0):
        assert torch.allclose(a_, b_, rtol=1e-3, atol=1e-6), (a_ - b_).abs().max().item()
    assert gp1.keys() == gp0.keys()
    for k in gp0:
        scale = gp0[k].abs().max().item() + 1e-12
        assert (gp1[k] - gp0[k]).abs().max().item() <= 2e-3 * scale, k
    # the unfold launch and its backward became the fused launches (torch's conv / SiLU / permute kernels are gone); the two-input
    # module unfolds each modality (the unfused path unfolds x1*x2 + x1 + x2 once): one more forward and one more backward launch
    assert n1 == n0 + (2 if cls_name == "SS2D_cross_new" else 0)


@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_merge_norm_with_gate(odt):
    """dst = LayerNorm(y^T) * SiLU(z) with z read from the channels-last in_proj output (SS2D.forward, models/cross.py:728-740)."""
    from fusionmamba_b200 import ss2d
    B, D, P = 2, 96, 77
    torch.manual_seed(9)
    y = torch.randn(B, D, P, device="cuda") * 2.0
    xz = torch.randn(B, P, 2 * D, device="cuda").to(odt)
    norm = torch.nn.LayerNorm(D).cuda()
    with torch.no_grad():
        ref = norm(y.transpose(1, 2).contiguous()).to(odt) * torch.nn.functional.silu(xz[..., D:])
        out = ss2d.merge_norm(y, norm, odt, gate=(xz, D))
    tol = dict(rtol=1e-5, atol=2e-5) if odt == torch.float32 else dict(rtol=1.6e-2, atol=1e-2)
    assert out.dtype == odt and torch.allclose(out.float(), ref.float(), **tol), (out.float() - ref.float()).abs().max().item()


@pytest.mark.parametrize("D", [8, 32, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 100, 30])
@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_row_norm_vectorised_variants(D, odt):
    """Channels-last LayerNorm (+ SiLU gate) for every lanes-per-position / vectors-per-lane instance of row_norm_vec_kernel
    (dim % 4 == 0, up to 2048) and the scalar fallback (dim 30: not a multiple of 4; gate offset not a multiple of 4), with a
    position count that leaves a ragged last warp pass, against torch (models/cross.py:334-337, 728-740)."""
    from fusionmamba_b200 import ss2d
    B, P = 2, 37
    torch.manual_seed(D)
    ycl = torch.randn(B, P, D, device="cuda") * 2.0 + 20.0 * torch.randn(B, P, 1, device="cuda")
    xz = torch.randn(B, P, 2 * D, device="cuda").to(odt)
    norm = torch.nn.LayerNorm(D).cuda()
    with torch.no_grad():
        norm.weight.copy_(1.0 + 0.3 * torch.randn(D, device="cuda")); norm.bias.copy_(0.2 * torch.randn(D, device="cuda"))
        ref = norm(ycl)
        out = ss2d.merge_norm(ycl, norm, odt, channels_last=True)
        refg = ref.to(odt) * torch.nn.functional.silu(xz[..., D:])
        outg = ss2d.merge_norm(ycl, norm, odt, gate=(xz, D), channels_last=True)
    tol = dict(rtol=1e-5, atol=3e-5) if odt == torch.float32 else dict(rtol=1.6e-2, atol=1e-2)
    assert out.dtype == odt and out.shape == (B, P, D)
    assert torch.allclose(out.float(), ref.to(odt).float(), **tol), (out.float() - ref.float()).abs().max().item()
    assert torch.allclose(outg.float(), refg.float(), **tol), (outg.float() - refg.float()).abs().max().item()


@pytest.mark.parametrize("shape", [(2, 8, 8, 8), (1, 12, 7, 5), (2, 16, 64, 64), (1, 20, 9, 16)])
def test_channels_last_fused_store_and_row_norm(shape):
    """out_map = EFFICIENT_V2_CL writes the same values as the channel-first fused store, transposed; the channels-last
    LayerNorm equals the transposing one bit for bit up to summation order."""
    from fusionmamba_b200 import scan_cuda, ss2d
    B, D, H, W = shape
    N, L = 16, ss2d.scan_len(H, W, ss2d.MAP_V2)
    torch.manual_seed(H * 11 + W)
    u = torch.randn(B, 4 * D, L, device="cuda")
    delta = 0.5 * torch.rand(B, 4 * D, L, device="cuda")
    A = -0.5 * torch.rand(4 * D, N, device="cuda")
    Bm, Cm = torch.randn(B, 4, N, L, device="cuda"), torch.randn(B, 4, N, L, device="cuda")
    Dp, bias = torch.randn(4 * D, device="cuda"), 0.5 * torch.rand(4 * D, device="cuda")
    norm = torch.nn.LayerNorm(D).cuda()
    with torch.no_grad():
        norm.weight.add_(0.2 * torch.randn(D, device="cuda")); norm.bias.add_(0.2 * torch.randn(D, device="cuda"))
        y = scan_cuda.fwd_merge_v2(u, delta, A, Bm, Cm, Dp, bias, True, H, W)
        ycl = scan_cuda.fwd_merge_v2(u, delta, A, Bm, Cm, Dp, bias, True, H, W, channels_last=True)
        assert ycl.shape == (B, H * W, D) and torch.equal(ycl, y.transpose(1, 2))
        a = ss2d.merge_norm(y, norm, torch.float32)
        c = ss2d.merge_norm(ycl, norm, torch.float32, channels_last=True)
        assert torch.allclose(a, c, rtol=1e-5, atol=1e-5)
        assert torch.allclose(c, norm(ycl), rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("shape", [(32, 16, 16, 96), (4, 8, 8, 768), (2, 5, 7, 192), (3, 100, 384), (6, 40)])
def test_fast_layer_norm_matches_torch(shape):
    """blocks.FastLayerNorm == nn.LayerNorm on the inference path (fp32 in, fp32 out like torch under autocast), and defers to
    nn.LayerNorm's own forward when gradients are needed or the row kernel's preconditions do not hold."""
    from fusionmamba_b200 import _lib, blocks
    torch.manual_seed(sum(shape))
    D = shape[-1]
    ref = torch.nn.LayerNorm(D, eps=1e-6).cuda()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5); ref.bias.uniform_(-0.5, 0.5)
    host = torch.nn.Sequential(ref)
    assert blocks.adopt_layer_norms(host) == 1 and isinstance(host[0], blocks.FastLayerNorm)
    assert host[0].weight is ref.weight and host[0].bias is ref.bias
    x = torch.randn(*shape, device="cuda") * 3 + 1
    n0 = _lib.launch_count()
    with torch.no_grad():
        y = host(x)
        with torch.autocast("cuda", torch.bfloat16):
            ya = host(x)
    assert _lib.launch_count() - n0 == 2
    want = ref(x)
    assert y.dtype == torch.float32 and ya.dtype == torch.float32
    assert (y - want).abs().max() <= 2e-5 * max(1.0, float(want.abs().max()))
    assert torch.equal(y, ya)
    # gradient path: forward kernel + fm_layer_norm_bwd (two launches: row pass + column-sum finish), against torch autograd
    g = torch.randn_like(x)
    xg = x.clone().requires_grad_()
    host[0].weight.grad = host[0].bias.grad = None
    n0 = _lib.launch_count()
    host(xg).backward(g)
    assert _lib.launch_count() - n0 == 3
    got = (xg.grad.clone(), host[0].weight.grad.clone(), host[0].bias.grad.clone())
    ref2 = torch.nn.LayerNorm(D, eps=1e-6).cuda()
    ref2.load_state_dict(ref.state_dict())
    xr = x.clone().requires_grad_()
    ref2(xr).backward(g)
    for a, b, name in zip(got, (xr.grad, ref2.weight.grad, ref2.bias.grad), ("dx", "dweight", "dbias")):
        assert (a - b).abs().max() <= 2e-5 * max(1.0, float(b.abs().max())) * (1 if name == "dx" else 8), name


@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 16, 16, 96), (2, 8, 8, 768), (2, 5, 7, 192), (4, 100, 384)])
def test_norm_lowp_equals_autocast_layer_norm_then_cast(xdt, shape):
    """blocks.norm_lowp (norm-only form of fm_block_combine_norm) == what an autocast Linear sees behind nn.LayerNorm: the fp32
    LayerNorm of the (fp32 or 16-bit) residual stream rounded once to the autocast dtype."""
    from fusionmamba_b200 import blocks
    torch.manual_seed(7)
    x = (3 * torch.randn(*shape, device="cuda") + 0.5).to(xdt)
    ln = torch.nn.LayerNorm(shape[-1]).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.5, 0.5)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got = blocks.norm_lowp(ln, x)
            want = ln(x).to(torch.bfloat16)
    assert got is not None and got.dtype == torch.bfloat16 and got.shape == x.shape
    d = (got.float() - want.float()).abs()
    assert d.max().item() <= 2 ** -7 * want.float().abs().max().item()          # at most one bf16 ulp of the largest value
    assert (d > 0).float().mean().item() < 0.02                                  # and almost everywhere identical
    with torch.no_grad():
        assert blocks.norm_lowp(ln, x) is None                                   # outside autocast: not applicable


@pytest.mark.parametrize("itype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,Cc", [(3, 16, 16, 96), (2, 8, 8, 768), (2, 5, 7, 192), (1, 64, 64, 96)])
def test_block_tail_kernels_match_torch(itype, B, H, W, Cc):
    """fm_block_gates / fm_block_scale / fm_block_combine_norm against the ATen ops of VSSBlock_new._forward they replace
    (eca_layer, BiAttn, the adds, norm2: models/cross.py:744-768, 1236-1259, 1362-1377), written out with torch."""
    import torch.nn.functional as F
    from fusionmamba_b200 import blocks
    torch.manual_seed(B * 1000 + Cc)
    dev = "cuda"

    class Se(torch.nn.Module):                       # BiAttn's parameter layout
        def __init__(self):
            super().__init__()
            self.norm = torch.nn.LayerNorm(Cc)
            self.global_reduce = torch.nn.Linear(Cc, Cc // 8)
            self.channel_select = torch.nn.Linear(Cc // 8, Cc)
    se = Se().to(dev)
    with torch.no_grad():
        se.norm.weight.uniform_(0.5, 1.5); se.norm.bias.uniform_(-0.3, 0.3)
    eca_w = torch.randn(1, 1, 3, device=dev)
    norm2 = torch.nn.LayerNorm(Cc).to(dev)
    with torch.no_grad():
        norm2.weight.uniform_(0.5, 1.5); norm2.bias.uniform_(-0.3, 0.3)
    x = (torch.randn(B, H, W, Cc, device=dev) * 2 + 0.5).to(itype)
    xc = torch.randn(B, H, W, Cc, device=dev).to(itype)
    inp = torch.randn(B, H, W, Cc, device=dev)
    P = H * W
    with torch.no_grad():
        scale, g1 = blocks.block_gates(x.view(B, P, Cc), se, eca_w)
        _, g2 = blocks.block_gates(xc.view(B, P, Cc), se, None)
        # torch reference in fp32 (the gates are fp32 here; the reference computes them in the autocast dtype)
        xf = x.float()
        ye = torch.sigmoid(F.conv1d(xf.mean((1, 2)).unsqueeze(1), eca_w, padding=1).squeeze(1))
        def se_gate(v):
            m = se.norm(v.float()).mean((1, 2))
            return torch.sigmoid(se.channel_select(F.gelu(se.global_reduce(m))))
        assert (scale - ye).abs().max() < 2e-5 and (g1 - se_gate(x)).abs().max() < 2e-5 and (g2 - se_gate(xc)).abs().max() < 2e-5
        y = blocks.block_scale(x, scale)
        ref_y = (x + (x * scale.to(itype).view(B, 1, 1, Cc)))
        assert y.dtype == itype
        tol = 0 if itype == torch.float32 else 1
        assert (y.float() - ref_y.float()).abs().max() <= (1e-6 if itype == torch.float32 else 2e-2) * max(1.0, float(ref_y.abs().max()))
        x_new, y2 = blocks.block_combine_norm(inp, x, xc, g1, g2, norm2)
        mix = (x * g1.to(itype).view(B, 1, 1, Cc)) + (xc * g2.to(itype).view(B, 1, 1, Cc))
        ref_new = inp + mix
        ref_y2 = norm2(ref_new).to(itype)
        assert x_new.dtype == torch.float32 and y2.dtype == itype
        assert (x_new - ref_new).abs().max() <= (2e-6 if itype == torch.float32 else 1e-6) * max(1.0, float(ref_new.abs().max())) + \
            (0 if itype == torch.float32 else 0.0)
        assert (y2.float() - ref_y2.float()).abs().max() <= (2e-5 if itype == torch.float32 else 2e-2) * max(1.0, float(ref_y2.abs().max()))

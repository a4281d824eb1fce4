"""Model-level records of bench.py: the unmodified reference VSSM_Fusion (tools/model_harness.py) on this library.

  inference_record   BASELINE configs[2]: full model [2,2,9,2]/[2,9,2,2], bf16 autocast, no_grad, GLOBAL batch 32 of synthetic
                     256x256 pairs sharded over the ranks (no collective) -> fused pairs/s, per arm:
                       reference_cuda  reference model + the reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref)
                       dropin          reference model + our selective_scan_cuda (model code untouched)
                       patched         + ss2d.patch_reference (our SS2D core behind models.cross.cross_selective_scan)
                       swapped         + our SS2D modules adopted from the reference modules' state_dicts
                       swapped_graph   + the whole forward captured in one CUDA graph and replayed
                       swapped_ln_graph  + nn.LayerNorm served by the row kernel of fm_norm.cu (blocks.FastLayerNorm)
                       fused_blocks(_graph)  + the tail of every VSSBlock_new (ECA, BiAttn, adds, norm2) on fm_block.cu
                     plus e2e (host images in, fused images out, copies inside the timed region) on the fastest arm
  training_record    BASELINE configs[3]: full model, train(), fp32 like train.py, Fusionloss, Adam, batch 8 per GPU of synthetic
                     512x640 pairs (weak scaling), gradient all-reduce overlapped with backward (dist.GradReducer) when N > 1
  longseq_record     BASELINE configs[4]: one 1024x1024 pair, bf16 inference latency per arm

Timing: CUDA events on the current stream around K steps after W warm-up steps, max over ranks.
"""
from __future__ import annotations

import copy
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from tools import model_harness as mh  # noqa: E402


def _max_over_ranks(ms: float, dev, world: int) -> float:
    if world == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _time_steps(fn, steps, warmup, dev, world):
    for _ in range(warmup):
        fn()
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    _barrier(world)
    return _max_over_ranks(e0.elapsed_time(e1), dev, world) / steps


def _have_ref_cuda() -> bool:
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "selective_scan_cuda_ref.so"))


def _arm_setup(arm: str, model, model_swapped):
    """returns the callable module for this arm (selecting backend / fuse mode as a side effect)."""
    if arm == "reference_cuda":
        mh.set_backend("ref_cuda"); mh.set_fuse(None)
        return model
    mh.set_backend("ours")
    mh.set_fuse("patch" if arm == "patched" else None)
    return model_swapped if arm.startswith(("swapped", "fused")) else model


def inference_record(dev, rank, world, steps=10, warmup=3, global_batch=32, res=256, kind="full", arms=None, seed=0):
    from fusionmamba_b200 import _lib
    from fusionmamba_b200.dist import shard_batch
    from fusionmamba_b200.graph import GraphedForward
    a, b = shard_batch(global_batch, world, rank)
    nb = b - a
    model = mh.fix_device_attrs(mh.build_model(kind, device="cpu", seed=seed).eval().to(dev), dev)
    swapped = mh.fix_device_attrs(copy.deepcopy(model), dev)
    mh.swap_ss2d(swapped)
    swapped_ln = mh.fix_device_attrs(copy.deepcopy(swapped), dev)
    mh.swap_layer_norms(swapped_ln)
    fused = mh.fix_device_attrs(copy.deepcopy(swapped_ln), dev)
    mh.swap_vss_blocks(fused)
    x1h, x2h = mh.make_pair(global_batch, res, res, seed=seed + 1)
    x1h, x2h = x1h[a:b].contiguous().pin_memory(), x2h[a:b].contiguous().pin_memory()
    x1, x2 = x1h.to(dev), x2h.to(dev)
    arms = arms or ["reference_cuda", "dropin", "patched", "swapped", "swapped_graph", "swapped_ln_graph", "fused_blocks", "fused_blocks_graph"]
    if not _have_ref_cuda():
        arms = [x for x in arms if x != "reference_cuda"]
    pick = lambda a_: fused if a_.startswith("fused") else (swapped_ln if "_ln" in a_ else swapped)
    out = {"workload": f"BASELINE configs[2]: {kind} FusionMamba, bf16 autocast, no_grad, global batch {global_batch} of "
                       f"{res}x{res} pairs sharded over {world} GPU(s)", "global_batch": global_batch, "per_gpu_batch": nb,
           "n_gpus": world, "scaling": "strong", "dtype": "bf16 autocast (scan in fp32, models/cross.py:94)", "steps": steps,
           "warmup": warmup, "unit": "pairs/s", "arms": {}}
    graphs = {}
    for arm in arms:
        m = _arm_setup(arm, model, pick(arm))
        if arm.endswith("_graph"):
            gf = graphs.setdefault(arm, GraphedForward(m, autocast_dtype=torch.bfloat16))
            fn = lambda gf=gf: gf(x1, x2)
        else:
            def fn(m=m):
                with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
                    return m(x1, x2)
        n0 = _lib.launch_count()
        try:
            ms = _time_steps(fn, steps, warmup, dev, world)
        except Exception as e:  # an arm that cannot run is reported, not hidden
            out["arms"][arm] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            continue
        out["arms"][arm] = {"pairs_per_s": global_batch / (ms * 1e-3), "ms_per_step": ms,
                            "our_launches_per_step": (_lib.launch_count() - n0) / (steps + warmup)}
    ok = {k: v for k, v in out["arms"].items() if "pairs_per_s" in v and k != "reference_cuda"}
    if ok:
        best = max(ok, key=lambda k: ok[k]["pairs_per_s"])
        out["best_arm"] = best
        out["pairs_per_s"] = ok[best]["pairs_per_s"]
        if "reference_cuda" in out["arms"] and "pairs_per_s" in out["arms"]["reference_cuda"]:
            out["speedup_vs_reference_cuda"] = ok[best]["pairs_per_s"] / out["arms"]["reference_cuda"]["pairs_per_s"]
        # end to end on the best arm: pinned host images -> device, forward, fused image -> pinned host, every step
        m = _arm_setup(best, model, pick(best))
        yh = torch.empty(nb, 1, res, res).pin_memory()
        gfb = graphs.get(best)

        def e2e():
            a1, a2 = x1h.to(dev, non_blocking=True), x2h.to(dev, non_blocking=True)
            if gfb is not None:
                y = gfb(a1, a2)
            else:
                with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
                    y = m(a1, a2)
            yh.copy_(y, non_blocking=True)
        ms = _time_steps(e2e, steps, warmup, dev, world)
        out["e2e"] = {"pairs_per_s": global_batch / (ms * 1e-3), "ms_per_step": ms, "arm": best,
                      "h2d_bytes_per_step": 2 * x1h.numel() * 4, "d2h_bytes_per_step": yh.numel() * 4}
        if world > 1:
            # weak-scaling companion: the same arm with the FULL batch on every GPU (independent replicas, no collective)
            w1, w2 = mh.make_pair(global_batch, res, res, seed=seed + 11 + rank, device=dev)

            def weak():
                if gfb is not None:
                    return gfb(w1, w2)
                with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
                    return m(w1, w2)
            ms = _time_steps(weak, steps, warmup, dev, world)
            out["weak"] = {"per_gpu_batch": global_batch, "global_batch": global_batch * world, "arm": best,
                           "pairs_per_s": global_batch * world / (ms * 1e-3), "ms_per_step": ms, "scaling": "weak"}
    mh.set_backend("ours"); mh.set_fuse(None)
    del model, swapped, swapped_ln, fused, graphs
    torch.cuda.empty_cache()
    return out


def training_record(dev, rank, world, steps=3, warmup=2, per_gpu_batch=8, H=512, W=640, kind="full", arms=None, seed=0,
                    bucket_mb=32.0):
    from fusionmamba_b200.dist import GradReducer
    loss_mod = mh.load_loss()
    crit = loss_mod.Fusionloss()
    arms = arms or ["reference_cuda", "dropin", "patched", "patched_ln", "swapped_ln"]
    if not _have_ref_cuda():
        arms = [x for x in arms if x != "reference_cuda"]
    x1, x2 = mh.make_pair(per_gpu_batch, H, W, seed=seed + 100 + rank, device=dev)
    out = {"workload": f"BASELINE configs[3]: {kind} FusionMamba training step (fwd + bwd + grad all-reduce + Adam), fp32 like "
                       f"train.py, batch {per_gpu_batch}/GPU of {H}x{W} pairs, {world} GPU(s)", "per_gpu_batch": per_gpu_batch,
           "global_batch": per_gpu_batch * world, "n_gpus": world, "scaling": "weak", "steps": steps, "warmup": warmup,
           "unit": "pairs/s", "arms": {}}
    for arm in arms:
        model = mh.fix_device_attrs(mh.build_model(kind, device="cpu", seed=seed).to(dev), dev).train()
        if arm.startswith("swapped"):
            mh.swap_ss2d(model)                  # our SS2D / SS2D_cross_new modules: conv + SiLU + unfold as one op, one-kernel backward
            model = mh.fix_device_attrs(model, dev).train()
        if arm.endswith("_ln"):
            mh.swap_layer_norms(model)           # LayerNorm forward + backward on this library's kernels
        _arm_setup("patched" if arm.startswith("patched") else arm, model, model)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)                       # train.py:107
        red = GradReducer(model.parameters(), bucket_mb=bucket_mb) if world > 1 else None
        ev = []

        def step(record=False):
            if red is not None:
                red.zero_grad()
            else:
                opt.zero_grad(set_to_none=True)
            y = model(x1, x2)
            ones, zeros = torch.ones_like(y), torch.zeros_like(y)
            y = torch.where(y > ones, ones, y)
            y = torch.where(y < zeros, zeros, y)                                  # train.py:149-152
            loss, *_ = crit(image_vis=x1, image_ir=x2, generate_img=y, i=0, labels=None)
            loss.backward()
            if record:
                e_b = torch.cuda.Event(enable_timing=True); e_b.record()
            if red is not None:
                red.finish()
            if record:
                e_r = torch.cuda.Event(enable_timing=True); e_r.record()
                ev.append((e_b, e_r))
            opt.step()
            return loss

        try:
            for _ in range(warmup):
                step()
            _barrier(world)
            torch.cuda.reset_peak_memory_stats(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step(record=True)
            e1.record()
            _barrier(world)
            ms = _max_over_ranks(e0.elapsed_time(e1), dev, world) / steps
            rec = {"pairs_per_s": per_gpu_batch * world / (ms * 1e-3), "ms_per_step": ms, "loss": float(loss),
                   "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
            if red is not None:
                exposed = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
                rec["allreduce_exposed_ms"] = _max_over_ranks(exposed, dev, world)
                # the same buckets reduced with nothing to overlap: NCCL time and bus bandwidth of the exchange itself
                _barrier(world)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for bkt in red.buckets:
                    dist.all_reduce(bkt["flat"])
                a1.record()
                _barrier(world)
                alone = _max_over_ranks(a0.elapsed_time(a1), dev, world)
                rec["allreduce_alone_ms"] = alone
                rec["allreduce_bytes"] = red.nbytes
                rec["allreduce_bus_gbs"] = 2 * (world - 1) / world * red.nbytes / (alone * 1e-3) / 1e9
                rec["buckets"] = len(red.buckets)
                red.remove()
            out["arms"][arm] = rec
        except Exception as e:
            out["arms"][arm] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        del model, opt, red
        torch.cuda.empty_cache()
    ok = {k: v for k, v in out["arms"].items() if "pairs_per_s" in v and k != "reference_cuda"}
    if ok:
        best = max(ok, key=lambda k: ok[k]["pairs_per_s"])
        out["best_arm"], out["pairs_per_s"], out["steps_per_s"] = best, ok[best]["pairs_per_s"], 1e3 / ok[best]["ms_per_step"]
        if "pairs_per_s" in out["arms"].get("reference_cuda", {}):
            out["speedup_vs_reference_cuda"] = ok[best]["pairs_per_s"] / out["arms"]["reference_cuda"]["pairs_per_s"]
    mh.set_backend("ours"); mh.set_fuse(None)
    return out


def longseq_record(dev, steps=5, warmup=2, res=1024, kind="full", arms=None, seed=0):
    """configs[4]: one res x res pair, bf16 autocast inference (stage-0 scans: L = 16384 through EfficientScan)."""
    model = mh.fix_device_attrs(mh.build_model(kind, device="cpu", seed=seed).eval().to(dev), dev)
    swapped = mh.fix_device_attrs(copy.deepcopy(model), dev)
    mh.swap_ss2d(swapped)
    x1, x2 = mh.make_pair(1, res, res, seed=seed + 7, device=dev)
    arms = arms or ["reference_cuda", "dropin", "swapped", "swapped_ln_graph", "fused_blocks_graph"]
    if not _have_ref_cuda():
        arms = [x for x in arms if x != "reference_cuda"]
    swapped_ln = fused = None
    if any("_ln" in a_ or a_.startswith("fused") for a_ in arms):
        swapped_ln = mh.fix_device_attrs(copy.deepcopy(swapped), dev)
        mh.swap_layer_norms(swapped_ln)
        fused = mh.fix_device_attrs(copy.deepcopy(swapped_ln), dev)
        mh.swap_vss_blocks(fused)
    out = {"workload": f"BASELINE configs[4]: {kind} FusionMamba, one {res}x{res} pair, bf16 autocast inference", "unit": "ms/pair",
           "arms": {}}
    ys = {}
    for arm in arms:
        m = _arm_setup(arm, model, fused if arm.startswith("fused") else (swapped_ln if "_ln" in arm else swapped))
        if arm.endswith("_graph"):
            from fusionmamba_b200.graph import GraphedForward
            gf = GraphedForward(m, autocast_dtype=torch.bfloat16)
            fn = lambda gf=gf: gf(x1, x2)
        else:
            def fn(m=m):
                with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
                    return m(x1, x2)
        try:
            ms = _time_steps(fn, steps, warmup, dev, 1)
            ys[arm] = fn().float().clone()
            out["arms"][arm] = {"ms_per_pair": ms}
        except Exception as e:
            out["arms"][arm] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    if "reference_cuda" in ys:
        sc = float(ys["reference_cuda"].abs().max())
        for k, v in ys.items():
            if k != "reference_cuda":
                out["arms"][k]["max_abs_diff_vs_reference_cuda_rel"] = float((v - ys["reference_cuda"]).abs().max()) / sc
    mh.set_backend("ours"); mh.set_fuse(None)
    return out


def kernel_breakdown(dev, batch=32, res=256, kind="full", arm="dropin", seed=0, top=25, train=False, res_w=None):
    """One profiled forward (torch.profiler, CUDA activities): GPU time per kernel name, grouped -- what fraction is the scan."""
    from torch.profiler import ProfilerActivity, profile
    model = mh.fix_device_attrs(mh.build_model(kind, device="cpu", seed=seed).eval().to(dev), dev)
    swapped = None
    if arm.startswith(("swapped", "fused")):
        swapped = mh.fix_device_attrs(copy.deepcopy(model), dev)
        mh.swap_ss2d(swapped)
        if "_ln" in arm or arm.startswith("fused"):
            mh.swap_layer_norms(swapped)
        if arm.startswith("fused"):
            mh.swap_vss_blocks(swapped)
    m = _arm_setup(arm, model, swapped)
    x1, x2 = mh.make_pair(batch, res, res_w or res, seed=seed + 1, device=dev)

    if train:
        crit = mh.load_loss().Fusionloss()
        m.train()

        def fn():
            m.zero_grad(set_to_none=True)
            y = m(x1, x2)
            loss, *_ = crit(image_vis=x1, image_ir=x2, generate_img=y.clamp(0, 1), i=0, labels=None)
            loss.backward()
    else:
        def fn():
            with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
                return m(x1, x2)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count)
            for e in prof.key_averages()]
    rows = [r for r in rows if r[1] > 0]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)

    def group(name):
        if name.startswith("fm::") or "fm::" in name:
            return "fusionmamba_b200 kernels"
        if "selective_scan" in name:
            return "reference scan kernels"
        if "gemm" in name.lower() or "cutlass" in name.lower() or "nvjet" in name.lower() or "xmma" in name.lower():
            return "GEMM / conv (cuBLAS, cuDNN)"
        if "cudnn" in name.lower() or "conv" in name.lower():
            return "GEMM / conv (cuBLAS, cuDNN)"
        return "elementwise / norm / copy (ATen)"
    groups = {}
    for n, t, c in rows:
        g = groups.setdefault(group(n), [0.0, 0])
        g[0] += t; g[1] += c
    mh.set_backend("ours"); mh.set_fuse(None)
    return {"arm": arm, "batch": batch, "res": res, "gpu_time_us": tot, "launches": sum(r[2] for r in rows),
            "groups": {k: {"us": v[0], "share": v[0] / tot, "launches": v[1]} for k, v in sorted(groups.items(), key=lambda kv: -kv[1][0])},
            "top": [{"name": n[:110], "us": t, "count": c} for n, t, c in rows[:top]]}


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["infer", "train", "long", "breakdown", "train_breakdown"])
    ap.add_argument("--kind", default="full")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--arm", default="dropin")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    if args.what == "infer":
        r = inference_record(dev, rank, world, args.steps, args.warmup, global_batch=args.batch or 32, kind=args.kind)
    elif args.what == "train":
        r = training_record(dev, rank, world, args.steps, args.warmup, per_gpu_batch=args.batch or 8, kind=args.kind)
    elif args.what == "long":
        r = longseq_record(dev, args.steps, args.warmup, kind=args.kind)
    elif args.what == "train_breakdown":
        r = kernel_breakdown(dev, batch=args.batch or 8, res=512, res_w=640, kind=args.kind, arm=args.arm, train=True, top=40)
    else:
        r = kernel_breakdown(dev, batch=args.batch or 32, kind=args.kind, arm=args.arm, top=40)
    r["wall_s"] = time.time() - t0
    if rank == 0:
        print(json.dumps(r), flush=True)
    if world > 1:
        dist.destroy_process_group()

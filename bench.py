#!/usr/bin/env python
"""bench.py -- selective-scan fwd+bwd throughput on BASELINE.json configs[1] (B=8, K=4, D=192, L=4096, N=16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f32|bf16] [--impl ours|reference]

A "step" is one forward + one backward pass of the scan over one batch of synthetic input
(distributions of mamba_ssm/ops/test_selective_scan.py:406-441).  Per rank the batch is fixed (weak scaling:
independent image-pair batches shard across GPUs with no data-path collective).

Printed JSON line (rank 0):
  value       whole-job algorithmic GB/s, inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e         same metric through the public API (selective_scan_fn + autograd) from pinned HOST buffers,
              H2D of every input and D2H of out + every gradient inside the timed region
  roofline    dominant kernel (backward): algorithmic bytes / mean kernel time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the CPU oracle (oracle/scan_oracle.c, OpenMP) on a bounded sample of the same workload
  ref_cuda    the reference's own CUDA kernels rebuilt unmodified for sm_100a (oracle/_ref), timed in the same run
  model       BASELINE configs[2]: fused pairs/s of the unmodified reference VSSM_Fusion at 256x256 (bf16, global batch 32
              sharded over the ranks), per arm (reference CUDA kernels / our drop-in / fused routes)  -- tools/model_bench.py
  train       BASELINE configs[3]: training step (fwd + bwd + overlapped NCCL gradient all-reduce + Adam) at 512x640
  longseq     BASELINE configs[4]: one 1024x1024 pair, bf16 inference latency (N = 1 only)
--impl reference times the CPU path as the reference arm: every step is the WHOLE configs[1] batch, forward + backward, by
oracle/scan_oracle.c on all host cores (kind "port").  The reference's own CPU path, the pure-PyTorch selective_scan_ref
(a Python loop over L, staged byte-code in baseline/_ref), cannot run the backward at L = 4096 (its autograd is O(L^2),
BASELINE.md section 2), so it is timed beside the port -- forward at the full shape once, backward at L = 512 -- and
reported in the same line as ``python_ref``.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(batch=8, n_groups=4, d_inner=192, dim=768, seqlen=4096, dstate=16)
METRIC = "selective-scan fwd+bwd algorithmic HBM GB/s (BASELINE configs[1]: B=8,K=4,D=192,L=4096,N=16)"


def bench_config(dtype):
    """One config dict for both arms (the driver compares them key by key)."""
    fb, bb = algo_bytes(CFG["batch"], CFG["dim"], CFG["seqlen"], CFG["dstate"], CFG["n_groups"], 4 if dtype == "f32" else 2)
    return {"workload": "BASELINE configs[1] selective_scan fwd+bwd", **CFG, "io_dtype": dtype, "per_gpu_batch": CFG["batch"],
            "l2": "working set 0.6 GB per step > 126 MB L2 (no flush needed)",
            "algorithmic_bytes_fwd": fb, "algorithmic_bytes_bwd": bb}


def algo_bytes(batch, dim, L, N, G, es, has_z=False):
    """SURVEY.md section 8d / BASELINE.md section 5: fwd (3E+2G)s+4P, bwd (5E+4G)s+8P (+2E s / +3E s with z)."""
    E, Gg, P = batch * dim * L, batch * G * N * L, dim * (N + 2)
    fwd = (3 * E + 2 * Gg) * es + 4 * P + (2 * E * es if has_z else 0)
    bwd = (5 * E + 4 * Gg) * es + 8 * P + (3 * E * es if has_z else 0)
    return fwd, bwd


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bwd_kernel_name():
    """The backward kernel the library's dispatch selects for the bench workload (scan_cuda._hck_len + fm_scan_bwd.cuh)."""
    from fusionmamba_b200 import scan_cuda
    c = CFG
    hl = scan_cuda._hck_len(c["dstate"], c["seqlen"], c["batch"] * c["dim"])
    if hl == 8:
        return "scan_bwd_ls2_kernel" if os.environ.get("FM_SCAN_BWD_LS2", "1") != "0" else "scan_bwd_ls_kernel"
    return "scan_bwd_rp_kernel"


def ncu_traffic(kernel_key, dtype):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, from the committed ncu --set full capture
    of this same workload (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); None if no capture is on file."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t[f"configs1_{dtype}"][kernel_key]["dram_bytes"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def bind_rank_cores(local, world):
    """Give each rank its own slice of the host cores (and, where the GPU reports one, the cores of its NUMA node), so that the
    ranks' pinned-memory copies and Python threads do not migrate over each other.  Returns a description for the JSON line."""
    try:
        import torch
        cores = sorted(os.sched_getaffinity(0))
        node = None
        try:
            bus = torch.cuda.get_device_properties(local).pci_bus_id
            dom = torch.cuda.get_device_properties(local).pci_domain_id
            dev = torch.cuda.get_device_properties(local).pci_device_id
            path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
            node = int(open(path).read().strip())
            if node >= 0:
                cl = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
                ncores = set()
                for part in cl.split(","):
                    a, _, b = part.partition("-")
                    ncores.update(range(int(a), int(b or a) + 1))
                if len(ncores & set(cores)) >= world:
                    cores = sorted(ncores & set(cores))
        except Exception:
            node = None
        per = max(1, len(cores) // world)
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return {"cores": len(mine), "first": mine[0], "numa_node": node}
    except Exception as e:  # binding is best effort
        return {"error": str(e)[:100]}


def cpu_oracle_run(sample_batch, sample_dim, L, N, G, steps=1):
    """Time the CPU oracle (fwd + bwd) on a bounded sample; returns (seconds per step, threads)."""
    import numpy as np
    from oracle import c_oracle
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    c_oracle.set_threads(cores)                  # every host core this process may use, whatever OMP_NUM_THREADS says
    rng = np.random.default_rng(0)
    f = lambda *s: rng.standard_normal(s, dtype=np.float32)
    r = lambda *s: rng.random(s, dtype=np.float32)
    u, delta = f(sample_batch, sample_dim, L), 0.5 * r(sample_batch, sample_dim, L)
    A, Bm, Cm = -0.5 * r(sample_dim, N), f(sample_batch, G, N, L), f(sample_batch, G, N, L)
    D, bias, g = f(sample_dim), 0.5 * r(sample_dim), f(sample_batch, sample_dim, L)
    c_oracle.scan_fwd(u[:, :8, :256], delta[:, :8, :256], A[:8], Bm[:, :1, :, :256], Cm[:, :1, :, :256], D[:8], None, bias[:8], True)
    t0 = time.perf_counter()
    for _ in range(steps):
        c_oracle.scan_fwd(u, delta, A, Bm, Cm, D, None, bias, True)
        c_oracle.scan_bwd(u, delta, A, Bm, Cm, D, None, bias, g, True)
    return (time.perf_counter() - t0) / steps, c_oracle.num_threads()


def python_ref_run(L_bwd=512):
    """The reference's OWN CPU path (selective_scan_ref, mamba_ssm/ops/selective_scan_interface.py:92-158) on the host cores:
    forward at the full configs[1] shape once; forward + autograd backward at L = L_bwd, batch 1 (O(L^2))."""
    import torch
    from tools import model_harness as mh
    if not mh.available():
        return {"unavailable": "baseline/_ref not staged"}
    iface = mh.ref_scan_interface()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    Bn, dim, L, N, G = CFG["batch"], CFG["dim"], CFG["seqlen"], CFG["dstate"], CFG["n_groups"]

    def inputs(bn, l):
        return (torch.randn(bn, dim, l), 0.5 * torch.rand(bn, dim, l), -0.5 * torch.rand(dim, N), torch.randn(bn, G, N, l),
                torch.randn(bn, G, N, l), torch.randn(dim), None, 0.5 * torch.rand(dim))
    u, dl, A, Bm, Cm, D, z, bias = inputs(Bn, L)
    t0 = time.perf_counter()
    with torch.no_grad():
        iface.selective_scan_ref(u, dl, A, Bm, Cm, D, z=None, delta_bias=bias, delta_softplus=True)
    fwd_s = time.perf_counter() - t0
    fb, _ = algo_bytes(Bn, dim, L, N, G, 4)
    rec = {"impl": "selective_scan_ref (pure PyTorch, reference byte-code)", "cores": cores, "fwd_s_configs1": fwd_s,
           "fwd_gbs_configs1": fb / fwd_s / 1e9}
    del u, dl, Bm, Cm
    u, dl, A, Bm, Cm, D, z, bias = inputs(1, L_bwd)
    for t in (u, dl, A, Bm, Cm, D, bias):
        t.requires_grad_()
    t0 = time.perf_counter()
    out = iface.selective_scan_ref(u, dl, A, Bm, Cm, D, z=None, delta_bias=bias, delta_softplus=True)
    out.backward(torch.randn_like(out))
    rec[f"fwd_bwd_s_batch1_L{L_bwd}"] = time.perf_counter() - t0
    f1, b1 = algo_bytes(1, dim, L_bwd, N, G, 4)
    rec[f"fwd_bwd_gbs_batch1_L{L_bwd}"] = (f1 + b1) / rec[f"fwd_bwd_s_batch1_L{L_bwd}"] / 1e9
    return rec


def run_reference(args):
    """Reference arm: the CPU path on the host cores (rank 0 only).  One step = the whole configs[1] batch, forward + backward."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    es = 4 if args.dtype == "f32" else 2
    Bn, sd, G = CFG["batch"], CFG["dim"], CFG["n_groups"]
    t0 = time.time()
    W = max(0, args.warmup)
    if W:
        cpu_oracle_run(Bn, sd, CFG["seqlen"], CFG["dstate"], G, steps=W)
    sec, thr = cpu_oracle_run(Bn, sd, CFG["seqlen"], CFG["dstate"], G, steps=max(1, args.steps))
    fb, bb = algo_bytes(Bn, sd, CFG["seqlen"], CFG["dstate"], G, es)
    val = (fb + bb) / sec / 1e9
    sample = (f"whole batch {Bn} (dim {sd}, L {CFG['seqlen']}, N {CFG['dstate']}), fwd+bwd per step, oracle/scan_oracle.c "
              f"(C + OpenMP restatement of selective_scan_ref, fp64 accumulate), {sec:.2f} s/step")
    try:
        pyref = python_ref_run() if not args.no_python_ref else None
    except Exception as e:  # reported, never fatal for the arm
        pyref = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": bench_config(args.dtype),
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "python_ref": pyref, "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


def long_kernel_record(dev, iters=10):
    """Scan forward at B=1, 768 rows, L=65536 (one 1024x1024 pair at the classic CrossScan length, BASELINE configs[4]): ours
    (time-split forward) vs the reference's CUDA kernel, fp32 and bf16 I/O."""
    import torch
    from fusionmamba_b200 import scan_cuda
    from oracle import build_ref
    ext = build_ref.load_ref()
    out = {"shape": {"batch": 1, "dim": 768, "seqlen": 65536, "dstate": 16, "n_groups": 4}}
    for name, it in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        torch.manual_seed(0)
        u = torch.randn(1, 768, 65536, device=dev).to(it)
        dl = (0.5 * torch.rand(1, 768, 65536, device=dev)).to(it)
        A = -0.5 * torch.rand(768, 16, device=dev)
        Bm, Cm = torch.randn(1, 4, 16, 65536, device=dev).to(it), torch.randn(1, 4, 16, 65536, device=dev).to(it)
        D, bias = torch.randn(768, device=dev), 0.5 * torch.rand(768, device=dev)
        pf, _ = scan_cuda.prepare_fwd(u, dl, A, Bm, Cm, D, None, bias, True, with_hck=False)

        def timeit(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        rec = {"ours_fwd_ms": timeit(lambda: scan_cuda.launch_fwd(pf, dev)), "time_split_workspace_bytes": int(pf.workspace_bytes)}
        if ext is not None:
            rec["reference_cuda_fwd_ms"] = timeit(lambda: ext.fwd(u, dl, A, Bm, Cm, D, None, bias, True))
            rec["speedup"] = rec["reference_cuda_fwd_ms"] / rec["ours_fwd_ms"]
        out[name] = rec
    return out


def ref_cuda_record(d, fb, bb, iters=10):
    """The reference's own CUDA kernels (selective_scan/*.cu built unmodified for sm_100a, oracle/_ref) on the same inputs."""
    import torch
    from oracle import build_ref
    ext = build_ref.load_ref()
    if ext is None:
        return {"unavailable": "oracle/_ref/selective_scan_cuda_ref.so not built"}
    a = (d["u"].detach(), d["delta"], d["A"], d["B"], d["C"], d["D"], None, d["delta_bias"])

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    _, x2 = ext.fwd(*a, True)
    f_ms = timeit(lambda: ext.fwd(*a, True))
    b_ms = timeit(lambda: ext.bwd(*a, d["g"], x2, None, None, True, False))     # includes its own zero-fill of dA/dB/dC
    return {"what": "reference selective_scan_cuda rebuilt unmodified for sm_100a, same inputs, same run",
            "fwd_ms": f_ms, "bwd_ms": b_ms, "fwd_gbs": fb / f_ms / 1e6, "bwd_gbs": bb / b_ms / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--e2e-slices", type=int, default=0, help="batch slices of the pipelined end-to-end step (0: default)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dtype", choices=["f32", "bf16"], default="f32")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-python-ref", action="store_true", help="reference arm: skip the pure-PyTorch selective_scan_ref timing")
    ap.add_argument("--no-model", action="store_true", help="skip the model-level records (configs[2], [3], [4])")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step record (configs[3])")
    ap.add_argument("--model-kind", choices=["full", "tiny"], default="full")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fusionmamba_b200 import _lib, scan_cuda, selective_scan_fn

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    _lib.lib()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    binding = bind_rank_cores(local, world) if world > 1 else None
    W = max(3, args.warmup)
    K = args.steps
    itype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    es = 4 if args.dtype == "f32" else 2
    Bn, dim, L, N, G = CFG["batch"], CFG["dim"], CFG["seqlen"], CFG["dstate"], CFG["n_groups"]
    fb, bb = algo_bytes(Bn, dim, L, N, G, es)

    torch.manual_seed(rank)
    host = dict(
        u=torch.randn(Bn, dim, L).to(itype), delta=(0.5 * torch.rand(Bn, dim, L)).to(itype),
        A=-0.5 * torch.rand(dim, N), B=torch.randn(Bn, G, N, L).to(itype), C=torch.randn(Bn, G, N, L).to(itype),
        D=torch.randn(dim), delta_bias=0.5 * torch.rand(dim), g=torch.randn(Bn, dim, L).to(itype))
    host = {k: v.pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in host.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel path, inputs resident in HBM ----------------------------------------------
    u = d["u"].requires_grad_()
    pf, (out, x) = scan_cuda.prepare_fwd(u, d["delta"], d["A"], d["B"], d["C"], d["D"], None, d["delta_bias"], True)
    scan_cuda.launch_fwd(pf, dev)
    pb, r = scan_cuda.prepare_bwd(u.detach(), d["delta"], d["A"], d["B"], d["C"], d["D"], None, d["delta_bias"], d["g"], x,
                                  None, None, True, False)
    acc = [r["dA"], r["dB"], r["dC"], r["dD"], r["ddelta_bias"]]

    def step(ev=None):
        if ev:
            ev[0].record()
        scan_cuda.launch_fwd(pf, dev)
        if ev:
            ev[1].record()
        for t in acc:
            t.zero_()
        if ev:
            ev[2].record()
        scan_cuda.launch_bwd(pb, dev)
        if ev:
            ev[3].record()

    for _ in range(W):
        step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = _lib.launch_count()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(evs[i])
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = _lib.launch_count() - n0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    bwd_ms = sum(e[2].elapsed_time(e[3]) for e in evs) / K
    value = (fb + bb) * K * world / (ms_total * 1e-3) / 1e9
    peak, peak_src = peaks()

    # ---------------- end to end through the public API, host buffers ----------------------------------
    e2e = None
    if not args.no_e2e:
        names = ("u", "delta", "A", "B", "C", "D", "delta_bias")

        # The batch is streamed in NB slices: H2D of slice i+1, compute of slice i and D2H of slice i-1 overlap on three
        # streams (PCIe is full duplex); every call is the public API (selective_scan_fn + autograd) on one slice.
        NB = args.e2e_slices if (args.e2e_slices > 0 and Bn % args.e2e_slices == 0) else (4 if Bn % 4 == 0 else 1)
        bs = Bn // NB
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        big, small = ("u", "delta", "B", "C", "g"), ("A", "D", "delta_bias")
        dev_in = [{k: torch.empty_like(host[k][i * bs:(i + 1) * bs], device=dev) for k in big} for i in range(NB)]
        dev_small = {k: torch.empty_like(host[k], device=dev) for k in small}
        done = [None] * NB                                       # compute-finished events of the previous step, per slice
        out_names = ("out", "du", "ddelta", "dB", "dC")
        outs_host = {"out": torch.empty_like(host["u"]).pin_memory(), "du": torch.empty_like(host["u"]).pin_memory(),
                     "ddelta": torch.empty_like(host["delta"]).pin_memory(), "dB": torch.empty_like(host["B"]).pin_memory(),
                     "dC": torch.empty_like(host["C"]).pin_memory(), "dA": torch.empty_like(host["A"]).pin_memory(),
                     "dD": torch.empty_like(host["D"]).pin_memory(), "ddelta_bias": torch.empty_like(host["delta_bias"]).pin_memory()}

        def e2e_step():
            main = torch.cuda.current_stream(dev)
            with torch.cuda.stream(s_in):
                s_in.wait_stream(main) if done[0] is None else None
                for k in small:
                    dev_small[k].copy_(host[k], non_blocking=True)
            pA, pD, pb = (dev_small[k].detach().requires_grad_() for k in small)
            for i in range(NB):
                sl = slice(i * bs, (i + 1) * bs)
                with torch.cuda.stream(s_in):
                    if done[i] is not None:
                        s_in.wait_event(done[i])                  # previous step finished reading this slice's buffers
                    for k in big:
                        dev_in[i][k].copy_(host[k][sl], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                main.wait_event(ev)
                lv = {k: dev_in[i][k].detach().requires_grad_() for k in ("u", "delta", "B", "C")}
                o = selective_scan_fn(lv["u"], lv["delta"], pA, lv["B"], lv["C"], pD, None, pb, True)
                o.backward(dev_in[i]["g"])
                done[i] = torch.cuda.Event()
                done[i].record(main)
                res = {"out": o.detach(), "du": lv["u"].grad, "ddelta": lv["delta"].grad, "dB": lv["B"].grad, "dC": lv["C"].grad}
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done[i])
                    for k in out_names:
                        res[k].record_stream(s_out)
                        outs_host[k][sl].copy_(res[k], non_blocking=True)
            with torch.cuda.stream(s_out):
                s_out.wait_stream(main)
                for k, t in (("dA", pA.grad), ("dD", pD.grad), ("ddelta_bias", pb.grad)):
                    t.record_stream(s_out)
                    outs_host[k].copy_(t, non_blocking=True)
            main.wait_stream(s_out)                               # the step ends when its results are in host memory
            return outs_host

        for _ in range(W):
            e2e_step()
        # a freshly booted box backs pinned host pages lazily: keep warming up (bounded) until the step time settles
        best, settled = float("inf"), 0
        for _ in range(30):
            torch.cuda.synchronize(dev)
            t0w = time.perf_counter()
            e2e_step()
            torch.cuda.synchronize(dev)
            dtw = time.perf_counter() - t0w
            settled = settled + 1 if dtw < 1.1 * best else 0
            best = min(best, dtw)
            if settled >= 3:
                break
        Ke = max(3, min(K, 10))
        for _ in range(Ke):          # back-to-back (pipelined) steps keep two steps' outputs alive: let the caching allocator
            e2e_step()               # grow to that footprint before anything is timed
        torch.cuda.synchronize(dev)
        barrier()
        # the host side of this path (pinned-page copies over PCIe on a shared multi-tenant host) is noisy -- the same
        # binary measures 40 ... 116 GB/s on different boxes / minutes (profiles/r01_e2e_noise.txt): five trials of Ke
        # steps, the MEDIAN trial is reported and every trial is listed next to it
        trials = []
        for _ in range(5):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(Ke):
                res = e2e_step()
            a1.record()
            barrier()
            tt = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            trials.append(float(tt.item()))
        te = torch.tensor([sorted(trials)[len(trials) // 2]], device=dev, dtype=torch.float64)      # median trial
        h2d = sum(host[k].numel() * host[k].element_size() for k in (*names, "g"))
        d2h = sum(t.numel() * t.element_size() for t in res.values())
        e2e = {"value": (fb + bb) * Ke * world / (float(te.item()) * 1e-3) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
               "ms_per_step": float(te.item()) / Ke, "trials_ms_per_step": [round(t / Ke, 3) for t in trials],
               "how": f"selective_scan_fn + backward per batch slice ({NB} slices), H2D / compute / D2H overlapped on 3 streams; median of 5 trials of {Ke} steps (all listed)"}

    # ---------------- CPU baseline (rank 0, N=1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, thr = cpu_oracle_run(1, dim, L, N, G)
        f1, b1 = algo_bytes(1, dim, L, N, G, es)
        cpu = {"value": (f1 + b1) / sec / 1e9, "unit": "GB/s", "cores": thr, "kind": "port",
               "sample": f"batch 1 of {Bn} (dim {dim}, L {L}, N {N}), fwd+bwd, oracle/scan_oracle.c fp64 accumulate, {sec:.2f} s"}

    # ---------------- the reference's own CUDA kernels on the same inputs (rank 0) -----------------------
    refc = None
    if rank == 0:
        try:
            refc = ref_cuda_record(d, fb, bb)
            refc["ours_fwd_ms"], refc["ours_bwd_ms"] = fwd_ms, bwd_ms
            if "fwd_ms" in refc:
                refc["speedup_fwd"], refc["speedup_bwd"] = refc["fwd_ms"] / fwd_ms, refc["bwd_ms"] / bwd_ms
        except Exception as e:
            refc = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    # ---------------- kernel-level companion of configs[4]: one 1024^2 pair at the V0 length (768 rows x 65536 steps) ----
    long_kernel = None
    if rank == 0 and world == 1:
        try:
            long_kernel = long_kernel_record(dev)
        except Exception as e:
            long_kernel = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    # ---------------- model-level records: pairs/s (configs[2]), training step (configs[3]), 1024^2 (configs[4]) ----
    model_rec = train_rec = long_rec = None
    if not args.no_model:
        try:                                       # release the scan benchmark's buffers (GPU and pinned host) first
            del d, host, out, x, r, acc, pf, pb, u
            if not args.no_e2e:
                del dev_in, outs_host, res
        except NameError:
            pass
        torch.cuda.empty_cache()
        from tools import model_bench
        try:
            model_rec = model_bench.inference_record(dev, rank, world, steps=10, warmup=3, kind=args.model_kind)
        except Exception as e:
            model_rec = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        if not args.no_train:
            try:
                train_rec = model_bench.training_record(dev, rank, world, steps=3, warmup=2, kind=args.model_kind)
            except Exception as e:
                train_rec = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        if world == 1:
            try:
                long_rec = model_bench.longseq_record(dev, steps=3, warmup=2, kind=args.model_kind)
            except Exception as e:
                long_rec = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    if long_rec is not None and long_kernel is not None:
        long_rec["scan_kernel"] = long_kernel
    elif long_kernel is not None:
        long_rec = {"scan_kernel": long_kernel}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": bench_config(args.dtype),
            "roofline": {"bound": "hbm", "kernel": bwd_kernel_name(), "achieved": bb / (bwd_ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": bb / (bwd_ms * 1e-3) / 1e9 / peak, "traffic": ncu_traffic("bwd", args.dtype),
                         "peak_source": peak_src, "kernel_ms": bwd_ms, "algorithmic_bytes": bb,
                         "note": "dominant kernel of the step; bound by the SM's shared-memory / L1 data pipe (ncu: LSU wavefronts "
                                 "64 % of peak on average, 75 % on the SMs that host 6 of the 768 warps) and fp32 issue at dstate 16, "
                                 "not by HBM (DESIGN.md section 4); traffic includes the dense 8-step state checkpoints "
                                 "(+201 MB read) that the algorithmic byte count excludes"},
            "roofline_fwd": {"bound": "hbm", "kernel": "scan_fwd16_kernel", "achieved": fb / (fwd_ms * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": fb / (fwd_ms * 1e-3) / 1e9 / peak, "kernel_ms": fwd_ms,
                             "traffic": ncu_traffic("fwd", args.dtype), "algorithmic_bytes": fb},
            "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches), "clocks": clocks,
            "ref_cuda": refc, "model": model_rec, "train": train_rec, "longseq": long_rec, "rank_binding": binding,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

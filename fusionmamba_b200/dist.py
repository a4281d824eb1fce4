"""Data-parallel plumbing for the SS2D path: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on the
B200 box, gloo in the CPU tests).  The reference has no distributed code at all (no DataParallel, no
NCCL call site); the path shards by image pair -- every (batch, channel) row of the scan is independent
(selective_scan_fwd_kernel.cuh:97-98 launches grid (batch, dim)) -- so

  * inference / the scan micro-benchmark: ``shard_batch`` splits the batch, NO data-path collective;
  * training: the only exchange is ONE gradient all-reduce per step (``allreduce_gradients``), bucketed so that NCCL's
    launch latency is amortised and the buckets of late layers can overlap the backward of early ones; parameters that
    did not receive a gradient (the reference has one: Differential_enhance.lastconv, models/cross.py:849-864) are
    skipped consistently on every rank because the skip depends only on the model structure.

``GradReducer`` is the training-step form (BASELINE configs[3]): gradients live in pre-flattened bucket buffers (each
``p.grad`` is a view, so nothing is gathered or scattered around the collective), every bucket's all-reduce is launched from
a post-accumulate-grad hook the moment its last gradient arrives, on a side stream, so the collectives of late layers overlap
the backward of early ones.  ``allreduce_gradients`` is the simple after-backward form kept for callers that own their grads.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_batch(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of this rank's contiguous share of ``n_items`` independent items (image pairs / scan rows); the first
    ``n_items % world_size`` ranks take one extra item.  Shares differ by at most one and cover the range exactly."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _buckets(params: List[torch.Tensor], bucket_bytes: int) -> List[List[torch.Tensor]]:
    out, cur, size = [], [], 0
    for p in params:
        nb = p.numel() * p.element_size()
        if cur and (size + nb > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
            out.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += nb
    if cur:
        out.append(cur)
    return out


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, average: bool = True,
                        group=None, async_op: bool = False):
    """Sum (or average) ``.grad`` of every parameter that has one across the process group, in flattened buckets of about
    ``bucket_mb`` MiB (reverse registration order: the gradients produced first by backward go out first).
    With ``async_op`` the NCCL work handles are returned together with a ``finish()`` callable that waits and scatters the
    reduced buckets back -- call it before the optimizer step."""
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("allreduce_gradients needs an initialised torch.distributed process group")
    world = dist.get_world_size(group)
    with_grad = [p for p in reversed(list(params)) if p.grad is not None]
    pending = []
    for bucket in _buckets([p.grad for p in with_grad], int(bucket_mb * (1 << 20))):
        flat = torch.cat([g.reshape(-1) for g in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        pending.append((work, flat, bucket))

    def finish():
        for work, flat, bucket in pending:
            work.wait()
            if average:
                flat.div_(world)
            off = 0
            for g in bucket:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()

    if async_op:
        return [w for w, _, _ in pending], finish
    finish()
    return None


class GradReducer:
    """Overlapped data-parallel gradient reduction for one model replica per rank.

        red = GradReducer(model.parameters(), bucket_mb=32)      # once, after the model is on its device
        for batch in ...:
            red.zero_grad()                                      # instead of optimizer.zero_grad()
            loss(model(batch)).backward()                        # buckets go out from hooks while backward runs
            red.finish()                                         # waits; grads now hold the mean over ranks
            optimizer.step()

    * Buckets follow reverse registration order (the order backward produces gradients); each ``p.grad`` is a view into its
      bucket's flat buffer for the life of the reducer: no torch.cat / copy-back around the collective.
    * A parameter the forward never uses (the reference has one, Differential_enhance.lastconv, models/cross.py:849-864)
      never fires its hook; its slot stays zero and its bucket is launched by ``finish()``.  Which buckets are late depends
      only on the model structure, so every rank issues the same collectives in the same order.
    * The collectives run on a side stream that waits for the event recorded when the bucket became ready; ``finish()`` makes
      the current stream wait for them (no host synchronisation on CUDA)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, average: bool = True, group=None):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("GradReducer needs an initialised torch.distributed process group")
        self.group, self.average = group, average
        self.world = dist.get_world_size(group)
        ps = [p for p in reversed(list(params)) if p.requires_grad]
        self.buckets = []                                        # dicts: flat, params, pending, ready
        self._bucket_of = {}
        for chunk in _buckets(ps, int(bucket_mb * (1 << 20))):
            flat = torch.zeros(sum(p.numel() for p in chunk), dtype=chunk[0].dtype, device=chunk[0].device)
            off, views = 0, []
            for p in chunk:
                views.append(flat[off:off + p.numel()].view_as(p))
                p.grad = views[-1]
                off += p.numel()
                self._bucket_of[p] = len(self.buckets)
            self.buckets.append({"flat": flat, "params": chunk, "views": views, "pending": len(chunk), "launched": False,
                                 "work": None})
        self.is_cuda = bool(ps) and ps[0].is_cuda
        self.stream = torch.cuda.Stream(device=ps[0].device) if self.is_cuda else None
        self.nbytes = sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets)
        self._next = 0                                           # buckets are launched strictly in index order on every rank
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in ps]

    # -- per-step protocol ------------------------------------------------------------------------------------------------
    def zero_grad(self) -> None:
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"], b["launched"], b["work"] = len(b["params"]), False, None
        self._next = 0

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        b = self.buckets[self._bucket_of[p]]
        b["pending"] -= 1
        self._launch_ready()

    def _launch_ready(self, force: bool = False) -> None:
        while self._next < len(self.buckets):
            b = self.buckets[self._next]
            if b["pending"] > 0 and not force:
                return
            self._launch(b)
            self._next += 1

    def _launch(self, b) -> None:
        for p, v in zip(b["params"], b["views"]):                # autograd keeps accumulating into the view; if something replaced
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():      # p.grad (a hook, set_to_none), fold it back in
                v.copy_(p.grad)
        if self.is_cuda:
            self.stream.wait_stream(torch.cuda.current_stream(b["flat"].device))
            with torch.cuda.stream(self.stream):
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        b["launched"] = True

    def finish(self) -> None:
        self._launch_ready(force=True)
        for b in self.buckets:
            if self.is_cuda:
                with torch.cuda.stream(self.stream):
                    b["work"].wait()
                    if self.average:
                        b["flat"].div_(self.world)
            else:
                b["work"].wait()
                if self.average:
                    b["flat"].div_(self.world)
        if self.is_cuda:
            torch.cuda.current_stream(self.buckets[0]["flat"].device).wait_stream(self.stream)
        for b in self.buckets:                                   # p.grad is the bucket view again for the optimizer / next step
            for p, v in zip(b["params"], b["views"]):
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    p.grad = v

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

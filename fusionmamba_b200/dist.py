"""Data-parallel plumbing for the SS2D path: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on the
B200 box, gloo in the CPU tests).  The reference has no distributed code at all (no DataParallel, no
NCCL call site); the path shards by image pair -- every (batch, channel) row of the scan is independent
(selective_scan_fwd_kernel.cuh:97-98 launches grid (batch, dim)) -- so

  * inference / the scan micro-benchmark: ``shard_batch`` splits the batch, NO data-path collective;
  * training: the only exchange is ONE gradient all-reduce per step (``allreduce_gradients``), bucketed so that NCCL's
    launch latency is amortised and the buckets of late layers can overlap the backward of early ones; parameters that
    did not receive a gradient (the reference has one: Differential_enhance.lastconv, models/cross.py:849-864) are
    skipped consistently on every rank because the skip depends only on the model structure.
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

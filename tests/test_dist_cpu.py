"""World-size-2 gloo tests (CPU) of the data-parallel plumbing: batch sharding covers the items exactly once, and the
bucketed gradient all-reduce equals the single-process gradient of the concatenated batch (the N>1 path of bench.py /
a training step uses exactly these two functions; NCCL replaces gloo on the B200 box)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_batch_partitions():
    from fusionmamba_b200.dist import shard_batch
    for n in (0, 1, 7, 8, 32, 33):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_batch(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_batch(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fusionmamba_b200.dist import allreduce_gradients, shard_batch
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        unused = torch.nn.Parameter(torch.zeros(4))            # never touched by forward: grad stays None on every rank
        x = torch.randn(8, 6)
        a, b = shard_batch(8, world, rank)
        loss = model(x[a:b]).pow(2).sum() / 8                 # per-rank share of the global-mean loss
        loss.backward()
        params = list(model.parameters()) + [unused]
        allreduce_gradients(params, bucket_mb=1e-4, average=False)   # tiny buckets: exercises the multi-bucket path
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        ref.load_state_dict(model.state_dict())
        (ref(x).pow(2).sum() / 8).backward()
        ok = all(torch.allclose(p.grad, r.grad, atol=1e-6) for p, r in zip(model.parameters(), ref.parameters()))
        ok = ok and unused.grad is None
        # async variant + averaging
        for p in model.parameters():
            p.grad.fill_(float(rank + 1))
        works, finish = allreduce_gradients(model.parameters(), average=True, async_op=True)
        finish()
        ok = ok and all(torch.allclose(p.grad, torch.full_like(p.grad, (1 + world) / 2)) for p in model.parameters())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_allreduce_gradients_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}


def _worker_reducer(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fusionmamba_b200.dist import GradReducer, shard_batch
        torch.manual_seed(0)

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a, self.b = torch.nn.Linear(6, 5), torch.nn.Linear(5, 3)
                self.unused = torch.nn.Linear(4, 4)           # like Differential_enhance.lastconv: never used in forward

            def forward(self, x):
                return self.b(torch.tanh(self.a(x)))
        model, ref = Net(), Net()
        ref.load_state_dict(model.state_dict())
        red = GradReducer(model.parameters(), bucket_mb=1e-4, average=True)      # tiny buckets: several collectives, in order
        assert len(red.buckets) > 2
        x = torch.randn(8, 6)
        a, b = shard_batch(8, world, rank)
        ok = True
        for step in range(2):                                  # second step: the bucket views are reused
            red.zero_grad()
            model(x[a:b]).pow(2).mean().backward()             # per-rank mean; the mean over ranks == global mean (equal shares)
            red.finish()
            ref.zero_grad(set_to_none=True)
            ref(x).pow(2).mean().backward()
            for (n, p), r in zip(model.named_parameters(), ref.parameters()):
                if n.startswith("unused"):
                    ok = ok and p.grad is not None and float(p.grad.abs().sum()) == 0.0 and r.grad is None
                else:
                    ok = ok and torch.allclose(p.grad, r.grad, atol=1e-6)
                ok = ok and p.grad.data_ptr() == red.buckets[red._bucket_of[p]]["views"][
                    [q is p for q in red.buckets[red._bucket_of[p]]["params"]].index(True)].data_ptr()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_reducer, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}

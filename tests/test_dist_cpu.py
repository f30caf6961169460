"""N > 1 host logic on the CPU: ray sharding and the training-step gradient all-reduce, gloo backend, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgnerf_b200.dist import allreduce_grads, gather_frame, shard_rays


def test_shard_rays_partitions_the_frame():
    for n, world, tile in [(307200, 8, 256), (1000, 3, 64), (5, 2, 256), (3136, 4, 56)]:
        parts = [shard_rays(n, r, world, tile) for r in range(world)]
        allidx = torch.cat(parts).sort()[0]
        assert torch.equal(allidx, torch.arange(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= tile
    assert torch.equal(shard_rays(10, 0, 1), torch.arange(10))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(7, 5))                  # an MLP weight
    table = torch.nn.Parameter(torch.randn(1, 100, 32))        # a point table
    frozen = torch.nn.Parameter(torch.randn(3), requires_grad=False)
    nograd = torch.nn.Parameter(torch.randn(4))                # no gradient on rank 1
    w.grad = torch.full_like(w, float(rank + 1))
    table.grad = torch.zeros_like(table)
    table.grad[0, rank * 10:(rank + 1) * 10] = 1.0             # ranks touch different rows (ray shards see different points)
    if rank == 0:
        nograd.grad = torch.ones_like(nograd)
    nbytes = allreduce_grads([w, table, frozen, nograd, None], average=False)
    ok = torch.allclose(w.grad, torch.full_like(w, 3.0)) and float(table.grad[0, :20].sum()) == 20 * 32 and float(table.grad[0, 20:].abs().sum()) == 0
    ok = ok and torch.allclose(nograd.grad, torch.ones(4)) and frozen.grad is None and nbytes == 4 * (35 + 3200 + 4)
    # frame assembly from ray shards
    n = 1000
    idx = shard_rays(n, rank, world, 64)
    full = gather_frame(idx[:, None].float().repeat(1, 3), idx, n, tile=64)
    ok = ok and torch.equal(full[:, 0], torch.arange(n).float())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_allreduce_grads_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]

"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding and the flat-buffer
gradient all-reduce used by the data-parallel trainer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsg_b200.parallel import FlatParams, shard_rows


def test_shard_rows_partitions_exactly():
    for n in (0, 1, 7, 128, 1_000_003):
        for world in (1, 2, 4, 8):
            sl = [shard_rows(n, r, world) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(sl, sl[1:]))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                   # identical initial weights on every rank
        model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
        flat = FlatParams(model)
        opt = torch.optim.SGD([flat.flat], lr=0.1)
        g = torch.Generator().manual_seed(100)
        X, Y = torch.randn(64, 6, generator=g), torch.randn(64, 3, generator=g)
        sl = shard_rows(64, rank, world)
        for _ in range(3):
            flat.zero_grad()
            loss = torch.nn.functional.mse_loss(model(X[sl]), Y[sl])
            loss.backward()
            flat.allreduce_grads()
            with torch.no_grad():
                flat.flat -= 0.1 * flat.grad
        assert all(p.data_ptr() >= flat.flat.data_ptr() for p in model.parameters())
        gathered = [torch.empty_like(flat.flat) for _ in range(world)]
        dist.all_gather(gathered, flat.flat)
        if rank == 0:
            out.put([t.numpy().copy() for t in gathered])      # by value: the child may exit before the parent reads
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_allreduce_matches_single_process_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gathered = [torch.from_numpy(a) for a in gathered]
    assert torch.equal(gathered[0], gathered[1])               # replicas stay bit-identical
    # equal shards + averaged gradients == full-batch gradient descent in one process
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    flat = FlatParams(model)
    g = torch.Generator().manual_seed(100)
    X, Y = torch.randn(64, 6, generator=g), torch.randn(64, 3, generator=g)
    for _ in range(3):
        flat.zero_grad()
        torch.nn.functional.mse_loss(model(X), Y).backward()
        with torch.no_grad():
            flat.flat -= 0.1 * flat.grad
    assert torch.allclose(gathered[0], flat.flat, rtol=1e-5, atol=1e-6)

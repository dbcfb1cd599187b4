"""N>1 host logic on CPU: world_size-2 gloo process group, blob broadcast + batch sharding + gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ctrlhair_b200 import parallel


def test_shard_range_partitions():
    for n in (1, 7, 64, 65, 512):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
                assert e0 == s1
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    parallel.init_process_group("gloo")
    nbytes = 4096 + 13
    blob = (torch.arange(nbytes) % 251).to(torch.uint8) if rank == 0 else None
    got = parallel.broadcast_blob(blob, nbytes, src=0)
    ok = bool(torch.equal(got, (torch.arange(nbytes) % 251).to(torch.uint8)))
    n_total = 7
    s, e = parallel.shard_range(n_total, rank, world)
    local = torch.arange(s, e, dtype=torch.float32)[:, None] * torch.ones(1, 3)  # "images" tagged by global index
    full = parallel.gather_shards(local, n_total, dst=0)
    if rank == 0:
        ok = ok and bool(torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float32)))
    else:
        ok = ok and full is None
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t) == float(world)
    # gradient all-reduce of the colour/texture training step (mean over ranks, as DDP: solver.py:68-74)
    grads = torch.arange(6, dtype=torch.float32) * (rank + 1)
    parallel.allreduce_mean_(grads)
    ok = ok and bool(torch.allclose(grads, torch.arange(6, dtype=torch.float32) * 1.5))
    ret[rank] = ok
    dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))

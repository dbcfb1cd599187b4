"""Multi-GPU plumbing: one process per GPU, image batch sharded contiguously, ONE weight-blob broadcast.

The path has no per-step exchange (eval-mode BatchNorm has no cross-sample term): the only
collective is the start-up broadcast of the packed weight blob (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import os

import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous partition of n images over `world` ranks; the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def broadcast_blob(blob, nbytes, src=0, device="cpu"):
    """Rank `src` passes the packed uint8 blob, the others pass None; everyone returns the blob on `device`."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return blob.to(device)
    if dist.get_rank() == src:
        if blob.numel() != nbytes:
            raise ValueError("blob has %d bytes, expected %d" % (blob.numel(), nbytes))
        buf = blob.to(device).contiguous()
    else:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(buf, src=src)
    return buf


def gather_shards(local, n_total, dst=0):
    """Optional: collect per-rank output shards [n_local, ...] on rank dst (all ranks must call)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(e - s for s, e in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    if rank != dst:
        return None
    return torch.cat([o[:e - s] for o, (s, e) in zip(outs, sizes)])


def allreduce_mean_(flat):
    """In-place mean of one flat gradient buffer over all ranks (what DDP does per bucket, solver.py:68-74): one
    collective per optimizer step, NCCL over NVLink on GPUs, gloo in the CPU tests.  No-op for a single process."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    return flat

"""Multi-GPU plumbing: one process per GPU, torch.distributed for the
rendezvous, NCCL (inside the CUDA library) for the per-level all-reduces of
the cut histograms.  The point set is sharded by contiguous index ranges
(SURVEY.md §8e); every rank passes its own shard to Rcb/Rib and gets the ids
of its shard back."""
from __future__ import annotations

import ctypes as C

from . import _lib
from .api import BackendError, Context


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) index range of `rank`: sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def make_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    err = _lib.lib().coupe_b200_nccl_unique_id(buf)
    if err != 0:
        raise BackendError(err)
    return buf.raw


def broadcast_unique_id(make_id=make_unique_id, group=None) -> bytes:
    """Rank 0 creates the NCCL unique id, torch.distributed broadcasts its 128 bytes."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def init_comm(ctx: Context, group=None) -> Context:
    """Attach an NCCL communicator spanning the torch.distributed group to `ctx`."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return ctx
    ctx.init_comm(broadcast_unique_id(group=group), rank, world)
    return ctx

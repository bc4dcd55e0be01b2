"""One-process-per-GPU plumbing around libogb's multi-rank contexts (torch.distributed only moves the
NCCL unique id and the timing scalars; the data-path collectives -- allgather of the index slices, of the packed pre-reduction
adjacency, flags and final edges -- are NCCL calls inside libogb, see csrc/ogb_device.cu)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (1 process: 0, 1, 0)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n, rank, world):
    """Query-read shard [lo, hi) of `rank`: contiguous ID ranges of ceil(n/world) reads -- the same rule
    as ogb_context::shard() in csrc/ogb_device.cu (read IDs are lexicographic ranks, i.e. effectively
    random with respect to genome position, so equal ranges carry equal work)."""
    per = (n + world - 1) // world
    return min(n, per * rank), min(n, per * (rank + 1))


def broadcast_bytes(payload, src=0):
    """Broadcasts a bytes object made on `src` (the 128-byte ncclUniqueId) to every rank."""
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def max_over_ranks(values, device="cpu"):
    """Element-wise max over ranks of a list of floats (per-step device times)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().tolist()


def make_context(backend="nccl"):
    """Initialises torch.distributed (when WORLD_SIZE > 1) and returns (Context, rank, world, local_rank)."""
    from .api import Context, nccl_unique_id
    rank, world, local = env_rank()
    uid = None
    if backend == "nccl" and torch.cuda.is_available():
        torch.cuda.set_device(local)      # object collectives stage through the current device
    if world > 1:
        if not dist.is_initialized():
            kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
            dist.init_process_group(backend, **kw)
        uid = broadcast_bytes(nccl_unique_id() if rank == 0 else None)
    return Context(local, rank, world, uid), rank, world, local

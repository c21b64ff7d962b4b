"""Batch sharding of the inference path over the GPUs of one box (one process per GPU).

Utterances are independent in the forward (every norm/attention/recurrence is per sample; eval
BatchNorm uses running statistics), so the N-GPU path is pure data parallelism with NO data-path
collective.  torch.distributed is used only for the rendezvous, the barrier
around the timed region and the max-over-ranks reduction of the device time.
"""
import os

import torch
import torch.distributed as dist


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment (1 process if absent)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)  # binds the communicator to this rank's GPU (no guessing)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_bounds(n_items, rank, world):
    """Contiguous [lo, hi) slice of n_items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (device time of the slowest rank)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_outputs(local_out, n_items, rank, world):
    """all_gather of per-rank output shards back into (n_items, ...) order (test/validation helper;
    the serving path leaves each shard on its own GPU)."""
    if world == 1:
        return local_out
    sizes = [hi - lo for lo, hi in (shard_bounds(n_items, r, world) for r in range(world))]
    n_max = max(sizes)  # all_gather needs equal shapes: pad the short shards
    padded = local_out.new_zeros((n_max,) + tuple(local_out.shape[1:]))
    padded[: local_out.shape[0]] = local_out
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)

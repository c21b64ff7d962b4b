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


class StreamingSeparator:
    """Serving loop over HOST batches (reference: the per-batch `.to(device)` -> `model(...)` -> `.cpu()` of test.py /
    src/system/core.py:86-92): pinned host inputs go to the device on a copy stream, the separated waveforms come back on a second
    one, both double-buffered, so the copies of batch n + 1 and n - 1 run under the kernels of batch n.  One model forward is in
    flight at a time (the model owns ONE workspace); only the copies overlap it.

        sep = StreamingSeparator(model, batch, samples, frames)
        for wav_h, lip_h, out_h in batches:      # pinned (B, L) / (B, 512, Tv) in, pinned (B, 1, L) out
            sep.submit(wav_h, lip_h, out_h)      # returns at once; out_h is valid after sep.drain() (or a later submit of its slot)
        sep.drain()
    """

    def __init__(self, model, batch, samples, frames, device=None, lip_channels=512):
        if not torch.cuda.is_available():
            raise RuntimeError("StreamingSeparator needs a CUDA device (the RTFS-Net B200 path has no CPU fallback)")
        self.model = model
        self.dev = torch.device(device if device is not None else torch.cuda.current_device())
        self.h2d = torch.cuda.Stream(self.dev)
        self.d2h = torch.cuda.Stream(self.dev)
        self.wav = [torch.empty(batch, samples, device=self.dev) for _ in range(2)]
        self.lip = [torch.empty(batch, lip_channels, frames, device=self.dev) for _ in range(2)]
        self.out = [torch.empty(batch, 1, samples, device=self.dev) for _ in range(2)]
        self.in_free = [None, None]   # event: the forward that read input slot s has finished
        self.out_free = [None, None]  # event: the device->host copy out of output slot s has finished
        self.n = 0

    def submit(self, wav_h, lip_h, out_h):
        s = self.n & 1
        self.n += 1
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.h2d):
            if self.in_free[s] is not None:
                self.h2d.wait_event(self.in_free[s])
            self.wav[s].copy_(wav_h, non_blocking=True)
            self.lip[s].copy_(lip_h, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.h2d)
        cur.wait_event(ready)
        if self.out_free[s] is not None:
            cur.wait_event(self.out_free[s])
        with torch.no_grad():
            self.out[s].copy_(self.model(self.wav[s], self.lip[s]))
        done = torch.cuda.Event()
        done.record(cur)
        self.in_free[s] = done
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(done)
            out_h.copy_(self.out[s], non_blocking=True)
            freed = torch.cuda.Event()
            freed.record(self.d2h)
        self.out_free[s] = freed

    def drain(self):
        """Blocks until every submitted batch has been copied back; also makes the caller's stream wait for the copies."""
        cur = torch.cuda.current_stream(self.dev)
        for e in self.out_free:
            if e is not None:
                cur.wait_event(e)
        self.d2h.synchronize()

"""world_size-2 gloo test (CPU) of the N>1 host path: shard bounds, barrier, max/sum reductions and
the output gather, with the oracle's STFT standing in for the per-utterance work."""
import os
import socket

import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, ret):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle import rtfs_oracle as O
    from rtfs_net_b200 import shard

    r, lr, w = shard.init("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(7)
    wav = torch.randn(n_items, 2048, generator=g)  # same on every rank
    lo, hi = shard.shard_bounds(n_items, rank, world)
    shard.barrier()
    local = O.stft_spec(wav[lo:hi])  # per-utterance work, no cross-rank dependency
    full = shard.gather_outputs(local, n_items, rank, world)
    t_max = shard.max_over_ranks(float(rank + 1))
    n_sum = shard.sum_over_ranks(float(hi - lo))
    shard.barrier()
    if rank == 0:
        ref = O.stft_spec(wav)
        ret["err"] = float((full - ref).abs().max())
        ret["t_max"] = t_max
        ret["n_sum"] = n_sum
    torch.distributed.destroy_process_group()


def test_shard_bounds_cover_everything():
    from rtfs_net_b200.shard import shard_bounds

    for n in (1, 7, 32, 33, 64):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharded_batch():
    world, n_items = 2, 5
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_items, ret), nprocs=world, join=True)
    assert ret["err"] == 0.0
    assert ret["t_max"] == 2.0
    assert ret["n_sum"] == n_items

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


def _train_worker(rank, world, port, ret):
    """Data-parallel training plumbing on CPU: flat parameter / gradient buffers, ONE all-reduce of the flat gradient, the
    SyncBatchNorm statistics pack (sums + count) of the CAF cell."""
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from rtfs_net_b200 import shard
    from rtfs_net_b200.train import FlatParams

    shard.init("gloo")
    torch.manual_seed(0)  # same initial weights on every rank, as DDP broadcasts them
    net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.PReLU(), torch.nn.Linear(3, 1))
    flat = FlatParams(net)
    assert flat.world == world
    assert all(o % 64 == 0 for o in flat.offsets)
    assert all(p.data_ptr() == flat.flat_p.data_ptr() + 4 * o for p, o in zip(flat.params, flat.offsets))
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()  # accumulates into the views of flat_g
    local = flat.flat_g.clone()
    flat.exchange()
    bufs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(bufs, local)
    ok = torch.allclose(flat.flat_g, sum(bufs))
    # statistics pack of the CAF BatchNorm under sync_batchnorm (train.py:145): channel sums + element count
    sums = torch.arange(6, dtype=torch.float64).view(3, 2) * (rank + 1)
    pack = torch.cat([sums.reshape(-1), torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)])
    dist.all_reduce(pack)
    if rank == 0:
        ret["grad_ok"] = bool(ok)
        ret["pad_zero"] = float(flat.flat_g[flat.params[0].numel():64].abs().max()) == 0.0
        ret["pack"] = pack.tolist()
    dist.destroy_process_group()


def test_two_rank_gloo_training_exchange():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_train_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["grad_ok"] and ret["pad_zero"]
    assert ret["pack"] == [0.0, 3.0, 6.0, 9.0, 12.0, 15.0, 30.0]


def _syncbn_worker(rank, world, port, ret):
    """train.FastSyncBatchNorm on two gloo ranks against one BatchNorm1d over the concatenated batch."""
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from rtfs_net_b200 import shard
    from rtfs_net_b200.train import FastSyncBatchNorm, use_fast_sync_batchnorm

    shard.init("gloo")
    g = torch.Generator().manual_seed(11)
    x_all = torch.randn(5, 6, 7, generator=g) * 2.0 + 0.5  # uneven shards: 3 + 2 utterances
    dy_all = torch.randn(5, 6, 7, generator=g)
    w0, b0 = torch.randn(6, generator=g), torch.randn(6, generator=g)
    lo, hi = shard.shard_bounds(5, rank, world)
    net = torch.nn.Sequential(torch.nn.SyncBatchNorm(6))
    assert use_fast_sync_batchnorm(net) == 1 and type(net[0]) is FastSyncBatchNorm
    bn = net[0]
    with torch.no_grad():
        bn.weight.copy_(w0)
        bn.bias.copy_(b0)
    bn.train()
    x = x_all[lo:hi].clone().requires_grad_(True)
    y = bn(x)
    y.backward(dy_all[lo:hi])
    gw, gb = bn.weight.grad.clone(), bn.bias.grad.clone()
    dist.all_reduce(gw)
    dist.all_reduce(gb)
    ref = torch.nn.BatchNorm1d(6)
    with torch.no_grad():
        ref.weight.copy_(w0)
        ref.bias.copy_(b0)
    ref.train()
    xr = x_all.clone().requires_grad_(True)
    yr = ref(xr)
    yr.backward(dy_all)
    errs = [float((y - yr[lo:hi]).abs().max()), float((x.grad - xr.grad[lo:hi]).abs().max()), float((gw - ref.weight.grad).abs().max()),
            float((gb - ref.bias.grad).abs().max()), float((bn.running_mean - ref.running_mean).abs().max()),
            float((bn.running_var - ref.running_var).abs().max()), float(abs(int(bn.num_batches_tracked) - int(ref.num_batches_tracked)))]
    ret[rank] = max(errs)
    dist.destroy_process_group()


def test_two_rank_gloo_fast_sync_batchnorm():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_syncbn_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert max(ret.values()) < 2e-5, dict(ret)

"""torchrun --nproc-per-node 2 tools/prof_train_dp.py : where the data-parallel training step spends its time (rank 0, torch.profiler)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import audionet_conf
from rtfs_net_b200 import AVNet, shard
from rtfs_net_b200.train import Trainer

rank, local, world = shard.init()
torch.cuda.set_device(local)
R, B, L = 6, 16, 32000
g = np.load(os.path.join(ROOT, "tests", "golden", "state_dict_rtfs.npz"))
sd = {k: torch.from_numpy(g[k]) for k in g.files}
m = AVNet(print_macs=False, **audionet_conf(R))
m.load_state_dict(sd, strict=True)
m = m.cuda()
if world > 1:
    m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)
tr = Trainer(m)
gen = torch.Generator().manual_seed(1 + rank)
tgt = (0.1 * torch.randn(B, 1, L, generator=gen)).cuda()
wav = tgt[:, 0] + 0.1 * torch.randn(B, L, generator=gen).cuda()
lip = torch.rand(B, 512, 50, generator=gen).cuda()
for _ in range(3):
    tr.step(wav, tgt, lip)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3):
    tr.step(wav, tgt, lip)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("world", world, "ms per step", e0.elapsed_time(e1) / 3)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step(wav, tgt, lip)
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))

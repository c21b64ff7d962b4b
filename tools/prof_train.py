"""One RTFS-Net-6 training step (B=16, 2 s) for profiling under ncu: 1 warm-up step outside the profiled range is not
possible with ncu's launch counting, so the list contains 2 steps; launch_summary.py aggregates per kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import audionet_conf
from rtfs_net_b200 import AVNet
from rtfs_net_b200.train import Trainer

R, B, L = 6, int(os.environ.get("PROF_B", 16)), 32000
g = np.load(os.path.join(ROOT, "tests", "golden", "state_dict_rtfs.npz"))
sd = {k: torch.from_numpy(g[k]) for k in g.files}
m = AVNet(print_macs=False, **audionet_conf(R))
m.load_state_dict(sd, strict=True)
m = m.cuda()
tr = Trainer(m)
gen = torch.Generator().manual_seed(1)
tgt = (0.1 * torch.randn(B, 1, L, generator=gen)).cuda()
wav = tgt[:, 0] + 0.1 * torch.randn(B, L, generator=gen).cuda()
lip = torch.rand(B, 512, 50, generator=gen).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    print(float(tr.step(wav, tgt, lip)))
torch.cuda.synchronize()

import os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import audionet_conf
from rtfs_net_b200 import AVNet
from rtfs_net_b200.train import live_tensors
from rtfs_net_b200.weights import prepare
g = np.load(os.path.join(ROOT, "tests", "golden", "state_dict_rtfs.npz"))
sd = {k: torch.from_numpy(g[k]) for k in g.files}
m = AVNet(print_macs=False, **audionet_conf(6)); m.load_state_dict(sd, strict=True); m = m.cuda().train()
for _ in range(3):
    s = prepare(live_tensors(m), torch.device("cuda"), train=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    s = prepare(live_tensors(m), torch.device("cuda"), train=True)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("prepare(train=True): host %.3f ms per call, with device work %.3f ms" % ((t1 - t0) / 20 * 1e3, (t2 - t0) / 20 * 1e3))
mouth = torch.rand(16, 512, 50).cuda()
rm = m.refinement_module
for _ in range(3):
    v = rm.video_net.get_block(0)(m.video_bottleneck(mouth))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    v = rm.video_net.get_block(0)(m.video_bottleneck(mouth))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("VP block eager forward (train mode, autograd on): host %.3f ms per call, with device work %.3f ms" % ((t1 - t0) / 20 * 1e3, (t2 - t0) / 20 * 1e3))

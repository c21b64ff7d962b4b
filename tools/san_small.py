import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from conftest import build_model
m = build_model(bench.load_state_dict(), 4, "cuda")
g = torch.Generator().manual_seed(3)
wav = (0.1 * torch.randn(2, 16000, generator=g)).cuda()
lip = torch.rand(2, 512, 25, generator=g).cuda()
with torch.no_grad():
    out = m(wav, lip)
torch.cuda.synchronize()
print("done", float(out.abs().mean()))

"""Runs N forwards of the BASELINE config-2 workload (B=32, 2 s) -- the target of ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:<pat> -s <skip> -c <n> -o gpurun_out/prof python tools/prof_forward.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from conftest import build_model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BATCH
model = build_model(bench.load_state_dict(), bench.REPEATS, "cuda")
wav, lip = bench.make_inputs(B, 1000)
wav, lip = wav.cuda(), lip.cuda()
with torch.no_grad():
    for _ in range(n):
        out = model(wav, lip)
torch.cuda.synchronize()
print("done", float(out.abs().mean()))

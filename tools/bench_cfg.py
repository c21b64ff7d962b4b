"""Throughput of other BASELINE configurations (information only; bench.py measures configs[1]).
    python tools/bench_cfg.py --repeats 12 --batch 64 --seconds 4 [--steps 5]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from conftest import build_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--repeats", type=int, default=12)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=int, default=4)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
L, Tv = 16000 * a.seconds, 25 * a.seconds
g = torch.Generator().manual_seed(7)
wav = (0.1 * torch.randn(a.batch, L, generator=g)).cuda()
lip = torch.rand(a.batch, 512, Tv, generator=g).cuda()
m = build_model(bench.load_state_dict(), a.repeats, "cuda")
with torch.no_grad():
    for _ in range(3):
        out = m(wav, lip)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(a.steps):
        out = m(wav, lip)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
T, Tc = L // 128 + 1, (L // 128 + 1 - 2) // 2 + 1
A, H, G = 4 * 256 * T * 129 * a.batch, 4 * 64 * T * 129 * a.batch, 4 * 64 * Tc * 64 * a.batch
fwd = (6 + 4 * a.repeats) * A + 14 * a.repeats * H + 36 * a.repeats * G
print(f"RTFS-Net-{a.repeats} B={a.batch} {a.seconds}s: {ms:.2f} ms/forward, {a.batch / ms * 1e3:.1f} utt/s, "
      f"forward roofline {(fwd / (ms * 1e-3)) / 1e9:.0f} GB/s algorithmic, finite={bool(torch.isfinite(out).all())}, "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

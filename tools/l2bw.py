"""Read bandwidth vs working-set size (L2-resident vs HBM) with a plain torch reduction."""
import torch
for mb in (8, 16, 32, 64, 96, 128, 256, 1024, 4096):
    x = torch.ones(mb * 1024 * 1024 // 4, device="cuda")
    for _ in range(3):
        x.sum()
    n = max(4, 4096 // mb)
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(n):
        x.sum()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    print(f"read {mb:5d} MB: {mb / 1024 / (ms / 1e3):8.1f} GB/s ({ms * 1e3:.1f} us)")

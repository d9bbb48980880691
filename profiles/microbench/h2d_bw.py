"""Pinned host<->device copy bandwidth of the box (what bounds bench.py's e2e number)."""
import torch
for mb in (4, 32, 128):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
    for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
        for _ in range(3): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} {mb:4d} MB: {10 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s")
# both directions at once on two streams
n = 32 * 1024 * 1024 // 4
h1 = torch.empty(n).pin_memory(); h2 = torch.empty(n).pin_memory(); d1 = torch.empty(n, device="cuda"); d2 = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
print(f"bidirectional 32 MB each way: {10 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s per direction")

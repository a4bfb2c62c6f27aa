"""Pinned host->device copy rate on this box (what bounds bench.py's e2e leg)."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (1, 8, 32, 93, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("H2D %4d MiB: %.3f ms  %.1f GB/s" % (mb, ms, n / ms / 1e6))
# two streams, halves
n = 93 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1):
        d[: n // 2].copy_(h[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2):
        d[n // 2:].copy_(h[n // 2:], non_blocking=True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 100
print("H2D 93 MiB on two streams: %.3f ms  %.1f GB/s" % (ms, n / ms / 1e6))

"""H2D bandwidth of several 360 MB pinned buffers (is the speed a property of the allocation?)."""
import time, torch
n = 360_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
bufs = []
for k in range(10):
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    bufs.append(h)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("buffer", k, round(n / best / 1e9, 2), "GB/s", flush=True)
    if k % 3 == 2:
        bufs.pop(0)

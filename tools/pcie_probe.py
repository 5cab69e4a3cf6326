"""H2D bandwidth of pinned copies vs copy size and number of streams (how much coalescing clouds could buy)."""
import time
import torch
for mb in (1.7, 3.4, 6.8, 13.6, 64.0):
    n = int(mb * 1e6)
    total = 512 * 1024 * 1024
    k = max(1, total // n)
    src = torch.empty(n * min(k, 64), dtype=torch.uint8).pin_memory()
    dst = torch.empty(n * min(k, 64), dtype=torch.uint8, device="cuda")
    for ns in (1, 4):
        streams = [torch.cuda.Stream() for _ in range(ns)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for rep in range(3):
            for i in range(min(k, 64)):
                with torch.cuda.stream(streams[i % ns]):
                    dst[i * n:(i + 1) * n].copy_(src[i * n:(i + 1) * n], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"copy {mb:5.1f} MB x {min(k,64)} x3, {ns} stream(s): {3 * min(k,64) * n / dt / 1e9:6.1f} GB/s")

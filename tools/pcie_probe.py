"""Raw device-to-host bandwidth of the box (pinned memory), for the e2e roofline: one big copy and 64 MB pieces."""
import time
import torch

n = 1600 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for label, piece in (("one copy", n), ("64 MiB pieces", 64 << 20), ("8 MiB pieces", 8 << 20)):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in range(0, n, piece):
            h[a : a + piece].copy_(d[a : a + piece], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(label, round(n / dt / 1e9, 2), "GB/s")

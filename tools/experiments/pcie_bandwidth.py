"""Pinned D2H bandwidth of one PCM step (4096 stereo 20 ms frames = 31.5 MB), whole and chunked, and the
latency of the packet upload.  The floor of bench.py's `e2e` number."""
import time

import torch

n = 4096 * 1920
d = torch.zeros(n, dtype=torch.float32, device="cuda")
h = torch.zeros(n, dtype=torch.float32).pin_memory()
s = torch.cuda.Stream()
for chunks in (1, 4, 8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s):
            for c in range(chunks):
                a, b = n * c // chunks, n * (c + 1) // chunks
                h[a:b].copy_(d[a:b], non_blocking=True)
        s.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(f"D2H {n * 4 / 1e6:.1f} MB in {chunks} chunk(s): {dt * 1e3:.3f} ms = {n * 4 / dt / 1e9:.1f} GB/s")
h2 = torch.zeros(4096 * 160, dtype=torch.uint8).pin_memory()
d2 = torch.zeros(4096 * 160, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
print("H2D 655 KB + sync: %.1f us" % ((time.perf_counter() - t0) / 50 * 1e6))

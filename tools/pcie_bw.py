#!/usr/bin/env python3
"""Pinned host -> device copy bandwidth on this box (the ceiling of bench.py's `e2e`)."""
import torch

n = 1024 * 1280 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("pinned H2D, 1.34 GB in one copy: %.1f GB/s -> at most %.0f frames/s of 1280x1024 u8"
      % (n / best / 1e6, n / best / 1e6 * 1e9 / (1280 * 1024)))
best = 1e9
chunk = 128 * 1280 * 1024
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("pinned H2D, 8 copies of 168 MB: %.1f GB/s" % (n / best / 1e6))

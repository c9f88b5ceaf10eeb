#!/usr/bin/env python3
"""Measured HBM bandwidth of pure-write and pure-read streams on this GPU (torch fill_ / sum),
next to the copy bandwidth MEASURED_PEAKS.json uses: the denominator that applies to a kernel
whose traffic is almost all writes (K1: 1 B/px read, 8 B/px written)."""
import torch

n = 1 << 30  # 4 GiB of f32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


t = timed(lambda: a.fill_(1.0))
print("write-only (fill_ 4 GiB): %.0f GB/s" % (4 * n / t / 1e9))
t = timed(lambda: a.sum())
print("read-only  (sum 4 GiB):   %.0f GB/s" % (4 * n / t / 1e9))
t = timed(lambda: b.copy_(a))
print("copy (read + write bytes): %.0f GB/s" % (8 * n / t / 1e9))

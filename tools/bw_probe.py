#!/usr/bin/env python
"""HBM calibration: pure-write (fill), pure-read (sum) and copy bandwidth, to put write-heavy kernels in context."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e-3
af = a.view(torch.float32)
print(f"fill  (write only): {n / t(lambda: a.zero_()) / 1e9:8.1f} GB/s")
print(f"sum   (read only) : {n / t(lambda: af.sum()) / 1e9:8.1f} GB/s")
print(f"copy  (read+write): {2 * n / t(lambda: b.copy_(a)) / 1e9:8.1f} GB/s")

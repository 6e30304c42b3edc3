#!/usr/bin/env python
"""PSNR reward kernel in isolation: achieved algorithmic GB/s (8 B/pixel) vs the measured HBM peak."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
for B, S in ((64, 256), (512, 256), (4096, 256), (4096, 128), (1024, 512)):
    x = torch.rand(B, S, S, device="cuda"); gt = torch.rand(B, S, S, device="cuda")
    for _ in range(3): ops.psnr(x, gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.psnr(x, gt)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    gbs = 8.0 * B * S * S / t / 1e9
    print(f"psnr B={B:5d} {S}x{S}: {t*1e6:8.1f} us  {gbs:7.1f} GB/s = {100*gbs/peak:5.1f}% of {peak:.0f} GB/s")

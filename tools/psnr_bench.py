#!/usr/bin/env python
"""PSNR reward kernel alone (direct C-ABI calls, CUDA events): us per launch and achieved GB/s (8 B/pixel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import _lib
l = _lib.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for S in (128, 256):
    for B in (1, 8, 64, 512, 4096):
        x = torch.rand(B, 1, S, S, device="cuda"); g = torch.rand(B, 1, S, S, device="cuda"); out = torch.empty(B, device="cuda")
        fn = lambda: l.pnp_psnr(x.data_ptr(), g.data_ptr(), S * S, out.data_ptr(), B, S * S, _lib.stream_ptr())
        for _ in range(5): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(100): fn()
        e1.record(); torch.cuda.synchronize(); t = e0.elapsed_time(e1) / 100 * 1e-3
        ref = 10 * torch.log10(1 / ((x.clamp(0, 1) - g) ** 2).reshape(B, -1).mean(1))
        print(f"psnr {S}x{S} B={B:5d}: {t * 1e6:7.1f} us  {8 * B * S * S / t / 1e9:7.0f} GB/s   max|d| vs torch {float((out - ref).abs().max()):.1e} dB")

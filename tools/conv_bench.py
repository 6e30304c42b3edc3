#!/usr/bin/env python
"""Time one tensor-core conv configuration in isolation (also the target of `ncu --set full`)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=64); ap.add_argument("--s", type=int, default=256)
ap.add_argument("--c0", type=int, default=32); ap.add_argument("--c1", type=int, default=0); ap.add_argument("--cout", type=int, default=32)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)
in0 = torch.randn(a.b, a.s, a.s, a.c0, device="cuda", generator=g).to(torch.bfloat16)
in1 = torch.randn(a.b, a.s, a.s, a.c1, device="cuda", generator=g).to(torch.bfloat16) if a.c1 else None
w = torch.randn(a.cout, a.c0 + a.c1, 3, 3, device="cuda", generator=g) * 0.05
bias = torch.zeros(a.cout, device="cuda")
for _ in range(2): out = ops.conv3x3_bf16(in0, w, bias, in1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(a.iters):
    e0.record(); out = ops.conv3x3_bf16(in0, w, bias, in1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
fl = 2.0 * 9 * (a.c0 + a.c1) * a.cout * a.s * a.s * a.b
t = min(ts)
print(f"conv {a.c0}+{a.c1}->{a.cout} @{a.s} B={a.b}: {t*1e3:.1f} us (incl. weight pack)  {fl/t/1e9:.1f} TFLOP/s")

#!/usr/bin/env python
"""Whole-denoiser timing (no per-launch events): ms per forward over back-to-back launches, best and median of several rounds."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--size", type=int, default=256)
ap.add_argument("--iters", type=int, default=20); ap.add_argument("--rounds", type=int, default=7); ap.add_argument("--tag", default=""); ap.add_argument("--splitk", type=int, default=-1)
a = ap.parse_args()
from dt4image_restoration_b200 import _lib
_lib.lib().pnp_unet_set_splitk(a.splitk)
a.tag = a.tag or f"splitk={a.splitk}"
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to("cuda")
v = torch.rand(a.batch, 1, a.size, a.size, device="cuda"); sg = torch.full((a.batch,), 0.1, device="cuda")
for _ in range(5): den(v, sg)
torch.cuda.synchronize()
ts = []
for r in range(a.rounds):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): den(v, sg)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / a.iters)
    torch.cuda._sleep(int(2e8))          # let the power governor settle between rounds
ts.sort()
print(f"{a.tag} B={a.batch} {a.size}x{a.size}: best {ts[0]:.4f} ms, median {ts[len(ts)//2]:.4f} ms per forward ({a.rounds} rounds of {a.iters})")

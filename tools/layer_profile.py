#!/usr/bin/env python
"""Per-launch timing table of the denoiser (CUDA events around every launch, pnp_unet_profile)."""
import argparse, ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import _lib
from dt4image_restoration_b200.noise import UNetDenoiser2D
from oracle import pnp_oracle as O

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--size", type=int, default=256)
ap.add_argument("--reps", type=int, default=5); ap.add_argument("--out", default="")
a = ap.parse_args()
B, S = a.batch, a.size
den = UNetDenoiser2D(state_dict=O.init_unet_params(0, "default")).to("cuda")
plan = den.plan(B, S, S)
v = torch.rand(B, 1, S, S, device="cuda"); sg = torch.full((B,), 0.1, device="cuda"); x = torch.empty_like(v)
l = _lib.lib()
names, shapes = ["first 2->32"], [(2, 32, 0)]
blocks = [("inc", [(32, 32), (32, 32)], 0), ("down1", [(32, 64), (64, 64), (64, 64)], 1), ("down2", [(64, 128), (128, 128), (128, 128)], 2),
          ("down3", [(128, 256), (256, 256), (256, 256)], 3), ("down4", [(256, 512), (512, 512), (512, 512)], 4),
          ("up1", [(768, 256), (256, 256), (256, 256)], 3), ("up2", [(384, 128), (128, 128), (128, 128)], 2),
          ("up3", [(192, 64), (64, 64), (64, 64)], 1), ("up4", [(96, 32), (32, 32), (32, 32)], 0)]
for bn, convs, lvl in blocks:
    if bn.startswith("down"): names.append(f"{bn}.pool"); shapes.append(None)
    if bn.startswith("up"): names.append(f"{bn}.upsample"); shapes.append(None)
    for i, (ci, co) in enumerate(convs):
        names.append(f"{bn}.conv {ci}->{co} @{S >> lvl}"); shapes.append((ci, co, lvl))
n = C.c_int(64); ms = (C.c_float * 64)(); kinds = (C.c_int * 64)()
acc = np.zeros(64)
for r in range(a.reps + 1):
    _lib.check(l.pnp_unet_profile(plan.handle, v.data_ptr(), sg.data_ptr(), x.data_ptr(), _lib.stream_ptr(), ms, kinds, C.byref(n)))
    if r: acc[:n.value] += np.array(ms[:n.value])
acc /= a.reps
rows = []
tot_f = tot_t = 0.0
for i in range(n.value):
    sh = shapes[i]
    fl = 2.0 * 9 * sh[0] * sh[1] * (S >> sh[2]) ** 2 * B if sh else 0.0
    tf = fl / (acc[i] * 1e-3) / 1e12 if sh else 0.0
    rows.append({"launch": names[i], "kind": int(kinds[i]), "ms": float(acc[i]), "gflop": fl / 1e9, "tflops": tf})
    if kinds[i] == 1: tot_f += fl; tot_t += acc[i]
print(f"{'launch':34s} {'ms':>8s} {'GFLOP':>9s} {'TFLOP/s':>8s} {'share':>6s}")
T = acc[:n.value].sum()
for r in rows: print(f"{r['launch']:34s} {r['ms']:8.4f} {r['gflop']:9.1f} {r['tflops']:8.1f} {100 * r['ms'] / T:5.1f}%")
print(f"total {T:.3f} ms; tcgen05 convs {tot_t:.3f} ms = {tot_f / tot_t / 1e9:.1f} TFLOP/s; image-iters/s (denoiser only) {B / T * 1e3:.0f}")
if a.out: json.dump({"batch": B, "size": S, "rows": rows, "total_ms": float(T), "umma_tflops": tot_f / tot_t / 1e9}, open(a.out, "w"), indent=1)

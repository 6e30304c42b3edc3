#!/usr/bin/env python
"""Where the end-to-end trajectory of bench.py spends its time: upload (reset from pinned host arrays), 30 steps with and
without the per-step action upload / reward download, final download of x.  CUDA events on one stream."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import synth
from dt4image_restoration_b200.engine import PnPEngine
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict

B, S, T = 64, 256, 30
dev = torch.device("cuda")
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to(dev)
batch = synth.make_batch(B, S, S, "cartesian", 4, 0.0, seed0=0)
h_item = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in batch.items()}
sig, mus = synth.fixed_schedule(T)
h_act = torch.stack([torch.tensor(sig, dtype=torch.float32).reshape(T, 1).expand(T, B),
                     torch.tensor(mus, dtype=torch.float32).reshape(T, 1).expand(T, B)], dim=1).contiguous().pin_memory()
h_rew = torch.empty(T, B).pin_memory(); h_x = torch.empty(B, 1, S, S).pin_memory()
eng = PnPEngine(den, B, S, S, dev)

def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e

def run(copies, psnr):
    t = [ev()]
    eng.reset(h_item, non_blocking=True); t.append(ev())
    for k in range(T):
        if copies: eng.actions.copy_(h_act[k], non_blocking=True)
        eng.step()
        if psnr:
            r = eng.psnr()
            if copies: h_rew[k].copy_(r, non_blocking=True)
    t.append(ev())
    h_x.copy_(eng.x, non_blocking=True); t.append(ev())
    torch.cuda.synchronize()
    return [t[i].elapsed_time(t[i + 1]) for i in range(3)]

eng.reset(h_item); eng.set_actions(float(sig[0]), float(mus[0]))
for _ in range(3): run(True, True)
cases = (("steps only", False, False), ("steps + psnr", False, True), ("steps + psnr + copies", True, True))
res = {n: [] for n, _, _ in cases}
for rnd in range(4):                      # alternate the cases: clocks / power state drift between runs of 76 ms
    for name, c, p in cases:
        time.sleep(1.0)
        res[name].append(run(c, p))
for name, _, _ in cases:
    rs = res[name]
    steps = sorted(r[1] for r in rs)
    print(f"{name:24s}: reset/upload {min(r[0] for r in rs):.2f} ms, 30 steps min {steps[0]:.2f} / median {steps[len(steps)//2]:.2f} ms "
          f"({steps[0]/T:.3f} / {steps[len(steps)//2]/T:.3f} per step), x download {min(r[2] for r in rs):.2f} ms")

#!/usr/bin/env python
"""What the reward kernel costs INSIDE a run of environment steps (B=64, 256x256): steps alone, steps + PSNR on the same stream,
steps + PSNR on a side stream (fork after the step, never joined into the next step: x is rewritten 2.4 ms later)."""
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
eng = PnPEngine(den, B, S, S, dev)
eng.reset({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in batch.items()})
eng.set_actions(0.1, 0.5)
side = torch.cuda.Stream()
def run(mode):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(T):
        eng.step()
        if mode == "same":
            eng.psnr()
        elif mode == "side":
            ev = torch.cuda.current_stream().record_event()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                eng.psnr()
        elif mode == "twice":
            eng.psnr(); eng.psnr()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / T
for m in ("none", "same", "side", "twice"):
    run(m)
res = {m: [] for m in ("none", "same", "side", "twice")}
for rnd in range(5):
    for m in res:
        time.sleep(0.7)
        res[m].append(run(m))
for m, v in res.items():
    v = sorted(v)
    print(f"{m:6s}: min {v[0]:.4f}  median {v[len(v)//2]:.4f} ms per step")

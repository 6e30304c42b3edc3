#!/usr/bin/env python
"""Where the DT-driven rollout spends its time: reset (host->device), eager vs graph-replayed iterations."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import synth
from dt4image_restoration_b200.engine import PnPEngine
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict
from dt4image_restoration_b200.policy import DecisionTransformer
from dt4image_restoration_b200.rollout import BatchedRollout
import numpy as np
B, S = 64, 256
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to("cuda")
eng = PnPEngine(den, B, S, S, "cuda")
base = synth.make_batch(8, S, S, "cartesian", 4, 0.0)
data = {k: torch.from_numpy(np.concatenate([v] * 8, axis=0)) for k, v in base.items()}
task = torch.full((B,), 4, dtype=torch.long); rtg0 = (10 + 1.08) / (16.6 + 1.08)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print(f"engine.reset (pageable host tensors): {t(lambda: eng.reset(data)):.1f} ms")
pinned = {k: v.pin_memory() for k, v in data.items()}
print(f"engine.reset (pinned host tensors):   {t(lambda: eng.reset(pinned)):.1f} ms")
dev = {k: v.cuda() for k, v in data.items()}
print(f"engine.reset (device tensors):        {t(lambda: eng.reset(dev)):.1f} ms")
print(f"30 env steps alone:                   {t(lambda: [eng.step() for _ in range(30)]):.1f} ms")
torch.manual_seed(0)
pol = DecisionTransformer()
for g in (False, True):
    ro = BatchedRollout(pol, eng, 6, 30, force_full_length=True, use_graph=g)
    ms = t(lambda: ro.run(dev, task, rtg0))
    print(f"rollout use_graph={g} (graph captured: {ro._graph is not None}): {ms:.1f} ms  -> {B * 30 / ms * 1e3:.0f} image-iters/s")

#!/usr/bin/env python
"""Two DT-driven rollouts (B=64, 256x256, 30 iterations, CUDA graph) - target of an ncu launch list."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import synth
from dt4image_restoration_b200.engine import PnPEngine
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict
from dt4image_restoration_b200.policy import DecisionTransformer
from dt4image_restoration_b200.rollout import BatchedRollout
B, S = 64, 256
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to("cuda")
eng = PnPEngine(den, B, S, S, "cuda")
base = synth.make_batch(8, S, S, "cartesian", 4, 0.0)
dev = {k: torch.from_numpy(np.concatenate([v] * 8, axis=0)).cuda() for k, v in base.items()}
task = torch.full((B,), 4, dtype=torch.long); rtg0 = (10 + 1.08) / (16.6 + 1.08)
torch.manual_seed(0)
ro = BatchedRollout(DecisionTransformer(), eng, 6, 30, force_full_length=True, use_graph=True)
for _ in range(2):
    ro.run(dev, task, rtg0)
torch.cuda.synchronize()
print("done")

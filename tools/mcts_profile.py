#!/usr/bin/env python
"""cProfile of one full tree search (mcts.BatchedMCTS, 128x128, width 5, 30 iterations) on one GPU."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import synth
from dt4image_restoration_b200.mcts import BatchedMCTS
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict
from dt4image_restoration_b200.policy import DecisionTransformer

dev = torch.device("cuda")
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to(dev)
torch.manual_seed(1234)
pol = DecisionTransformer(block_size=18, n_embeds=9, mode="norm")
with torch.no_grad():
    pol.predict_action[0].bias[0] = -2.0
s = BatchedMCTS(pol, den, 128, 128, width=5, n_iters=30, device=dev, rank=0, world=1, peer=None)
itm = synth.make_item(synth.phantom(128, 128, 2), synth.radial_mask(128, 128, 0.3), 0.0, 2)
dm = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in itm.items()}
rtg = (10 + 1.08) / (16.6 + 1.08)
torch.manual_seed(99); s.search(dm, rtg, torch.tensor([[3]]))
torch.cuda.synchronize()
torch.manual_seed(99); t0 = time.perf_counter(); s.env_steps = 0
pr = cProfile.Profile(); pr.enable()
fin, best, progs = s.search(dm, rtg, torch.tensor([[3]]))
torch.cuda.synchronize()
pr.disable()
print(f"search: {time.perf_counter() - t0:.3f} s (under cProfile), env steps {s.env_steps}, final {float(fin):.3f} dB")
pstats.Stats(pr).sort_stats("cumulative").print_stats(int(sys.argv[1]) if len(sys.argv) > 1 else 45)

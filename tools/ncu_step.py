#!/usr/bin/env python
"""One PnP-ADMM step (B x S x S, Cartesian mask -> row-only prox; then the same prox with a radial mask) between
cudaProfilerStart / Stop, for `ncu --profile-from-start off` captures of per-kernel DRAM bytes and durations:

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/step.csv python tools/ncu_step.py [B] [S]
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops, synth
from dt4image_restoration_b200.engine import PnPEngine
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to("cuda")
eng = PnPEngine(den, B, S, S, "cuda")
base = synth.make_batch(min(B, 8), S, S, "cartesian", 4, 0.0, seed0=0)
reps = (B + min(B, 8) - 1) // min(B, 8)
eng.reset({k: torch.from_numpy(np.concatenate([v] * reps, axis=0)[:B]) for k, v in base.items()})
eng.set_actions(0.1, 0.5)
rm = torch.from_numpy(synth.radial_mask(S, S, 0.3)).to("cuda").reshape(1, 1, S, S)
prep_r = ops.ProxPrepared(eng.y0, rm)
zr, ur, vr = torch.empty_like(eng.z), torch.empty_like(eng.u), torch.empty_like(eng.v)
for _ in range(3):
    eng.step(); prep_r.prox_dual(eng.x, eng.u, eng.mu, out=(zr, ur, vr))
print("mask kind known:", eng.probe.get(), prep_r.probe.get())
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step()
prep_r.prox_dual(eng.x, eng.u, eng.mu, out=(zr, ur, vr))
eng.psnr()
torch.cuda.synchronize()
torch.cuda.profiler.stop()

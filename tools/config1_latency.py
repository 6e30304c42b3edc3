#!/usr/bin/env python
"""BASELINE config 1: single 256x256 phantom, radial 30 % mask, 30 PnP-ADMM iterations with the fixed schedule, batch 1,
through the drop-in PnPEnv.reset/step API (the call the reference's eval loop makes) and through the batched engine."""
import os, sys, time
from collections import OrderedDict
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import synth
from dt4image_restoration_b200.env import PnPEnv
from dt4image_restoration_b200.engine import PnPEngine
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict
from oracle import pnp_oracle as O
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to("cuda")
env = PnPEnv(30, den, "cuda")
item = synth.make_item(synth.phantom(S, S, 0), synth.radial_mask(S, S, 0.3), 0.0, 0)
data = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in item.items()}
sig, mus = synth.fixed_schedule(30)
def traj():
    st = env.reset(dict(data), "cuda")
    for k in range(30):
        a = OrderedDict(T=torch.zeros(1, device="cuda"), sigma_d=torch.full((1,), float(sig[k]), device="cuda"),
                        mu=torch.full((1,), float(mus[k]), device="cuda"))
        st, done = env.step(st, a)
    return env.compute_reward(st["x"].reshape(1, S, S), st["gt"].reshape(1, S, S))
traj(); torch.cuda.synchronize()
t0 = time.perf_counter(); n = 5
for _ in range(n): r = traj()
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n
print(f"PnPEnv drop-in, B=1 {S}x{S} radial 30%, 30 iterations: {t*1e3:.1f} ms per trajectory = {t/30*1e3:.3f} ms/iter = {30/t:.0f} image-iters/s; PSNR {float(r):.2f} dB")
def traj_host_actions():
    st = env.reset(dict(data), "cuda")
    for k in range(30):      # actions as host scalars (a fixed schedule): no tensor -> bool synchronisation in `if T > 0.5`
        st, done = env.step(st, OrderedDict(T=0.0, sigma_d=torch.tensor([float(sig[k])]), mu=torch.tensor(float(mus[k]))))
    return env.compute_reward(st["x"].reshape(1, S, S), st["gt"].reshape(1, S, S))
traj_host_actions(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n): r = traj_host_actions()
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n
print(f"PnPEnv drop-in, host-side actions, B=1 {S}x{S}: {t*1e3:.1f} ms per trajectory = {t/30*1e3:.3f} ms/iter = {30/t:.0f} image-iters/s; PSNR {float(r):.2f} dB")
env_eager = PnPEnv(30, den, "cuda", use_graph=False)
def traj_eager():
    st = env_eager.reset(dict(data), "cuda")
    for k in range(30):
        st, done = env_eager.step(st, OrderedDict(T=0.0, sigma_d=torch.tensor([float(sig[k])]), mu=torch.tensor(float(mus[k]))))
    return env_eager.compute_reward(st["x"].reshape(1, S, S), st["gt"].reshape(1, S, S))
traj_eager(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n): r = traj_eager()
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n
print(f"PnPEnv drop-in, use_graph=False, B=1 {S}x{S}: {t*1e3:.1f} ms per trajectory = {t/30*1e3:.3f} ms/iter = {30/t:.0f} image-iters/s; PSNR {float(r):.2f} dB")
eng = PnPEngine(den, 1, S, S, "cuda")
def traj2():
    eng.reset(data)
    for k in range(30):
        eng.set_actions(float(sig[k]), float(mus[k])); eng.step()
    return eng.psnr()
traj2(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n): r = traj2()
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n
print(f"PnPEngine,      B=1 {S}x{S}: {t*1e3:.1f} ms per trajectory = {t/30*1e3:.3f} ms/iter = {30/t:.0f} image-iters/s; PSNR {float(r):.2f} dB")
torch.set_num_threads(os.cpu_count())
params = O.init_unet_params(0, "default")
st = O.reset(data)
t0 = time.perf_counter()
for k in range(30):
    st, _ = O.step(params, st, {"T": torch.zeros(1), "mu": torch.tensor([mus[k]]), "sigma_d": torch.tensor([float(sig[k])])})
t = (time.perf_counter() - t0) / 30
r_cpu = float(O.psnr(st["x"].reshape(1, S, S), st["gt"].reshape(1, S, S)))
print(f"CPU oracle (reference path), {os.cpu_count()} threads: {t*1e3:.1f} ms/iter = {1/t:.1f} image-iters/s; PSNR after 30 iterations {r_cpu:.4f} dB")
# parity of the whole trajectory (north_star: <= 1e-3 max-abs, <= 0.05 dB per trajectory over 30 iterations)
eng.reset(data)
for k in range(30):
    eng.set_actions(float(sig[k]), float(mus[k])); eng.step()
r_gpu = float(eng.psnr()); dx = float((eng.x.reshape(S, S).cpu() - st["x"].reshape(S, S)).abs().max())
print(f"parity after 30 iterations: max|x - oracle| = {dx:.3e}, PSNR {r_gpu:.4f} dB vs {r_cpu:.4f} dB (diff {abs(r_gpu - r_cpu):.2e} dB)")
# the same engine step replayed from a CUDA graph (launch overhead removed)
eng.reset(data); eng.set_actions(float(sig[0]), float(mus[0]))
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    eng.step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eng.step()
def traj3():
    eng.reset(data)
    for k in range(30):
        eng.set_actions(float(sig[k]), float(mus[k])); g.replay()
    return eng.psnr()
traj3(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n): r = traj3()
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n
print(f"PnPEngine + CUDA graph, B=1 {S}x{S}: {t*1e3:.1f} ms per trajectory = {t/30*1e3:.3f} ms/iter = {30/t:.0f} image-iters/s; PSNR {float(r):.2f} dB")

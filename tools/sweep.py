#!/usr/bin/env python
"""BASELINE config 5: throughput sweep over the batch size at 128x128 and 256x256, FFT-prox kernel and U-Net denoiser
timed in isolation on one GPU (the path is embarrassingly parallel over images, so N GPUs run N such sweeps).

    python tools/sweep.py [--max-gb 60]  ->  one table per size: images/s and roofline fraction per batch
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops, _lib
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict

ap = argparse.ArgumentParser(); ap.add_argument("--max-gb", type=float, default=60.0); ap.add_argument("--out", default="")
a = ap.parse_args()
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
GF = {128: 9.684, 256: 38.734}
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to("cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timeit(fn, it):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3

rows = []
for S in (128, 256):
    print(f"--- {S}x{S} ---")
    print(f"{'batch':>6s} | {'prox Cartesian us':>18s} {'Mimg/s':>8s} {'HBM frac':>8s} | {'prox radial us':>15s} {'HBM frac':>8s} | {'U-Net ms':>9s} {'kimg/s':>8s} {'TFLOP/s':>8s} {'of bf16 peak':>12s}")
    B = 1
    while B <= 4096:
        g = torch.Generator(device="cuda").manual_seed(B)
        x = torch.rand(B, 1, S, S, device="cuda", generator=g)
        u = torch.complex(torch.randn(B, 1, S, S, device="cuda", generator=g), torch.randn(B, 1, S, S, device="cuda", generator=g)) * 0.1
        y0 = torch.complex(torch.randn(B, 1, S, S, device="cuda", generator=g), torch.randn(B, 1, S, S, device="cuda", generator=g))
        mu = torch.full((B,), 0.5, device="cuda")
        out = (torch.empty_like(u), torch.empty_like(u), torch.empty_like(x))
        res = {"size": S, "batch": B}
        for kind in ("cartesian", "radial"):
            if kind == "cartesian":
                mask = (torch.rand(B, 1, 1, S, device="cuda", generator=g) < 0.25).expand(B, 1, S, S).contiguous()
            else:
                mask = torch.rand(B, 1, S, S, device="cuda", generator=g) < 0.25
            prep = ops.ProxPrepared(y0, mask)
            t = timeit(lambda: prep.prox_dual(x, u, mu, out=out), 20 if B <= 256 else 5)
            res[f"prox_{kind}_us"] = t * 1e6
            res[f"prox_{kind}_hbm_frac"] = 37.0 * B * S * S / t / 1e9 / peaks["hbm_gbs"]
            del prep
        ws_gb = _lib.lib().pnp_unet_workspace_bytes(B, S, S) / 1e9
        if ws_gb <= a.max_gb:
            plan = den.plan(B, S, S)
            v = torch.rand(B, 1, S, S, device="cuda"); sg = torch.full((B,), 0.1, device="cuda")
            t = timeit(lambda: plan.forward(v, sg), 10 if B <= 64 else 3)
            res["unet_ms"] = t * 1e3
            res["unet_tflops"] = GF[S] * B / t / 1e3
            den._plans.clear(); del plan
            torch.cuda.empty_cache()
        un = f"{res['unet_ms']:9.3f} {B / res['unet_ms']:8.2f} {res['unet_tflops']:8.1f} {res['unet_tflops'] / peaks['bf16_tflops_sustained']:12.3f}" if "unet_ms" in res else f"{'(workspace > ' + str(int(a.max_gb)) + ' GB)':>40s}"
        print(f"{B:6d} | {res['prox_cartesian_us']:18.1f} {B / res['prox_cartesian_us']:8.3f} {res['prox_cartesian_hbm_frac']:8.3f} | "
              f"{res['prox_radial_us']:15.1f} {res['prox_radial_hbm_frac']:8.3f} | {un}")
        rows.append(res)
        B *= 4 if B >= 16 else 2
if a.out:
    json.dump(rows, open(a.out, "w"), indent=1)

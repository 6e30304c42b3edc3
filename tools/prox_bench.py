#!/usr/bin/env python
"""FFT-prox + dual kernel in isolation: time, image-iters/s, achieved algorithmic GB/s (37 B/pixel) vs HBM peak.
Cases BxS[c|r]: c = Cartesian 4x mask (column-only -> row-only kernel), r = random mask (general cluster kernel)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops, _lib
ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cases", default="64x256c,256x256c,1024x256c,64x256r,256x256r,1024x256r,1024x128r,128x512r")
a = ap.parse_args()
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
for cs in a.cases.split(","):
    kind = cs[-1] if cs[-1] in "cr" else "r"
    B, S = map(int, cs.rstrip("cr").split("x"))
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(B, 1, S, S, device="cuda", generator=g)
    u = torch.complex(torch.randn(B, 1, S, S, device="cuda", generator=g), torch.randn(B, 1, S, S, device="cuda", generator=g)) * 0.1
    y0 = torch.complex(torch.randn(B, 1, S, S, device="cuda", generator=g), torch.randn(B, 1, S, S, device="cuda", generator=g))
    if kind == "c":
        cols = torch.rand(B, 1, 1, S, device="cuda", generator=g) < 0.25
        mask = cols.expand(B, 1, S, S).contiguous()
    else:
        mask = (torch.rand(B, 1, S, S, device="cuda", generator=g) < 0.25)
    mu = torch.full((B,), 0.5, device="cuda")
    z = torch.empty_like(u); un = torch.empty_like(u); v = torch.empty_like(x)
    if ops.ProxPrepared.supported(S, S):
        prep = ops.ProxPrepared(y0, mask)
        tag = " (prepared, row-only kernel)" if prep.column_only else (" (prepared, cluster kernel)" if S in (128, 256) else " (prepared, general 3-launch)")
        run = lambda: prep.prox_dual(x, u, mu, out=(z, un, v))
    else:
        ws = torch.empty(_lib.lib().pnp_prox_workspace_bytes(B, S, S), dtype=torch.uint8, device="cuda")
        tag = " (general 3-launch)"
        run = lambda: ops.prox_dual(x, u, y0, mask, mu, out=(z, un, v), workspace=ws)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): run()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / a.iters * 1e-3
    gbs = 37.0 * B * S * S / t / 1e9
    print(f"prox{tag} B={B:5d} {S}x{S}: {t*1e6:9.1f} us  {B/t/1e6:7.3f} M image-iters/s  {gbs:7.1f} GB/s algorithmic = {100*gbs/peak:5.1f}% of {peak:.0f} GB/s")

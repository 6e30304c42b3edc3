#!/usr/bin/env python
"""FFT-prox + dual kernel in isolation: time, image-iters/s, achieved algorithmic GB/s (37 B/pixel) vs HBM peak."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops, _lib
ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=20); ap.add_argument("--cases", default="64x256,256x128,1024x128,256x256,1024x256,32x512,128x512,4096x128")
a = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
for cs in a.cases.split(","):
    B, S = map(int, cs.split("x"))
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(B, 1, S, S, device="cuda", generator=g)
    u = torch.complex(torch.randn(B, 1, S, S, device="cuda", generator=g), torch.randn(B, 1, S, S, device="cuda", generator=g)) * 0.1
    y0 = torch.complex(torch.randn(B, 1, S, S, device="cuda", generator=g), torch.randn(B, 1, S, S, device="cuda", generator=g))
    mask = (torch.rand(B, 1, S, S, device="cuda", generator=g) < 0.25)
    mu = torch.full((B,), 0.5, device="cuda")
    z = torch.empty_like(u); un = torch.empty_like(u); v = torch.empty_like(x)
    ws = torch.empty(_lib.lib().pnp_prox_workspace_bytes(B, S, S), dtype=torch.uint8, device="cuda")
    l = _lib.lib()
    prepared = bool(l.pnp_prox_prepared_supported(S, S))
    if prepared:
        y0T = torch.empty_like(y0); mT = torch.empty(B, 1, S, S, dtype=torch.uint8, device="cuda")
        mk8 = mask.view(torch.uint8)
        _lib.check(l.pnp_prox_prepare(y0.data_ptr(), mk8.data_ptr(), S * S, y0T.data_ptr(), mT.data_ptr(), B, S, S, _lib.stream_ptr()))
        def run():
            _lib.check(l.pnp_prox_dual_prepared(x.data_ptr(), u.data_ptr(), y0T.data_ptr(), mT.data_ptr(), S * S, mu.data_ptr(), 1,
                                                z.data_ptr(), un.data_ptr(), v.data_ptr(), B, S, S, _lib.stream_ptr()))
    else:
        def run(): ops.prox_dual(x, u, y0, mask, mu, out=(z, un, v), workspace=ws)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): run()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / a.iters * 1e-3
    gbs = 37.0 * B * S * S / t / 1e9
    print(f"prox{' (prepared)' if prepared else ''} B={B:5d} {S}x{S}: {t*1e6:9.1f} us  {B/t/1e6:7.3f} M image-iters/s  {gbs:7.1f} GB/s algorithmic = {100*gbs/peak:5.1f}% of {peak:.0f} GB/s")

#!/usr/bin/env python
"""Stress of the cluster FFT-prox kernels against torch.fft ON THE GPU (test infrastructure; not a product path): many rounds per
cluster, repeated launches with u_out aliasing u_in, per-image random masks and mu.  Any protocol race (exchange buffers, credits,
bulk-load prefetch, TMEM slots) shows up as a mismatch in some image of some repeat."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops
def cfft(z, inv=False):
    f = torch.fft.ifftn if inv else torch.fft.fftn
    return torch.fft.fftshift(f(torch.fft.ifftshift(z, dim=(-2, -1)), dim=(-2, -1), norm="ortho"), dim=(-2, -1))
worst = 0.0
for S, B in ((256, 333), (256, 57), (128, 1500), (128, 301)):
    g = torch.Generator(device="cuda").manual_seed(S + B)
    x = torch.rand(B, 1, S, S, device="cuda", generator=g)
    u = torch.complex(torch.randn(B, 1, S, S, device="cuda", generator=g), torch.randn(B, 1, S, S, device="cuda", generator=g)) * 0.1
    y0 = torch.complex(torch.randn(B, 1, S, S, device="cuda", generator=g), torch.randn(B, 1, S, S, device="cuda", generator=g))
    mask = torch.rand(B, 1, S, S, device="cuda", generator=g) < 0.3
    mu = torch.rand(B, device="cuda", generator=g) * 0.9 + 0.05
    prep = ops.ProxPrepared(y0, mask)
    z, v = torch.empty_like(u), torch.empty_like(x)
    u_ref = u.clone()
    for rep in range(8):
        Z = cfft(x + u_ref)
        m4 = mu.view(B, 1, 1, 1)
        Z = torch.where(mask, (m4 * Z + y0) / (1 + m4), Z)
        z_ref = cfft(Z, inv=True)
        u_ref = u_ref + x - z_ref
        prep.prox_dual(x, u, mu, out=(z, u, v))          # u updated in place (u_out aliases u_in)
        dz = (z - z_ref).abs().amax(dim=(1, 2, 3)); du = (u - u_ref).abs().amax(dim=(1, 2, 3))
        dv = (v - (z_ref - u_ref).real).abs().amax(dim=(1, 2, 3))
        w = max(dz.max().item(), du.max().item(), dv.max().item())
        worst = max(worst, w)
        assert w < 2e-4, f"S={S} B={B} repeat {rep}: image {int(torch.argmax(dz))} off by {w:.3e}"
        x = (x * 0.9 + 0.1 * z_ref.real).clamp(0, 1)      # a different x every repeat
    print(f"S={S} B={B}: 8 repeats ok")
print(f"prox_stress: ok, worst |diff| {worst:.2e}")

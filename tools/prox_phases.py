#!/usr/bin/env python
"""Per-phase cycle breakdown of fftprox_cl_kernel (general-mask FFT-prox at 256x256, 16-CTA clusters).

    python tools/prox_phases.py --build      # here (no GPU needed): csrc/libpnp_b200_phases.so with -DPNP_PROX_PHASE_TIMING
    python tools/prox_phases.py [B ...]      # on a B200: average cycles per image and CTA for every phase

The instrumented library is a separate file; the product library is not touched.  Thread 0 of every CTA reads %clock64
at the phase boundaries, so a phase's figure is that warp's time in it, waits included.
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dt4image_restoration_b200 import build as B  # noqa: E402

LIB = os.path.join(B.CSRC, "libpnp_b200_phases.so")
PHASES = ["wait for the bulk loads of u, x (next image)", "wait for the A-free credits of the peers",
          "rows forward of the next image (registers)", "wait for the peers' rows (exchange 1)",
          "columns (FFT, blend, inverse FFT) in place", "column sends (st.async)", "wait for the peers' columns (exchange 2)",
          "row sends (st.async) + CTA barrier + next bulk loads", "rows inverse + epilogue"]


def build():
    objs = []
    for s in B.SOURCES:
        o = os.path.join(B.CSRC, s.replace(".cu", ".phases.o"))
        subprocess.check_call([B._nvcc(), *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")], "-DPNP_PROX_PHASE_TIMING",
                               "-c", os.path.join(B.CSRC, s), "-o", o])
        objs.append(o)
    subprocess.check_call([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart"])
    print(LIB)


def run(batches):
    import torch
    from dt4image_restoration_b200 import _lib
    _lib.LIB_PATH = LIB                      # before the first load(): every op of this process uses the instrumented build
    from dt4image_restoration_b200 import ops
    lib = _lib.lib()
    lib.pnp_debug_prox_phases.restype = C.c_int
    lib.pnp_debug_prox_phases.argtypes = [C.POINTER(C.c_ulonglong)]
    out = (C.c_ulonglong * 16)()
    S = 256
    for Bn in batches:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.rand(Bn, 1, S, S, device="cuda", generator=g)
        cplx = lambda: torch.complex(torch.randn(Bn, 1, S, S, device="cuda", generator=g), torch.randn(Bn, 1, S, S, device="cuda", generator=g))
        u, y0 = cplx() * 0.1, cplx()
        mask = torch.rand(Bn, 1, S, S, device="cuda", generator=g) < 0.25
        mu = torch.full((Bn,), 0.5, device="cuda")
        z, un, v = torch.empty_like(u), torch.empty_like(u), torch.empty_like(x)
        prep = ops.ProxPrepared(y0, mask)
        for _ in range(3):
            prep.prox_dual(x, u, mu, out=(z, un, v))
        _lib.check(lib.pnp_debug_prox_phases(out))            # clear
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            prep.prox_dual(x, u, mu, out=(z, un, v))
        e1.record()
        _lib.check(lib.pnp_debug_prox_phases(out))
        cta_images = max(int(out[15]), 1)                      # one count per CTA and image
        names = PHASES
        tot = sum(int(out[i]) for i in range(len(names)))
        print(f"B={Bn} 256x256 random 25 % mask: {e0.elapsed_time(e1) / n * 1e3:.1f} us per launch (instrumented build), "
              f"{tot / cta_images:.0f} cycles per image and CTA")
        for i, name in enumerate(names):
            if name is None:
                continue
            print(f"  {int(out[i]) / cta_images:9.0f} clk  {100.0 * int(out[i]) / max(tot, 1):5.1f} %  {name}")


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    else:
        run([int(a) for a in sys.argv[1:]] or [64, 1024])

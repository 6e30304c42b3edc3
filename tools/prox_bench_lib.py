#!/usr/bin/env python
"""tools/prox_bench.py against another build of the library (A/B of kernel variants in ONE gpurun call):

    python tools/prox_bench_lib.py libpnp_b200_<tag>.so --cases 64x256r,1024x256r

The variant libraries are built here with the product flags plus -D switches (e.g. -DPNP_CL_LOCAL_ONLY: every cluster exchange
of fftprox_cl_kernel targets the sender's own CTA - wrong results, no SM-to-SM traffic: how much of the time is the fabric)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dt4image_restoration_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.__file__), "csrc", sys.argv[1])
sys.argv = ["prox_bench.py"] + sys.argv[2:]
runpy.run_path(os.path.join(ROOT, "tools", "prox_bench.py"), run_name="__main__")

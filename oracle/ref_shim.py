"""Import the real reference (``/root/reference``) read-only.  TEST INFRASTRUCTURE ONLY.

Works only in the build container (the GPU box has no ``/root/reference``); used by
``oracle/make_golden.py`` and by the ``ref``-marked tests, which skip when the tree is absent.

Shims:
* ``skimage.metrics`` and ``h5py`` are not installed and are only used by dead / off-path code
  (reference ``evaluation/env.py:7,141-143``, ``dataset/datasets.py:8``) -> stub modules.
* ``PnPEnv._load_no_ref`` downloads ARNIQA through ``torch.hub`` (``env.py:36-40``) -> no-op.
* ``UNetDenoiser2D`` insists on a checkpoint file (``noise.py:140-148``) -> we save the seeded
  state_dict to a temp file and pass ``ckpt_path``.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

REF_ROOT = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "evaluation", "env.py"))


def load():
    """Returns a namespace with the reference's PnPEnv, UNet, UNetDenoiser2D, fft, ifft, torch_psnr."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        skm = types.ModuleType("skimage.metrics")
        skm.peak_signal_noise_ratio = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
        sk.metrics = skm
        sys.modules["skimage"] = sk
        sys.modules["skimage.metrics"] = skm
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import evaluation.env as renv
    import evaluation.noise as rnoise
    import evaluation.utils.transformations as rtr

    renv.PnPEnv._load_no_ref = lambda self: None
    ns = types.SimpleNamespace(PnPEnv=renv.PnPEnv, torch_psnr=renv.torch_psnr, UNet=rnoise.UNet,
                               UNetDenoiser2D=rnoise.UNetDenoiser2D, fft=rtr.fft, ifft=rtr.ifft,
                               env_module=renv, noise_module=rnoise)
    return ns


def make_denoiser(ns, params):
    """Reference ``UNetDenoiser2D`` holding the given state_dict."""
    import torch
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "unet.pt")
        torch.save({k: v.clone() for k, v in params.items()}, p)
        return ns.UNetDenoiser2D(ckpt_path=p)

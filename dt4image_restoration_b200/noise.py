"""Drop-in for the reference denoiser ``UNetDenoiser2D`` (reference ``evaluation/noise.py:139-164``).

Same constructor (``ckpt_path``), same checkpoint format (the ``state_dict`` of reference ``UNet(2, 1)``,
56 tensors, ``noise.py:147-148``), same call ``denoiser(x[B,1,H,W], sigma) -> [B,1,H,W]`` in [0,1].
The network itself (``noise.py:101-133``) runs as hand-written sm_100a kernels behind the C-ABI
(``pnp_unet_forward``): tcgen05 implicit-GEMM convs with fused epilogues; see DESIGN.md.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import torch

from . import ops

CURRENT_DIR = os.path.dirname(os.path.abspath(__file__))


def random_init_state_dict(seed: int = 0, kind: str = "default") -> "OrderedDict[str, torch.Tensor]":
    """Seeded synthetic weights in the reference checkpoint format (no checkpoint is shipped with the reference,
    ``.gitignore:2,6,7``; BASELINE's configs all say "random-init U-Net").

    ``default``: what ``nn.Conv2d`` draws for the reference's ``UNet(2, 1)`` (``noise.py:80``): weight and bias
    uniform in +-1/sqrt(fan_in).  ``kaiming``: variance-preserving normal weights for LeakyReLU(0.2), biases
    N(0, 0.02^2), the 1x1 output conv damped by 0.15 (a net whose layers all carry signal, used by the layer tests).
    One CPU generator, tensors drawn in registration order in float64 and rounded to fp32, so the values depend on
    ``(seed, kind)`` only.  ``tests/test_oracle_golden.py`` pins it to the oracle's initialiser bit for bit.
    """
    if kind not in ("default", "kaiming"):
        raise ValueError(kind)
    gen = torch.Generator().manual_seed(int(seed) * 2654435761 % (2 ** 31) + 12345)
    sd, fan_in = OrderedDict(), 1
    for key, shape in ops.unet_state_dict_shapes():
        is_weight = key.endswith("weight")
        if is_weight:
            fan_in = shape[1] * shape[2] * shape[3]      # the bias that follows belongs to this conv
        if kind == "default":
            t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * (1.0 / math.sqrt(fan_in))
        elif is_weight:
            std = math.sqrt(2.0 / ((1.0 + 0.2 ** 2) * fan_in)) * (0.15 if key.startswith("outc") else 1.0)
            t = torch.randn(shape, generator=gen, dtype=torch.float64) * std
        else:
            t = torch.randn(shape, generator=gen, dtype=torch.float64) * 0.02
        sd[key] = t.to(torch.float32)
    return sd


class UNetDenoiser2D(torch.nn.Module):
    def __init__(self, ckpt_path=None, state_dict=None):
        super().__init__()
        if state_dict is None:
            if ckpt_path is None:
                ckpt_path = os.path.join(CURRENT_DIR, "pretrained", "unet-nm.pt")
                if not os.path.exists(ckpt_path):
                    raise ValueError("Default ckpt not found, you have to provide a ckpt path")
            state_dict = torch.load(ckpt_path, map_location="cpu")
        self.register_buffer("flat_params", ops.flatten_state_dict(state_dict), persistent=False)
        self._packed = None
        self._packed_key = None
        self._plans = {}

    # -- device-side preparation -----------------------------------------------------------------
    def _ensure_packed(self, device):
        key = (self.flat_params.data_ptr(), str(device))
        if self._packed is None or self._packed_key != key:
            if not self.flat_params.is_cuda:
                raise ops._lib.PnpError("UNetDenoiser2D has no CPU path: move it to a CUDA device with .to('cuda')")
            self._packed = ops.pack_unet_weights(self.flat_params)
            self._packed_key = key
            self._plans = {}
        return self._packed

    def plan(self, B: int, H: int, W: int) -> ops.UNetPlan:
        packed = self._ensure_packed(self.flat_params.device)
        k = (B, H, W)
        if k not in self._plans:
            self._plans[k] = ops.UNetPlan(packed, B, H, W)
            while len(self._plans) > 6:                   # a plan owns up to 8 GiB of workspace: keep the six newest shapes
                self._plans.pop(next(iter(self._plans)))
        else:
            self._plans[k] = self._plans.pop(k)           # most recently used last
        return self._plans[k]

    # -- reference interface ----------------------------------------------------------------------
    def forward(self, x, sigma, preclamp: bool = False):
        # x: [B,1,H,W]
        N, C, H, W = x.shape
        if C != 1:
            raise RuntimeError(f"expected a single-channel image, got C={C}")
        sigma = torch.as_tensor(sigma, dtype=torch.float32, device=x.device).reshape(-1)
        if sigma.numel() != N:
            # the reference's sigma.view(N,1,1,1) raises for a size mismatch (noise.py:159)
            raise RuntimeError(f"shape '[{N}, 1, 1, 1]' is invalid for input of size {sigma.numel()}")
        plan = self.plan(N, H, W)
        if preclamp:
            out, pre = plan.forward(x.float(), sigma, preclamp=True)
            return out.view(N, 1, H, W), pre.view(N, 1, H, W)
        return plan.forward(x.float(), sigma).view(N, 1, H, W)

"""Drop-in for the reference denoiser ``UNetDenoiser2D`` (reference ``evaluation/noise.py:139-164``).

Same constructor (``ckpt_path``), same checkpoint format (the ``state_dict`` of reference ``UNet(2, 1)``,
56 tensors, ``noise.py:147-148``), same call ``denoiser(x[B,1,H,W], sigma) -> [B,1,H,W]`` in [0,1].
The network itself (``noise.py:101-133``) runs as hand-written sm_100a kernels behind the C-ABI
(``pnp_unet_forward``): tcgen05 implicit-GEMM convs with fused epilogues; see DESIGN.md.
"""
from __future__ import annotations

import os

import torch

from . import ops

CURRENT_DIR = os.path.dirname(os.path.abspath(__file__))


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

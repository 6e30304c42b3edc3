"""Pin ``PnPEnv.no_ref_inputs`` / ``run_no_ref_reward`` against the REAL reference ``run_no_ref_reward`` (``evaluation/env.py:42-54``)
and write ``tests/golden/ref_noref_inputs.npz``.  TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python -m oracle.make_golden_noref

ARNIQA itself (torch.hub, network) is not available; a recording stand-in module with ARNIQA's call signature is installed on
the reference environment, the reference's own method is run on a 128 x 128 state, and what the model RECEIVED (full- and
half-resolution 3-channel images) and the returned score are saved.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dt4image_restoration_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class RecordingModel(torch.nn.Module):
    """Stand-in with ARNIQA's forward signature: a fixed functional of both inputs as the 'score', inputs kept."""

    def forward(self, img, img_ds, return_embedding=False, scale_score=True):
        assert return_embedding is False and scale_score is True
        self.seen = (img.detach().clone(), img_ds.detach().clone())
        w = torch.linspace(0.5, 1.5, img.shape[-1], dtype=torch.float32)
        wd = torch.linspace(1.5, 0.5, img_ds.shape[-1], dtype=torch.float32)
        return ((img.float() * w).mean(dim=(1, 2, 3)) + 2.0 * (img_ds.float() * wd).mean(dim=(1, 2, 3))).reshape(-1, 1)


def main():
    ns = ref_shim.load()
    env = ns.PnPEnv(30, torch.nn.Identity(), "cpu")
    env.no_ref_model = RecordingModel().eval()
    x = torch.from_numpy(synth.phantom(128, 128, 7)).float().reshape(1, 1, 128, 128)
    x = x + 0.05 * torch.randn(x.shape, generator=torch.Generator().manual_seed(3))
    score = env.run_no_ref_reward({"x": x})
    img, img_ds = env.no_ref_model.seen
    assert img.shape == (1, 3, 128, 128) and img_ds.shape == (1, 3, 64, 64)
    np.savez_compressed(os.path.join(GOLD, "ref_noref_inputs.npz"), x=x.numpy(), img=img.numpy(), img_ds=img_ds.numpy(),
                        score=np.float64(score))
    print("score", score, "->", os.path.join(GOLD, "ref_noref_inputs.npz"),
          f"{os.path.getsize(os.path.join(GOLD, 'ref_noref_inputs.npz')) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()

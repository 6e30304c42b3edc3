"""Pin the oracle against the REAL reference ``PnPEnv.step`` at sizes that are NOT powers of two and write
``tests/golden/ref_env_anysize.npz``.  TEST INFRASTRUCTURE ONLY.  Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden_anysize

The reference's ``reset`` / reward hard-code 128 x 128 (``evaluation/env.py:64,115``), but its ``step`` (``env.py:74-100``) is
shape-agnostic: ``torch.fft`` is mixed-radix and the U-Net pads ragged levels (``noise.py:49-53``).  SURVEY 8a lists 130 x 130
and 136 x 120 as sizes it accepts; an odd x odd size is added because there ``ifftshift`` and ``fftshift`` differ
(``transformations.py:6-19``).  Three steps each on the reference's own ``PnPEnv`` / ``UNetDenoiser2D`` (default init, seed 0),
states built with the reference's reset semantics; the oracle is asserted bit-identical (max|d| = 0) before anything is saved.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dt4image_restoration_b200 import synth  # noqa: E402
from oracle import pnp_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import act, eq, ref_states_any_shape  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
# (H, W, mask kind, mask parameter, k-space noise sigma, seed)
CASES = [(130, 130, "radial", 0.3, 0.0, 4), (136, 120, "cartesian", 4, 5.0, 5), (45, 51, "radial", 0.4, 0.0, 6)]
ACTIONS = [(0.0, 0.30, 40.0 / 255), (0.1, 0.55, 20.0 / 255), (0.2, 0.90, 8.0 / 255)]


def make_case_item(H, W, kind, par, sn, seed):
    mask = synth.radial_mask(H, W, par) if kind == "radial" else synth.cartesian_mask(H, W, par, seed)
    return synth.make_item(synth.phantom(H, W, seed), mask, sn, seed)


def main():
    torch.set_num_threads(1)
    ns = ref_shim.load()
    params = O.init_unet_params(seed=0, kind="default")
    env = ns.PnPEnv(30, ref_shim.make_denoiser(ns, params), "cpu")
    out = {}
    for (H, W, kind, par, sn, seed) in CASES:
        item = make_case_item(H, W, kind, par, sn, seed)
        rs = ref_states_any_shape(ns, item)
        os_ = O.reset(item)
        xs = []
        for (T, mu, sg) in ACTIONS:
            rs, rdone = env.step(rs, act(T, mu, sg))
            os_, odone = O.step(params, os_, act(T, mu, sg))
            assert rdone == odone is False
            eq(rs["x"], os_["x"], f"{H}x{W} step.x")
            eq(torch.view_as_real(rs["z"]), torch.view_as_real(os_["z"]), f"{H}x{W} step.z")
            eq(torch.view_as_real(rs["u"]), torch.view_as_real(os_["u"]), f"{H}x{W} step.u")
            xs.append(rs["x"].numpy().copy())
        tag = f"{H}x{W}"
        out[f"x_steps_{tag}"] = np.stack(xs)
        out[f"z_{tag}"] = torch.view_as_real(rs["z"]).numpy()
        out[f"u_{tag}"] = torch.view_as_real(rs["u"]).numpy()
        print(f"{tag}: oracle == reference over {len(ACTIONS)} steps (max|d| = 0)")
    out["actions"] = np.array(ACTIONS, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "ref_env_anysize.npz"), **out)
    print("written", os.path.join(GOLD, "ref_env_anysize.npz"),
          f"{os.path.getsize(os.path.join(GOLD, 'ref_env_anysize.npz')) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()

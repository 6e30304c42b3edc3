"""No-reference reward hook (reference ``evaluation/env.py:36-54``): the drop-in's preprocessing and model call against what
the REAL reference's ``run_no_ref_reward`` handed to a recording stand-in model (``oracle/make_golden_noref.py``)."""
import os

import numpy as np
import pytest
import torch

from dt4image_restoration_b200.env import PnPEnv
from oracle.make_golden_noref import RecordingModel


def _env(model):
    env = PnPEnv.__new__(PnPEnv)          # no denoiser / device needed for the reward hook
    env.no_ref_model = model
    return env


def test_no_ref_inputs_and_score_match_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_noref_inputs.npz"))
    state = {"x": torch.from_numpy(g["x"])}
    img, img_ds = PnPEnv.no_ref_inputs(state)
    assert img.shape == (1, 3, 128, 128) and img_ds.shape == (1, 3, 64, 64)
    assert np.array_equal(img.numpy(), g["img"])                      # grey channel + two zero channels: bit-exact
    np.testing.assert_allclose(img_ds.numpy(), g["img_ds"], rtol=0, atol=1e-6)   # antialiased bilinear x1/2
    m = RecordingModel().eval()
    score = _env(m).run_no_ref_reward(state)
    assert isinstance(score, float) and abs(score - float(g["score"])) < 1e-6
    assert torch.equal(m.seen[0], img)


def test_no_ref_hook_variants():
    with pytest.raises(NotImplementedError):
        _env(None).run_no_ref_reward({"x": torch.zeros(1, 1, 128, 128)})
    assert _env(lambda st: 3).run_no_ref_reward({"x": torch.zeros(1, 1, 128, 128)}) == 3.0   # plain callable(state)
    # complex x (the state right after reset, env.py:59) and another size
    st = {"x": torch.complex(torch.rand(1, 1, 96, 80), torch.zeros(1, 1, 96, 80))}
    img, img_ds = PnPEnv.no_ref_inputs(st)
    assert img.shape == (1, 3, 96, 80) and img_ds.shape == (1, 3, 48, 40) and float(img[:, 1:].abs().max()) == 0.0


@pytest.mark.gpu
def test_no_ref_reward_on_cuda_state(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_noref_inputs.npz"))

    class M(RecordingModel):
        def forward(self, img, img_ds, return_embedding=False, scale_score=True):
            return super().forward(img.float().cpu(), img_ds.float().cpu(), return_embedding, scale_score).to(img.device)

    score = _env(M().eval()).run_no_ref_reward({"x": torch.from_numpy(g["x"]).cuda()})
    assert abs(score - float(g["score"])) < 1e-4

"""The action-producing policy (reference transformer/decision_transformer.py) against reference fixtures."""
import os

import numpy as np
import torch

from dt4image_restoration_b200.policy import DecisionTransformer


def _inputs():
    g = torch.Generator().manual_seed(55)
    B, K = 1, 6
    rtg = torch.rand(B, K, 1, generator=g)
    st = torch.rand(B, K, 128 * 128, generator=g)
    ts = torch.arange(K).reshape(1, K, 1)
    task = torch.full((B, K), 3, dtype=torch.long)
    acts = torch.rand(B, K, 3, generator=g)
    return rtg, st, ts, task, acts


def test_policy_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_dt_seed1234.npz"))
    torch.manual_seed(1234)
    m = DecisionTransformer(block_size=18, n_embeds=9, mode="norm")
    assert sum(p.numel() for p in m.parameters()) == 1297836
    rtg, st, ts, task, acts = _inputs()
    a, ad = m(rtg, st, ts, task, acts, eval_actions=True)
    assert list(ad) == ["T", "sigma_d", "mu"]                      # key order of mode 'norm' (reference :147-154)
    np.testing.assert_allclose(a.numpy(), g["actions"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(m(rtg, st, ts, task, acts, eval_rtg=True).numpy(), g["rtg"], rtol=0, atol=5e-6)
    a0, _ = m(rtg, st, ts, task, actions=None)
    np.testing.assert_allclose(a0.numpy(), g["actions_noact"], rtol=0, atol=5e-6)
    assert float(ad["sigma_d"].max()) <= 70 / 255 and float(ad["T"].max()) <= 1.0


def test_policy_flex_key_order_and_resampling():
    torch.manual_seed(0)
    m = DecisionTransformer(mode="flex")
    rtg, st, ts, task, acts = _inputs()
    _, ad = m(rtg, st, ts, task, acts, eval_actions=True)
    assert list(ad) == ["mu", "sigma_d", "T"]
    # 256x256 observations are area-resampled to the encoder's 128x128: a constant image maps to the same output
    big = torch.full((1, 6, 256 * 256), 0.37)
    small = torch.full((1, 6, 128 * 128), 0.37)
    a_big, _ = m(rtg, big, ts, task, acts, eval_actions=True, hw=(256, 256))
    a_small, _ = m(rtg, small, ts, task, acts, eval_actions=True)
    assert (a_big - a_small).abs().max() < 1e-6

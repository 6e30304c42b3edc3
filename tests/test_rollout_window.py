"""Host logic of the batched policy rollout on CPU: the static K-entry context window with zero
padding and device-side position / shift must give the same actions as the reference-style growing / sliding window
(``evaluation/eval.py:147-220``: the policy sees the last K triples).  The environment is replaced by a toy engine."""
import copy

import pytest
import torch
import torch.nn.functional as F

from dt4image_restoration_b200.policy import ENC, DecisionTransformer
from dt4image_restoration_b200.rollout import BatchedRollout


class ToyEngine:
    """Same surface as PnPEngine as far as BatchedRollout uses it; the 'step' is a cheap deterministic map."""

    def __init__(self, B, H, W):
        self.B, self.H, self.W = B, H, W
        self.device = torch.device("cpu")
        self.x = torch.zeros(B, 1, H, W)
        self.z = torch.zeros(B, 1, H, W)
        self.u = torch.zeros(B, 1, H, W)
        self.v = torch.zeros(B, 1, H, W)
        self.sigma = torch.zeros(B)
        self.mu = torch.zeros(B)

    def reset(self, data):
        self.x.copy_(data["x"])

    @staticmethod
    def advance(x, sigma, mu):
        return (0.8 * x + 0.5 * torch.roll(x, 1, -1) * sigma.reshape(-1, 1, 1, 1) + 0.1 * mu.reshape(-1, 1, 1, 1)).clamp(0, 1)

    def step(self, active=None):
        new = self.advance(self.x, self.sigma, self.mu)
        if active is not None:
            new = torch.where(active.reshape(-1, 1, 1, 1), new, self.x)
        self.x.copy_(new)

    def psnr(self):
        return self.x.mean(dim=(1, 2, 3))


def growing_window_loop(pol, x0, task, rtg0, K, Tmax, force):
    B = x0.shape[0]
    x = x0.clone()
    obs = torch.zeros(B, Tmax + 1, ENC * ENC); rtg = torch.zeros(B, Tmax + 1, 1); act = torch.zeros(B, Tmax + 1, 3)
    ts = torch.arange(Tmax + 1).reshape(1, -1, 1).expand(B, -1, -1) % pol.time_embed.num_embeddings
    rs = lambda img: (F.interpolate(img, size=(ENC, ENC), mode="area") if img.shape[-2:] != (ENC, ENC) else img).reshape(B, -1)
    rtg[:, 0] = rtg0
    obs[:, 0] = rs(x)
    active = torch.ones(B, dtype=torch.bool); executed = torch.zeros(B, dtype=torch.int32)
    for t in range(Tmax):
        lo = max(0, t - K + 1); sl = slice(lo, t + 1); tk = task.reshape(B, 1).expand(B, t + 1 - lo)
        pa, ad = pol(rtg[:, sl], obs[:, sl], ts[:, sl], tk, act[:, sl], eval_actions=True, hw=(ENC, ENC))
        act[:, t] = pa[:, -1]
        a = {k: ad[k][:, -1, 0] for k in ad}
        if not force:
            active = active & ~(a["T"] > 0.5)
        new = ToyEngine.advance(x, a["sigma_d"], a["mu"])
        x = torch.where(active.reshape(-1, 1, 1, 1), new, x)
        executed += active.to(torch.int32)
        rtg[:, t + 1] = pol(rtg[:, sl], obs[:, sl], ts[:, sl], tk, act[:, sl], eval_rtg=True, hw=(ENC, ENC))[:, -1]
        obs[:, t + 1] = rs(x)
    return x, act[:, :Tmax], rtg, executed


@pytest.mark.parametrize("force,K,Tmax", [(True, 6, 14), (False, 6, 14), (True, 3, 4), (True, 6, 3)])
def test_static_window_equals_growing_window(force, K, Tmax):
    B, H, W = 3, 64, 64
    torch.manual_seed(5)
    pol = DecisionTransformer()
    for m in pol.modules():                 # larger weights than the 0.02 init so that actions really depend on the history
        if isinstance(m, torch.nn.Linear):
            m.weight.data.mul_(8.0)
    if not force:
        pol.predict_action[0].bias.data[0] = -0.2
    x0 = torch.rand(B, 1, H, W)
    task = torch.tensor([1, 4, 7])
    eng = ToyEngine(B, H, W)
    ro = BatchedRollout(copy.deepcopy(pol), eng, context_length=K, max_timesteps=Tmax, force_full_length=force, use_graph=False)
    out = ro.run({"x": x0}, task, rtg0=0.62)
    x_ref, act_ref, rtg_ref, executed = growing_window_loop(pol, x0, task, 0.62, K, Tmax, force)
    assert out["executed"].tolist() == executed.tolist()
    assert (ro.act - act_ref).abs().max() < 1e-5
    assert (ro.rtg - rtg_ref).abs().max() < 1e-5
    assert (out["x"] - x_ref).abs().max() < 1e-5
    if not force:
        assert 0 < int(executed.sum()) < B * Tmax      # the early exit really happened for some trajectories
    # a second run on the same object starts from a clean window
    out2 = ro.run({"x": x0}, task, rtg0=0.62)
    assert (out2["x"] - x_ref).abs().max() < 1e-5

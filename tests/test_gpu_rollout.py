"""Batched policy-driven rollout and MCTS-style candidate expansion on the GPU vs a CPU loop
built from the oracle step and the same policy weights."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dt4image_restoration_b200 import synth
from dt4image_restoration_b200.engine import PnPEngine
from dt4image_restoration_b200.noise import UNetDenoiser2D
from dt4image_restoration_b200.policy import ENC, DecisionTransformer
from dt4image_restoration_b200.rollout import BatchedRollout, CandidateExpander
from oracle import pnp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def to_t(item):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in item.items()}


def cpu_rollout(pol, params, batch, task, rtg0, K, Tmax, force):
    """Same loop as BatchedRollout.run with the oracle step (CPU)."""
    st = O.reset(batch)
    B, _, H, W = st["gt"].shape
    st["x"] = st["x"].real.clone()
    x = st["x"]
    obs = torch.zeros(B, Tmax + 1, ENC * ENC); rtg = torch.zeros(B, Tmax + 1, 1); act = torch.zeros(B, Tmax + 1, 3)
    ts = torch.arange(Tmax + 1).reshape(1, -1, 1).expand(B, -1, -1)
    rtg[:, 0] = rtg0
    rs = lambda img: (F.interpolate(img, size=(ENC, ENC), mode="area") if img.shape[-2:] != (ENC, ENC) else img).reshape(B, -1)
    obs[:, 0] = rs(x)
    active = torch.ones(B, dtype=torch.bool); executed = torch.zeros(B, dtype=torch.int32)
    for t in range(Tmax):
        lo = max(0, t - K + 1); sl = slice(lo, t + 1); tk = task.reshape(B, 1).expand(B, t + 1 - lo)
        pa, ad = pol(rtg[:, sl], obs[:, sl], ts[:, sl], tk, act[:, sl], eval_actions=True, hw=(ENC, ENC))
        act[:, t] = pa[:, -1]
        a = {k: ad[k][:, -1, 0] for k in ad}
        if not force:
            active = active & ~(a["T"] > 0.5)
        old = {k: st[k].clone() for k in ("x", "z", "u")}
        for b in range(B):      # reference semantics: scalar mu, one image at a time
            if not active[b]:
                continue
            one = {k: (v[b:b + 1] if torch.is_tensor(v) and v.dim() >= 4 else v) for k, v in st.items()}
            one, _ = O.step(params, one, {"T": torch.zeros(1), "mu": a["mu"][b:b + 1], "sigma_d": a["sigma_d"][b:b + 1]})
            for k in ("x", "z", "u"):
                old[k][b:b + 1] = one[k]
        for k in ("x", "z", "u"):
            st[k] = old[k]
        executed += active.to(torch.int32)
        nxt = pol(rtg[:, sl], obs[:, sl], ts[:, sl], tk, act[:, sl], eval_rtg=True, hw=(ENC, ENC))
        rtg[:, t + 1] = nxt[:, -1]
        obs[:, t + 1] = rs(st["x"])
    return st, executed


@pytest.mark.parametrize("force", [True, False])
def test_rollout_matches_cpu_loop(force):
    B, H, W, Tmax = 3, 64, 64, 8
    params = O.init_unet_params(0, "default")
    batch = synth.make_batch(B, H, W, "cartesian", 4, 0.0, seed0=40)
    torch.manual_seed(7)
    pol = DecisionTransformer()
    if not force:   # bias the T head so that some trajectories stop early
        pol.predict_action[0].bias.data[0] = 0.3
    task = torch.tensor([3, 4, 5])
    eng = PnPEngine(UNetDenoiser2D(state_dict=params), B, H, W, DEV)
    import copy
    ro = BatchedRollout(copy.deepcopy(pol), eng, context_length=6, max_timesteps=Tmax, force_full_length=force)
    out = ro.run(to_t(batch), task, rtg0=0.62)
    st, executed = cpu_rollout(pol, params, batch, task, 0.62, 6, Tmax, force)
    assert out["executed"].cpu().tolist() == executed.tolist()
    assert (out["x"].cpu() - st["x"]).abs().max() < 1e-3
    assert out["image_iters"] == int(executed.sum())
    if force:
        assert executed.tolist() == [Tmax] * B


def test_candidate_expansion_matches_oracle():
    H = W = 128
    K = 8
    params = O.init_unet_params(0, "default")
    item = synth.make_item(synth.phantom(H, W, 9), synth.radial_mask(H, W, 0.2), 0.0, 9)
    st = O.reset(item)
    st, _ = O.step(params, st, {"T": torch.zeros(1), "mu": torch.tensor([0.5]), "sigma_d": torch.tensor([0.1])})
    g = torch.Generator().manual_seed(3)
    sig, mu = CandidateExpander.sample_actions(0.08, 0.4, K, g)
    eng = PnPEngine(UNetDenoiser2D(state_dict=params), K, H, W, DEV)
    dev_state = {k: st[k].to(DEV) for k in ("z", "u", "y0", "mask", "gt")}
    rew = CandidateExpander(eng).expand(dev_state, sig, mu).cpu()
    # same fan-out with the reward kernel fused with the (here: one-rank) peer-memory all-gather
    from dt4image_restoration_b200.dist import PeerRewardGather
    rew2 = CandidateExpander(eng).expand_and_gather(dev_state, sig, mu, K, PeerRewardGather(K, DEV, local_only=True)).cpu()
    assert torch.equal(rew, rew2)
    for k in range(K):
        one = {kk: (v.clone() if torch.is_tensor(v) else v) for kk, v in st.items()}
        one, _ = O.step(params, one, {"T": torch.zeros(1), "mu": mu[k:k + 1], "sigma_d": sig[k:k + 1]})
        ref = O.psnr(one["x"].reshape(1, H, W), one["gt"].reshape(1, H, W)).item()
        assert abs(rew[k].item() - ref) < 0.05


@pytest.mark.parametrize("pos", [0, 1, 3, 5])
def test_fused_policy_step_matches_the_two_pytorch_forwards(pos):
    """``pnp_policy_step`` (one kernel: action head at the newest observation token, then the return head at the new action
    token against cached keys / values) vs the two ``forward_tokens`` calls it replaces (reference eval.py:147-186)."""
    from dt4image_restoration_b200.policy import FusedPolicy
    torch.manual_seed(7)
    pol = DecisionTransformer().to(DEV).eval()
    with torch.no_grad():                                  # default init is N(0, 0.02): give the heads something to show
        for prm in pol.parameters():
            prm.mul_(3.0).add_(0.01 * torch.randn_like(prm))
    B, K = 5, 6
    g = torch.Generator(device=DEV).manual_seed(pos)
    w_rtg = torch.rand(B, K, 1, device=DEV, generator=g)
    w_emb = torch.randn(B, K, pol.embed_dim, device=DEV, generator=g) * 0.5
    w_act = torch.rand(B, K, 3, device=DEV, generator=g)
    w_ts = torch.randint(0, 30, (B, K, 1), device=DEV, generator=g)
    w_task = torch.randint(0, 9, (B, 1), device=DEV, generator=g).expand(B, K).contiguous()
    w_rtg[:, pos + 1:] = 0; w_emb[:, pos + 1:] = 0; w_act[:, pos:] = 0; w_ts[:, pos + 1:] = 0
    p = torch.tensor([pos], device=DEV)
    # reference: two forwards
    act_ref = w_act.clone()
    pa, ad = pol.forward_tokens(w_rtg, w_emb, w_ts, w_task, act_ref, eval_actions=True)
    pa_t = pa[:, pos]
    act_ref[:, pos] = pa_t
    rtg_ref = pol.forward_tokens(w_rtg, w_emb, w_ts, w_task, act_ref, eval_rtg=True)[:, pos]
    # fused
    fp = FusedPolicy(pol)
    act_io = w_act.clone()
    act_out = torch.zeros(B, 3, device=DEV); rtg_out = torch.zeros(B, 1, 1, device=DEV)
    fp.step(w_rtg, w_emb, act_io, w_ts, w_task, p, act_out, rtg_out)
    assert (act_out - pa_t).abs().max() < 2e-5
    assert torch.equal(act_io[:, pos], act_out) and torch.equal(act_io[:, :pos], w_act[:, :pos])
    assert (rtg_out.reshape(B) - rtg_ref.reshape(B)).abs().max() < 5e-5
    for i, k in enumerate(pol.action_keys):
        assert (act_out[:, i] - ad[k][:, pos, 0]).abs().max() < 2e-5


@pytest.mark.parametrize("S,pos", [(128, 0), (128, 5), (256, 2), (256, 5), (384, 5)])
def test_policy_observe_matches_encoder_and_window_update(S, pos):
    """``pnp_policy_observe`` (area mean + state encoder + context-window append in one kernel) vs the PyTorch encoder in fp32
    on the CPU (reference decision_transformer.py:128-132,215 after the area resize) and the rollout's index_select /
    index_copy window update."""
    import torch.nn.functional as F
    from dt4image_restoration_b200.policy import FusedPolicy
    torch.manual_seed(11)
    pol = DecisionTransformer().eval()
    with torch.no_grad():
        for prm in pol.state_encoder.parameters():
            prm.mul_(3.0).add_(0.01 * torch.randn_like(prm))
    B, K, d = 5, 6, pol.embed_dim
    g = torch.Generator().manual_seed(S + pos)
    x = torch.rand(B, 1, S, S, generator=g)
    w_rtg = torch.rand(B, K, 1, generator=g); w_emb = torch.randn(B, K, d, generator=g)
    w_act = torch.rand(B, K, 3, generator=g); w_ts = torch.randint(0, 30, (B, K, 1), generator=g)
    nxt = torch.rand(B, generator=g)
    t_dev = 17 + pos
    # reference (CPU fp32)
    with torch.no_grad():
        emb = pol.state_encoder(F.interpolate(x, size=(128, 128), mode="area") if S != 128 else x)
    r_rtg, r_emb, r_act, r_ts = w_rtg.clone(), w_emb.clone(), w_act.clone(), w_ts.clone()
    if pos == K - 1:
        for w in (r_rtg, r_emb, r_act, r_ts):
            w[:, :-1] = w[:, 1:].clone()
    npos = min(pos + 1, K - 1)
    r_rtg[:, npos, 0] = nxt; r_emb[:, npos] = emb; r_act[:, npos] = 0; r_ts[:, npos, 0] = (t_dev + 1) % pol.time_embed.num_embeddings
    # kernel
    fp = FusedPolicy(pol.to(DEV))
    assert fp.observe_supported(S, S)
    c = [t.to(DEV).contiguous() for t in (w_rtg, w_emb, w_act, w_ts)]
    fp.observe(x.to(DEV), nxt.to(DEV), c[0], c[1], c[2], c[3], torch.tensor([pos], device=DEV), torch.tensor([t_dev], device=DEV))
    assert torch.equal(c[0].cpu(), r_rtg) and torch.equal(c[2].cpu(), r_act) and torch.equal(c[3].cpu(), r_ts)
    assert (c[1].cpu()[:, npos] - r_emb[:, npos]).abs().max() < 2e-5
    keep = [i for i in range(K) if i != npos]
    assert torch.equal(c[1].cpu()[:, keep], r_emb[:, keep])

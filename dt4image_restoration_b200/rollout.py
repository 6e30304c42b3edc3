"""Batched policy <-> environment rollout: the batch-B, any-size, sync-free
counterpart of the reference's greedy loop (``Evaluator.run_greedy`` / ``predict_action_and_rtg``,
``evaluation/eval.py:147-220``) and of the candidate fan-out of its tree search (``expand_tree``,
``evaluation/mcts.py:103-143``).

Semantics are the standard decision-transformer rollout the reference intends: per step the policy sees the last
``K`` (return-to-go, observation, action) triples, the action head at the newest observation gives
``{T, sigma_d, mu}``, the return head at the newest action gives the next return-to-go; a trajectory whose
``T > 0.5`` stops (its state is left untouched, as reference ``env.py:79-81``) while the rest of the batch goes
on.  (The reference loop is batch-1, 128x128 only and has indexing quirks - ``eval.py:90-95,168-184`` - that are
not reproduced; the environment underneath is the parity-tested drop-in.)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .engine import PnPEngine
from .policy import ENC, DecisionTransformer


class BatchedRollout:
    def __init__(self, policy: DecisionTransformer, engine: PnPEngine, context_length: int = 6,
                 max_timesteps: int = 30, force_full_length: bool = False):
        self.policy, self.eng = policy.to(engine.device).eval(), engine
        self.K, self.Tmax = context_length, max_timesteps
        self.force = force_full_length     # hold T at 0: fixed-length trajectories (throughput runs)
        B, dev = engine.B, engine.device
        self.obs = torch.zeros(B, max_timesteps + 1, ENC * ENC, device=dev)
        self.rtg = torch.zeros(B, max_timesteps + 1, 1, device=dev)
        self.act = torch.zeros(B, max_timesteps + 1, policy.action_dim, device=dev)
        self.ts = torch.arange(max_timesteps + 1, device=dev).reshape(1, -1, 1).expand(B, -1, -1)

    def _observe(self, t: int):
        x = self.eng.x                                    # [B,1,H,W] fp32 (reference get_policy_ob, env.py:103-109)
        if x.shape[-2:] != (ENC, ENC):
            x = F.interpolate(x, size=(ENC, ENC), mode="area")
        self.obs[:, t] = x.reshape(self.eng.B, -1)

    @torch.no_grad()
    def run(self, data: dict, task: torch.Tensor, rtg0: float):
        eng, pol, K = self.eng, self.policy, self.K
        B, dev = eng.B, eng.device
        eng.reset(data)
        task = task.to(dev).reshape(B, 1)
        self.act.zero_(); self.rtg.zero_()
        self.rtg[:, 0] = rtg0
        self._observe(0)
        active = torch.ones(B, dtype=torch.bool, device=dev)
        executed = torch.zeros(B, dtype=torch.int32, device=dev)
        keys = list(pol.action_keys)
        for t in range(self.Tmax):
            lo = max(0, t - K + 1)
            sl = slice(lo, t + 1)
            tk = task.expand(B, t + 1 - lo)
            pa, ad = pol(self.rtg[:, sl], self.obs[:, sl], self.ts[:, sl] % pol.time_embed.num_embeddings, tk,
                         self.act[:, sl], eval_actions=True, hw=(ENC, ENC))
            self.act[:, t] = pa[:, -1]
            a = {k: ad[k][:, -1, 0] for k in keys}
            if not self.force:
                active = active & ~(a["T"] > 0.5)
            eng.sigma.copy_(a["sigma_d"]); eng.mu.copy_(a["mu"])
            eng.step(None if self.force else active)
            executed += active.to(torch.int32)
            nxt = pol(self.rtg[:, sl], self.obs[:, sl], self.ts[:, sl] % pol.time_embed.num_embeddings, tk,
                      self.act[:, sl], eval_rtg=True, hw=(ENC, ENC))
            self.rtg[:, t + 1] = nxt[:, -1]
            self._observe(t + 1)
        return {"x": eng.x, "psnr": eng.psnr().clone(), "executed": executed,
                "image_iters": int(executed.sum().item())}


class CandidateExpander:
    """MCTS-style fan-out: ONE shared state x ``K`` sampled ``(sigma_d, mu)`` actions -> ``K`` one-step children and
    their rewards (reference ``expand_tree``, mcts.py:111-126, does 1 + 5 sequential ``env.step`` calls on an aliased
    dict; here the K children are a batch, sharded over ranks, rewards all-gathered by ``dist.gather_rewards``)."""

    def __init__(self, engine: PnPEngine):
        self.eng = engine

    @staticmethod
    def sample_actions(sigma_d: float, mu: float, n: int, generator: torch.Generator | None = None):
        """|N(sigma_d, 0.2)| and |N(mu, 0.001)| as mcts.py:64-70,114-116 (host side, seeded)."""
        s = (torch.randn(n, generator=generator) * 0.2 + sigma_d).abs()
        m = (torch.randn(n, generator=generator) * 0.001 + mu).abs()
        return s, m

    @torch.no_grad()
    def expand(self, state: dict, sigma_d: torch.Tensor, mu: torch.Tensor):
        """``state``: device tensors ``x? z u y0 mask gt`` of ONE image ``[1,1,H,W]``; ``sigma_d, mu``: ``[B_local]``."""
        e = self.eng
        B = e.B
        e.z.copy_(state["z"].expand(B, -1, -1, -1)); e.u.copy_(state["u"].expand(B, -1, -1, -1))
        e.y0.copy_(state["y0"].expand(B, -1, -1, -1))
        e.mask.copy_(state["mask"].to(torch.uint8).expand(B, -1, -1, -1))
        e.gt.copy_(state["gt"].reshape(1, 1, e.H, e.W).expand(B, -1, -1, -1))
        e.v.copy_((e.z - e.u).real)
        e.prepare()
        e.set_actions(sigma_d, mu)
        e.step()
        return e.psnr()

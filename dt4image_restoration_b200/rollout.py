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
    """Every iteration runs on the same static ``K``-entry context window, so its body (policy action head ->
    environment step -> policy return head -> observation encoding -> window update) has the same shapes from the
    first iteration on and is captured ONCE in a CUDA graph (``use_graph=False``: same body, eager).

    While the window is still filling (t < K-1) the entries behind the newest one are zero padding: attention is
    causal, so the outputs at the newest position do not depend on them and equal those of the reference's growing
    window.  The position of the newest entry, the time step and the shift of a full window live in device tensors
    (``index_select`` / ``index_copy_``), nothing synchronises with the host."""

    def __init__(self, policy: DecisionTransformer, engine: PnPEngine, context_length: int = 6,
                 max_timesteps: int = 30, force_full_length: bool = False, use_graph: bool | None = None,
                 fused_policy: bool = True):
        self.policy, self.eng = policy.to(engine.device).eval(), engine
        # both policy heads of an iteration in ONE kernel (pnp_policy_step) instead of two PyTorch forwards
        self.fused = None
        if fused_policy and context_length <= 6:
            from .policy import FusedPolicy
            try:
                self.fused = FusedPolicy(self.policy)
            except Exception:
                self.fused = None
        self.K, self.Tmax = context_length, max_timesteps
        self.force = force_full_length     # hold T at 0: fixed-length trajectories (throughput runs)
        # One CUDA-graph replay per iteration pays when the host cannot keep up (small batches).  With the fused policy and
        # observation kernels an iteration is ~40 launches for >= 2.6 ms of GPU work at batch 64 / 256^2, and the eager
        # launches keep their programmatic-dependent-launch overlap: measured 19.3 k (eager) vs 19.0 k (graph) image-iters/s.
        if use_graph is None:
            use_graph = engine.B * getattr(engine, "H", 0) * getattr(engine, "W", 0) <= 8 * 256 * 256
        self.use_graph = use_graph
        B, dev, d, A = engine.B, engine.device, policy.embed_dim, policy.action_dim
        K = context_length
        # the context window (graph inputs / outputs); entry `pos` is the newest one
        self.w_rtg = torch.zeros(B, K, 1, device=dev)
        self.w_emb = torch.zeros(B, K, d, device=dev)
        self.w_act = torch.zeros(B, K, A, device=dev)
        self.w_ts = torch.zeros(B, K, 1, dtype=torch.int64, device=dev)
        self.w_task = torch.zeros(B, K, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int64, device=dev)
        self.t_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._ar = torch.arange(K, device=dev)
        self._zero_act = torch.zeros(B, 1, A, device=dev)
        self._out_act = torch.zeros(B, A, device=dev)
        self._out_rtg = torch.zeros(B, 1, 1, device=dev)
        # per-step record of the rollout (actions taken, return-to-go fed to the policy)
        self.act = torch.zeros(B, max_timesteps, A, device=dev)
        self.rtg = torch.zeros(B, max_timesteps + 1, 1, device=dev)
        self.active = torch.ones(B, dtype=torch.bool, device=dev)
        self.executed = torch.zeros(B, dtype=torch.int32, device=dev)
        self._graph = None
        self._captured = False

    def _encode_obs(self) -> torch.Tensor:
        x = self.eng.x                                    # [B,1,H,W] fp32 (reference get_policy_ob, env.py:103-109)
        if x.shape[-2:] != (ENC, ENC):
            x = F.interpolate(x, size=(ENC, ENC), mode="area")
        return self.policy.encode_states(x.reshape(self.eng.B, 1, ENC, ENC))[:, 0]

    def _state(self):
        return (self.w_rtg, self.w_emb, self.w_act, self.w_ts, self.pos, self.t_dev, self.active, self.executed,
                self._out_act, self.eng.x, self.eng.z, self.eng.u, self.eng.v)

    def _body(self):
        """One policy/environment iteration on the static window."""
        eng, pol, K, B = self.eng, self.policy, self.K, self.eng.B
        pos = self.pos
        if self.fused is not None:
            # action head at the newest observation AND return head at the new action token, one launch; the kernel also
            # writes the action into the window entry
            self.fused.step(self.w_rtg, self.w_emb, self.w_act, self.w_ts, self.w_task, pos, self._out_act, self._out_rtg)
            a = {k: self._out_act[:, i] for i, k in enumerate(pol.action_keys)}
        else:
            pa, ad = pol.forward_tokens(self.w_rtg, self.w_emb, self.w_ts, self.w_task, self.w_act, eval_actions=True)
            pa_t = pa.index_select(1, pos)                    # [B,1,A]: the action head at the newest observation
            self.w_act.index_copy_(1, pos, pa_t)
            self._out_act.copy_(pa_t[:, 0])
            a = {k: ad[k].index_select(1, pos)[:, 0, 0] for k in pol.action_keys}
        if not self.force:
            self.active &= ~(a["T"] > 0.5)
        eng.sigma.copy_(a["sigma_d"]); eng.mu.copy_(a["mu"])
        eng.step(None if self.force else self.active)
        self.executed += self.active.to(torch.int32)
        # the return head at the newest action [B,1,1].  (Running it on a side stream next to the environment step was
        # measured: no gain, the persistent conv kernels leave the tiny policy kernels no SM to run on.)
        if self.fused is not None:
            nxt = self._out_rtg
        else:
            nxt = pol.forward_tokens(self.w_rtg, self.w_emb, self.w_ts, self.w_task, self.w_act,
                                     eval_rtg=True).index_select(1, pos)
        if self.fused is not None and self.fused.observe_supported(eng.H, eng.W):
            # encoder + window update in one kernel (pnp_policy_observe); only the two scalars advance here
            self.fused.observe(eng.x, nxt.reshape(-1), self.w_rtg, self.w_emb, self.w_act, self.w_ts, pos, self.t_dev)
            self.t_dev += 1
            self.pos.clamp_(max=K - 2).add_(1)
            return
        emb = self._encode_obs()
        # window update: a full window moves one entry to the left, then the new (return-to-go, observation, empty
        # action, time step) entry goes behind the newest one
        full = (pos == K - 1).to(torch.int64)
        idx = torch.clamp(self._ar + full, max=K - 1)
        for w in (self.w_rtg, self.w_emb, self.w_act, self.w_ts):
            w.copy_(w.index_select(1, idx))
        npos = torch.clamp(pos + 1, max=K - 1)
        self.t_dev += 1
        self.w_rtg.index_copy_(1, npos, nxt)
        self.w_emb.index_copy_(1, npos, emb.unsqueeze(1))
        self.w_act.index_copy_(1, npos, self._zero_act)
        self.w_ts.index_copy_(1, npos, (self.t_dev % pol.time_embed.num_embeddings).reshape(1, 1, 1).expand(B, 1, 1))
        self.pos.copy_(npos)

    def _capture(self):
        self._captured = True
        if not self.use_graph:
            return
        saved = [t.clone() for t in self._state()]
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._body()                             # warm-up on a side stream (lazy initialisations)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self._graph = g
        except Exception:                                # capture not possible: stay eager
            self._graph = None
            self.use_graph = False
            torch.cuda.synchronize()
        for t, sv in zip(self._state(), saved):
            t.copy_(sv)

    @torch.no_grad()
    def run(self, data: dict, task: torch.Tensor, rtg0: float):
        eng, K = self.eng, self.K
        B, dev = eng.B, eng.device
        eng.reset(data)
        for w in (self.w_rtg, self.w_emb, self.w_act, self.w_ts, self.pos, self.t_dev, self.act, self.rtg, self.executed):
            w.zero_()
        self.w_task.copy_(task.to(dev).reshape(B, 1).expand(B, K))
        self.w_rtg[:, 0] = rtg0
        self.rtg[:, 0] = rtg0
        self.w_emb[:, 0] = self._encode_obs()
        self.active.fill_(True)
        if not self._captured:
            self._capture()
        for t in range(self.Tmax):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._body()
            self.act[:, t] = self._out_act
            self.rtg[:, t + 1] = self.w_rtg.index_select(1, self.pos)[:, 0]
        return {"x": eng.x, "psnr": eng.psnr().clone(), "executed": self.executed.clone(),
                "image_iters": int(self.executed.sum().item())}


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
        self._expand_no_reward(state, sigma_d, mu)
        return self.eng.psnr()

    def _expand_no_reward(self, state: dict, sigma_d: torch.Tensor, mu: torch.Tensor):
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

    @torch.no_grad()
    def expand_and_gather(self, state: dict, sigma_d: torch.Tensor, mu: torch.Tensor, n_units: int, peer=None):
        """``expand`` + the rewards of ALL ranks' candidates in global candidate order ``[n_units]``.  ``peer``: a
        ``dist.PeerRewardGather`` (reward kernel and all-gather fused over NVLink peer memory) or ``None`` (NCCL)."""
        from . import dist as pdist
        if peer is None:
            return pdist.gather_rewards(self.expand(state, sigma_d, mu), n_units)
        e = self.eng
        self._expand_no_reward(state, sigma_d, mu)
        allr = peer.psnr_allgather(e.x, e.gt, check=True)   # NaN rewards if a rank never arrived (10 s timeout)
        sizes = [pdist.shard_range(n_units, r, peer.world) for r in range(peer.world)]
        return torch.cat([allr[r, :hi - lo] for r, (lo, hi) in enumerate(sizes)])

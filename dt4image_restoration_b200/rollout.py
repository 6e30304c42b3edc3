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
    """``use_graph``: once the context window is full (t >= K-1) every iteration has the same shapes, so its body
    (policy action head -> environment step -> policy return head -> observation encoding -> window shift) is captured
    once in a CUDA graph and replayed; set False to run everything eagerly (same results)."""

    def __init__(self, policy: DecisionTransformer, engine: PnPEngine, context_length: int = 6,
                 max_timesteps: int = 30, force_full_length: bool = False, use_graph: bool = True):
        self.policy, self.eng = policy.to(engine.device).eval(), engine
        self.K, self.Tmax = context_length, max_timesteps
        self.force = force_full_length     # hold T at 0: fixed-length trajectories (throughput runs)
        self.use_graph = use_graph
        B, dev, d = engine.B, engine.device, policy.embed_dim
        self.emb = torch.zeros(B, max_timesteps + 1, d, device=dev)          # encoded observations, one per time step
        self.rtg = torch.zeros(B, max_timesteps + 1, 1, device=dev)
        self.act = torch.zeros(B, max_timesteps + 1, policy.action_dim, device=dev)
        self.ts = torch.arange(max_timesteps + 1, device=dev).reshape(1, -1, 1).expand(B, -1, -1)
        # static window of the steady state (graph inputs / outputs)
        K = context_length
        self.w_rtg = torch.zeros(B, K, 1, device=dev)
        self.w_emb = torch.zeros(B, K, d, device=dev)
        self.w_act = torch.zeros(B, K, policy.action_dim, device=dev)
        self.w_ts = torch.zeros(B, K, 1, dtype=torch.int64, device=dev)
        self.w_task = torch.zeros(B, K, dtype=torch.int64, device=dev)
        self.active = torch.ones(B, dtype=torch.bool, device=dev)
        self.executed = torch.zeros(B, dtype=torch.int32, device=dev)
        self._graph = None

    def _encode_obs(self) -> torch.Tensor:
        x = self.eng.x                                    # [B,1,H,W] fp32 (reference get_policy_ob, env.py:103-109)
        if x.shape[-2:] != (ENC, ENC):
            x = F.interpolate(x, size=(ENC, ENC), mode="area")
        return self.policy.encode_states(x.reshape(self.eng.B, 1, ENC, ENC))[:, 0]

    def _iteration(self, rtg, emb, ts, task, act):
        """One policy/environment iteration on a context window (views or static buffers); the newest entry is last.
        Writes the chosen action into ``act[:, -1]`` and returns (next return-to-go [B,1], next observation emb [B,d])."""
        eng, pol = self.eng, self.policy
        pa, ad = pol.forward_tokens(rtg, emb, ts, task, act, eval_actions=True)
        act[:, -1] = pa[:, -1]
        a = {k: ad[k][:, -1, 0] for k in pol.action_keys}
        if not self.force:
            self.active &= ~(a["T"] > 0.5)
        eng.sigma.copy_(a["sigma_d"]); eng.mu.copy_(a["mu"])
        eng.step(None if self.force else self.active)
        self.executed += self.active.to(torch.int32)
        nxt = pol.forward_tokens(rtg, emb, ts, task, act, eval_rtg=True)
        return nxt[:, -1], self._encode_obs()

    def _steady_body(self):
        nxt_rtg, nxt_emb = self._iteration(self.w_rtg, self.w_emb, self.w_ts, self.w_task, self.w_act)
        self._out_act.copy_(self.w_act[:, -1])
        # shift the window by one step and append the new (return-to-go, observation, empty action) triple
        for w in (self.w_rtg, self.w_emb, self.w_act, self.w_ts):
            w[:, :-1] = w[:, 1:].clone()
        self.w_rtg[:, -1] = nxt_rtg
        self.w_emb[:, -1] = nxt_emb
        self.w_act[:, -1] = 0
        self.w_ts[:, -1] = (self.w_ts[:, -2] + 1) % self.policy.time_embed.num_embeddings

    def _capture(self):
        self._out_act = torch.zeros_like(self.w_act[:, 0])
        if not self.use_graph:
            return
        saved = [t.clone() for t in (self.w_rtg, self.w_emb, self.w_act, self.w_ts, self.active, self.executed,
                                     self.eng.x, self.eng.z, self.eng.u, self.eng.v)]
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._steady_body()                      # warm-up on a side stream (lazy initialisations)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._steady_body()
            self._graph = g
        except Exception:                                # capture not possible: stay eager
            self._graph = None
            self.use_graph = False
            torch.cuda.synchronize()
        for t, sv in zip((self.w_rtg, self.w_emb, self.w_act, self.w_ts, self.active, self.executed,
                          self.eng.x, self.eng.z, self.eng.u, self.eng.v), saved):
            t.copy_(sv)

    @torch.no_grad()
    def run(self, data: dict, task: torch.Tensor, rtg0: float):
        eng, pol, K = self.eng, self.policy, self.K
        B, dev = eng.B, eng.device
        eng.reset(data)
        task = task.to(dev).reshape(B, 1)
        self.act.zero_(); self.rtg.zero_()
        self.rtg[:, 0] = rtg0
        self.emb[:, 0] = self._encode_obs()
        self.active.fill_(True); self.executed.zero_()
        nT = pol.time_embed.num_embeddings
        t = 0
        while t < self.Tmax and t < K - 1:               # growing window: eager
            sl = slice(0, t + 1)
            nxt_rtg, nxt_emb = self._iteration(self.rtg[:, sl], self.emb[:, sl], self.ts[:, sl] % nT,
                                               task.expand(B, t + 1), self.act[:, sl])
            self.rtg[:, t + 1] = nxt_rtg
            self.emb[:, t + 1] = nxt_emb
            t += 1
        if t < self.Tmax:                                 # full window: static buffers, one graph replay per iteration
            sl = slice(t - K + 1, t + 1)
            self.w_rtg.copy_(self.rtg[:, sl]); self.w_emb.copy_(self.emb[:, sl]); self.w_act.copy_(self.act[:, sl])
            self.w_ts.copy_(self.ts[:, sl] % nT); self.w_task.copy_(task.expand(B, K))
            if self._graph is None and not hasattr(self, "_out_act"):
                self._capture()
            while t < self.Tmax:
                if self._graph is not None:
                    self._graph.replay()
                else:
                    self._steady_body()
                self.act[:, t] = self._out_act
                self.rtg[:, t + 1] = self.w_rtg[:, -1]
                self.emb[:, t + 1] = self.w_emb[:, -1]
                t += 1
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

"""Tree search over PnP-ADMM programs on the GPU: the reference's ``run_mcts``
(``evaluation/mcts.py:212-258``) - tree of ``Node`` s (``:4-59``), p-UCB selection (``:74-88``), expansion by one policy
action + ``width`` sampled ``(sigma_d, mu)`` actions (``:103-143``, samples as ``:64-70``), evaluation of the expanded node
by a greedy policy rollout to the horizon (``run_beam_search`` ``:198-207`` -> ``Evaluator.run_greedy``,
``evaluation/eval.py:189-220``) and max-reward back-propagation (``:34-38``) - with the expansion done as ONE batched
environment step over all candidates, optionally sharded over the ranks of a process group.

Aliasing, explicitly FIXED.  The reference runs the policy step and the five child steps on the SAME state dict
(``mcts.py:118,126``: ``env.step`` re-binds ``x, z, u`` of the dict it is given), so its six steps chain and all children
share one state.  A batched expansion cannot chain; here every candidate is ONE step from the expanded node's own state,
which is what the tree is meant to hold.  ``oracle/ref_drivers.run_mcts(independent_children=True)`` is the same fix applied
to a line-by-line restatement of the reference loop (pinned bit-exact to the real ``run_mcts`` with the aliasing ON by
``oracle/make_golden_drivers.py``); ``tests/test_drivers.py`` checks this module against it: same programs, same best
program, rewards within 0.05 dB.

Kept from the reference on purpose (they define which programs get explored): the policy-context look-ups of
``Evaluator.predict_action_and_rtg`` (``eval.py:147-186``: window ``[time-K, time)`` once ``time >= K``, "latest" indices
``:39-60``), the horizon of 30, the cache keyed by ``(time, edge, iteration)``, children starting at reward 0, p-UCB without
the unused ``beta`` term, samples drawn from the global torch RNG in the reference's order (``torch.manual_seed`` to
reproduce).  The reward of a rollout is ``env.run_no_ref_reward`` (ARNIQA in the reference, ``env.py:42-54``; a hook
here, PSNR against ``gt`` by default).

Multi-GPU (one process per GPU): the tree and the RNG stream are replicated, the ``1 + width`` candidates of an expansion
are split over the ranks, each rank steps its share with the batched engine, the children's states stay on the rank that
computed them and move (one broadcast of ``x, z, u``) only when the search descends into them; the per-candidate rewards
are all-gathered (``dist.PeerRewardGather`` / NCCL) when ``child_prior='psnr'`` asks for them.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.distributions as tdist

from . import dist as pdist
from . import ops
from .engine import PnPEngine
from .env import PnPEnv
from .rollout import CandidateExpander


class TreeNode:
    __slots__ = ("parent", "children", "reward", "prob", "visits", "time", "edge", "index", "state", "owner", "rtg",
                 "policy_x", "action", "action_dict", "policy_emb")

    def __init__(self, rtg, state, time, prob, parent, edge, index, policy_x, owner=0, action_dict=None):
        self.parent, self.children = parent, []
        self.reward, self.prob, self.visits, self.time, self.edge, self.index = 0.0, prob, 0, time, edge, index
        self.state, self.owner = state, owner        # {'x','z','u'} device tensors [1,1,H,W] (None on ranks that do not hold it)
        self.rtg, self.policy_x, self.action, self.action_dict = rtg, policy_x, None, action_dict
        self.policy_emb = None                       # state-encoder output of policy_x, computed once (BatchedMCTS._node_emb)

    @property
    def key(self) -> str:                              # the reference's repr(node), its cache key (mcts.py:25-26)
        return f"Node(time = {self.time}, edge = {self.edge})_{self.index}"

    def backprop(self, reward: float):                 # mcts.py:34-38
        node = self
        while node is not None and reward > node.reward:
            node.reward = reward
            node = node.parent


def select_p_ucb(parent: TreeNode, children):
    """mcts.py:74-88: (child.reward - parent.reward) + prob * sqrt(log(parent visits)) / (1 + child visits); first maximum."""
    best, best_val = parent, -1000.0
    bonus = torch.sqrt(torch.log(torch.tensor([float(parent.visits)])))
    for c in children:
        val = (c.reward - parent.reward) + float(c.prob * bonus / (1 + c.visits))
        if val > best_val:
            best, best_val = c, val
    return best


def sample_actions(center: float, scale: float, n: int):
    """mcts.py:64-70: |N(center, scale)| samples sorted by decreasing density (global torch RNG, as the reference)."""
    d = tdist.Normal(center, scale)
    a = d.sample(torch.Size([n])).abs()
    p = torch.exp(d.log_prob(a))
    p, idx = torch.sort(p, descending=True)
    return a[idx], p


class _PolicyGraph:
    """``Evaluator.predict_action_and_rtg`` (eval.py:147-186) on a K-entry context window as ONE CUDA-graph replay: the
    action-head forward, the (conditional) write of the new action into the window, the return-head forward.  The eager
    version is ~140 tiny kernels issued from Python twice per call (3.2 of the 3.6 s of a search before this class); the
    observations enter already encoded (every observation is encoded once, not K times per call).

    Static inputs: the window (``rtg, emb, ts, task, act``), ``ka`` / ``kr`` = window positions whose action / return
    predictions are wanted, ``write`` = 1 when the new action belongs into the window at ``ka`` (``time < K``)."""

    def __init__(self, policy, K: int, dev):
        d, A = policy.embed_dim, policy.action_dim
        self.policy, self.K = policy, K
        self.rtg = torch.zeros(1, K, 1, device=dev)
        self.emb = torch.zeros(1, K, d, device=dev)
        self.ts = torch.zeros(1, K, 1, dtype=torch.int64, device=dev)
        self.task = torch.zeros(1, K, dtype=torch.int64, device=dev)
        self.act = torch.zeros(1, K, A, device=dev)
        self.ka = torch.zeros(1, dtype=torch.int64, device=dev)
        self.kr = torch.zeros(1, dtype=torch.int64, device=dev)
        self.write = torch.zeros(1, 1, 1, device=dev)
        self.out_pa = torch.zeros(A, device=dev)
        self.out_pr = torch.zeros(1, device=dev)
        self.graph = None

    def _body(self):
        pol = self.policy
        pa, _ = pol.forward_tokens(self.rtg, self.emb, self.ts, self.task, self.act, eval_actions=True)
        pa_sel = pa.index_select(1, self.ka)                                   # [1,1,A]
        act2 = self.act.clone()
        act2.index_copy_(1, self.ka, torch.where(self.write > 0, pa_sel, self.act.index_select(1, self.ka)))
        pr = pol.forward_tokens(self.rtg, self.emb, self.ts, self.task, act2, eval_rtg=True)
        self.out_pa.copy_(pa_sel.reshape(-1))
        self.out_pr.copy_(pr.index_select(1, self.kr).reshape(-1))

    def run(self, rtg, emb, ts, task, act, ka: int, kr: int, write: bool):
        self.rtg.copy_(rtg); self.emb.copy_(emb); self.ts.copy_(ts); self.task.copy_(task); self.act.copy_(act)
        self.ka.fill_(ka); self.kr.fill_(kr); self.write.fill_(1.0 if write else 0.0)
        if self.graph is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._body()                                                   # warm-up outside the capture
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
        self.graph.replay()
        return self.out_pa.clone(), self.out_pr.clone()


class BatchedMCTS:
    def __init__(self, policy, denoiser, H: int, W: int, width: int = 5, n_iters: int = 30, max_timesteps: int = 30,
                 context_length: int = 6, device="cuda", reward_fn=None, rank: int = 0, world: int = 1, peer=None,
                 child_prior: str = "zero", graph_policy: bool = True):
        self.H, self.W, self.width, self.n_iters = H, W, width, n_iters
        self.Tmax, self.K, self.dev = max_timesteps, context_length, torch.device(device)
        self.policy = policy.to(self.dev).eval()
        self.env = PnPEnv(max_timesteps, denoiser, self.dev)
        self.env.no_ref_model = reward_fn if reward_fn is not None else self._psnr_reward
        self.rank, self.world, self.peer, self.child_prior = rank, world, peer, child_prior
        lo, hi = pdist.shard_range(1 + width, rank, world)
        self.lo, self.hi = lo, hi
        self.expander = CandidateExpander(PnPEngine(denoiser, max(hi - lo, 1), H, W, self.dev)) if hi > lo else None
        self.env_steps = 0
        # graph_policy: policy calls replay one CUDA graph on encoded observations (_PolicyGraph); False = the eager forwards
        self._pg = _PolicyGraph(self.policy, self.K, self.dev) if graph_policy else None

    # ------------------------------------------------------------------------------------------
    def _psnr_reward(self, state) -> float:
        x = state["x"]
        x = x.real if x.is_complex() else x
        return float(ops.psnr(x.reshape(1, -1), state["gt"].reshape(1, -1))[0])

    def _buffers(self, task):
        """tasks, timesteps, actions, observations, returns per time index.  With the graphed policy the observation
        buffer holds ENCODED observations ``[1, T, d]`` (see ``_ob``)."""
        T, dev = self.Tmax, self.dev
        n_ob = self.policy.embed_dim if self._pg is not None else self.H * self.W
        return (task.to(dev).reshape(1, -1)[:, :1].repeat(1, T), torch.arange(T, device=dev).reshape(1, T, 1),
                torch.zeros(1, T, 3, device=dev), torch.zeros(1, T, n_ob, device=dev), torch.zeros(1, T, 1, device=dev))

    def _ob(self, x):
        """What goes into the observation buffer for the image ``x``: the flat image (eager policy) or its state-encoder
        output (graphed policy; ``DecisionTransformer.encode_states`` of this ONE image)."""
        x = x.real if x.is_complex() else x
        flat = x.reshape(1, 1, -1)
        if self._pg is None:
            return flat[:, 0]
        return self.policy.encode_states(flat, (self.H, self.W))[:, 0]

    def _node_ob(self, n: TreeNode):
        if self._pg is None:
            return n.policy_x.reshape(1, -1)
        if n.policy_emb is None:
            n.policy_emb = self._ob(n.policy_x)
        return n.policy_emb

    def _fill_history(self, node: TreeNode, obs, rtgs, acts):
        """``Node.build_eval`` / ``build_action`` (mcts.py:40-58): observations and returns of the path, actions of the
        ancestors."""
        n = node
        while True:
            t = n.time if n.time >= 1 else 0
            obs[:, t] = self._node_ob(n)
            rtgs[:, t] = n.rtg
            if n.time < 1:
                break
            n = n.parent
        n = node.parent
        while n is not None:
            t = n.time if n.time >= 1 else 0
            acts[:, t] = n.action
            if n.time < 1:
                break
            n = n.parent

    @torch.no_grad()
    def _predict(self, obs, acts, rtgs, ts, tasks, time: int):
        """``Evaluator.predict_action_and_rtg`` (eval.py:147-186) with its look-up rules (:39-60)."""
        K = self.K
        sl = slice(0, K) if time < K else slice(time - K, time)
        hw = (self.H, self.W)
        if self._pg is not None:
            # same look-ups as below: action at window entry `time` (or the last one), written into the window only when
            # `time` lies inside it; return at entry (k - 1) with k = time + 1 (or -1)
            ka = time if time < K else K - 1
            kr = time if time + 1 <= K else K - 2
            pa, pr = self._pg.run(rtgs[:, sl], obs[:, sl], ts[:, sl], tasks[:, sl], acts[:, sl], ka, kr, time < K)
            acts[:, time] = pa
            ad = OrderedDict((key, pa[i:i + 1]) for i, key in enumerate(self.policy.action_keys))
            return pa, ad, pr
        pa, ad = self.policy(rtgs[:, sl], obs[:, sl], ts[:, sl], tasks[:, sl], acts[:, sl], eval_actions=True, hw=hw)
        k = -1 if time >= K else time
        ad = OrderedDict((key, ad[key][0][k]) for key in ad)
        pa = pa[0][k]
        acts[:, time] = pa
        pr = self.policy(rtgs[:, sl], obs[:, sl], ts[:, sl], tasks[:, sl], acts[:, sl], eval_rtg=True, hw=hw)
        k = -1 if time + 1 > K else time + 1
        return pa, ad, pr[0][k - 1]

    def _env_state(self, node: TreeNode, consts) -> OrderedDict:
        st = OrderedDict(consts)
        st.update(node.state)
        st["T"] = node.time / 30
        return st

    # ------------------------------------------------------------------------------------------
    def _fetch_state(self, node: TreeNode):
        """Make ``node.state`` available on every rank (one broadcast of x, z, u from the rank that computed it)."""
        if self.world == 1 or node.owner < 0:
            return
        import torch.distributed as tdd
        shapes = {"x": torch.float32, "z": torch.complex64, "u": torch.complex64}
        if node.state is None:
            node.state = {k: torch.empty(1, 1, self.H, self.W, dtype=dt, device=self.dev) for k, dt in shapes.items()}
        for k in ("x", "z", "u"):
            buf = torch.view_as_real(node.state[k]) if node.state[k].is_complex() else node.state[k]
            tdd.broadcast(buf, src=node.owner)
        node.owner = -1                                  # now replicated

    @torch.no_grad()
    def _expand(self, node: TreeNode, consts, task, index: int):
        tasks, ts, acts, obs, rtgs = self._buffers(task)
        self._fill_history(node, obs, rtgs, acts)
        pa, ad, pr = self._predict(obs, acts, rtgs, ts, tasks, node.time)
        node.action = pa
        sig, _ = sample_actions(float(ad["sigma_d"]), 0.2, self.width)
        mu, probs = sample_actions(float(ad["mu"]), 0.001, self.width)
        # candidate 0 = the policy's own action (the reference's `policy_state`), 1.. = the samples
        all_sig = torch.cat([ad["sigma_d"].reshape(1).cpu().float(), sig.float()])
        all_mu = torch.cat([ad["mu"].reshape(1).cpu().float(), mu.float()])
        stop = bool(ad["T"] > 0.5)                        # env.py:79-81: the step returns the state untouched
        lo, hi = self.lo, self.hi
        prior = None
        if stop or self.expander is None:
            local = {k: node.state[k].expand(max(hi - lo, 1), -1, -1, -1) for k in ("x", "z", "u")} if hi > lo else None
        else:
            st = dict(node.state, y0=consts["y0"], mask=consts["mask"], gt=consts["gt"])
            e = self.expander.eng
            if self.child_prior == "psnr":
                prior = self.expander.expand_and_gather(st, all_sig[lo:hi].to(self.dev), all_mu[lo:hi].to(self.dev),
                                                        1 + self.width, self.peer) if self.world > 1 else \
                    self.expander.expand(st, all_sig[lo:hi].to(self.dev), all_mu[lo:hi].to(self.dev)).clone()
            else:
                self.expander._expand_no_reward(st, all_sig[lo:hi].to(self.dev), all_mu[lo:hi].to(self.dev))
            self.env_steps += hi - lo
            local = {"x": e.x, "z": e.z, "u": e.u}

        def cand_state(c):                                # fresh tensors: engine buffers are reused by the next expansion
            if not (lo <= c < hi):
                return None
            return {k: local[k][c - lo:c - lo + 1].clone() for k in ("x", "z", "u")}

        owner_of = lambda c: next(r for r in range(self.world) if pdist.shard_range(1 + self.width, r, self.world)[0] <= c
                                  < pdist.shard_range(1 + self.width, r, self.world)[1])
        # the policy child's x is every child's observation for the policy (mcts.py:135: policy_state)
        pol = cand_state(0)
        if self.world > 1:
            import torch.distributed as tdd
            px = pol["x"] if pol is not None else torch.empty(1, 1, self.H, self.W, device=self.dev)
            tdd.broadcast(px, src=owner_of(0))
            policy_x = px
        else:
            policy_x = pol["x"]
        ad_children = ad
        for i in range(self.width):
            c = 1 + i
            child = TreeNode(pr, cand_state(c), node.time + 1, probs[i], node, i, index, policy_x,
                             owner=(owner_of(c) if self.world > 1 else -1),
                             action_dict=OrderedDict(T=ad_children["T"], sigma_d=sig[i], mu=mu[i]))
            if prior is not None:
                child.reward = float(prior[c])
            node.children.append(child)
        return node

    @torch.no_grad()
    def _rollout(self, node: TreeNode, consts, task):
        """``run_beam_search`` + ``run_greedy`` (mcts.py:198-207, eval.py:189-220) from ``node`` to the horizon."""
        tasks, ts, acts, obs, rtgs = self._buffers(task)
        self._fill_history(node, obs, rtgs, acts)
        _, ad, _ = self._predict(obs, acts, rtgs, ts, tasks, node.time)
        st = self._env_state(node, consts)
        pred_rtg = node.rtg
        for time in range(node.time, self.Tmax + 1):
            st, done = self.env.step(st, ad)
            self.env_steps += 0 if done else 1
            if time == self.Tmax or done:
                return self.env.run_no_ref_reward(st), time, st["x"].real if st["x"].is_complex() else st["x"]
            obs[:, time] = self._ob(st["x"])
            rtgs[:, time] = pred_rtg
            _, ad, pred_rtg = self._predict(obs, acts, rtgs, ts, tasks, time)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def search(self, item: dict, rtg0, task):
        """``item``: the reference's eval item (``x0, y0, mask, ATy0, gt``; tensors or arrays), ``rtg0`` the normalised
        target return, ``task`` the task token.  Returns the best program's final PSNR ``[1,1]`` (CPU), its key and the
        cached program rewards."""
        st0 = self.env.reset({k: torch.as_tensor(v) for k, v in item.items()}, self.dev)
        consts = OrderedDict((k, st0[k]) for k in ("y0", "mask", "gt", "ATy0", "complex_y0"))
        x0 = st0["x"].real.contiguous() if st0["x"].is_complex() else st0["x"]
        rtg0 = torch.as_tensor(rtg0, dtype=torch.float32, device=self.dev).reshape(1, 1, 1)
        task = torch.as_tensor(task).reshape(1, -1)
        root = TreeNode(rtg0, {"x": st0["x"], "z": st0["z"], "u": st0["u"]}, 0, 1.0, None, 0, 0, x0, owner=-1)
        programs, finals, nodes = OrderedDict(), {}, {}
        root.visits += 1
        for i in range(self.n_iters):
            node = root
            node.visits += 1
            while node.children:
                node = select_p_ucb(node, node.children)
                node.visits += 1
            self._fetch_state(node)
            self._expand(node, consts, task, i)
            key = node.key
            if key not in programs:
                reward, _, final = self._rollout(node, consts, task)
                node.reward = reward
                programs[key], finals[key], nodes[key] = reward, final, node
            node.backprop(programs[key])
        best_key = None
        best = -1000.0
        for k, r in programs.items():
            if r > best:
                best, best_key = r, k
        final_x = finals[best_key]
        return PnPEnv.compute_reward(consts["gt"].reshape(1, self.H, self.W), final_x.reshape(1, self.H, self.W)), best_key, programs

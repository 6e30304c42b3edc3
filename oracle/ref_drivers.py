"""Restatement of the reference's DRIVERS of the environment (the callers of ``PnPEnv``).  TEST INFRASTRUCTURE ONLY.

The GPU box has no ``/root/reference``, so the reference's own loops cannot be imported there.  This module restates
them - generic over the environment object and the policy model, device-agnostic - so that the drop-in ``PnPEnv`` can
be driven exactly the way the reference drives its own environment:

* ``GreedyDriver``            <- ``Evaluator`` (``evaluation/eval.py``): ``_get_latest_action`` ``:39-50``,
                                 ``_get_latest_rtg`` ``:53-60``, ``get_initial_policy_setup`` ``:62-100``,
                                 ``predict_action_and_rtg`` ``:147-186``, ``run_greedy`` ``:189-220``
* ``Node`` / ``sample_action_dict`` / ``select_p_ucb`` / ``expand_tree`` / ``run_beam_search`` / ``run_mcts`` /
  ``get_best_program``        <- ``evaluation/mcts.py:4-59, 64-70, 74-88, 103-143, 198-207, 212-258, 165-192``

Quirks are kept on purpose because they define "drop-in": the index arithmetic of the latest-action / latest-return
look-ups, the broadcast single-position calls of the initial set-up (``eval.py:90-95``), the ONE state dict and ONE
action dict that all six ``env.step`` calls of an expansion share (``mcts.py:118-136``), the cache keyed by
``repr(node)``.  ``run_mcts(..., independent_children=True)`` is the explicit FIX of the aliasing (every child is one
step from the parent's own state), which is what a batched expansion computes.

Pin: ``oracle/make_golden_drivers.py`` runs the REAL reference drivers (imported from ``/root/reference``) and these on
the same seeded inputs in the build container, asserts equal decisions and rewards, and writes
``tests/golden/ref_drivers.npz``; ``tests/test_drivers.py`` re-checks the restatement on the CPU and drives the CUDA
drop-in with it on the GPU box.  The 128 x 128 / batch-1 shapes are the reference's own (``eval.py:66,207``).
"""
from __future__ import annotations

import copy
from collections import OrderedDict

import torch
import torch.distributions as dist


class GreedyDriver:
    """``Evaluator`` without the checkpoint / dataset plumbing: model, env, context length, horizon."""

    def __init__(self, model, env, device, context_length: int = 6, max_timesteps: int = 30, action_dim: int = 3,
                 side: int = 128):
        self.model, self.env, self.device = model, env, device
        self.context_length, self.max_timesteps, self.action_dim, self.side = context_length, max_timesteps, action_dim, side

    # eval.py:39-50 - NOTE `>=` here and `>` in latest_rtg, as in the reference
    def latest_action(self, action_dict, preds, index):
        k = -1 if index >= self.context_length else index
        preds = preds[0][k]
        for key in ("T", "mu", "sigma_d"):
            action_dict[key] = action_dict[key][0][k]
        return action_dict, preds

    # eval.py:53-60
    def latest_rtg(self, rtg_preds, index):
        k = -1 if index > self.context_length else index
        return rtg_preds[0][k - 1]

    # eval.py:62-100
    def initial_setup(self, policy_inputs, mat):
        states, rtg, _, task = policy_inputs
        dev, T, K = self.device, self.max_timesteps, self.context_length
        states, rtg = states.to(dev), rtg.to(dev)
        acts = torch.zeros((1, T, self.action_dim), device=dev)
        obs = torch.zeros((1, T, self.side * self.side), device=dev)
        rtgs = torch.zeros((1, T, 1), device=dev)
        ts = torch.arange(0, T).reshape(1, T, 1).contiguous().to(dev)
        tasks = task.repeat(1, T).to(dev)
        obs[0, 0] = states
        rtgs[0, 0] = rtg
        env_state = self.env.reset(mat, dev)
        with torch.no_grad():
            self.model.eval()
            pa, ad = self.model(rtgs[:, :K], obs[:, :K], ts[:, :K], tasks[:, :K], actions=None)
        ad, pa = self.latest_action(ad, pa, index=0)
        acts[:, 0] = pa
        with torch.no_grad():      # eval.py:90-95: single POSITION K (not a slice) for rtg and actions, broadcast by the model
            pr = self.model(rtgs[:, K], obs[:, :K], ts[:, :K], tasks[:, :K], acts[:, K], eval_rtg=True)
        pr = self.latest_rtg(pr, index=1)
        return (obs, acts, rtgs, ts, tasks), (env_state, pr, pa, ad)

    # eval.py:147-186
    @torch.no_grad()
    def predict_action_and_rtg(self, obs, acts, rtgs, ts, tasks, time):
        self.model.eval()
        K = self.context_length
        sl = slice(0, K) if time < K else slice(time - K, time)
        pa, ad = self.model(rtgs[:, sl], obs[:, sl], ts[:, sl], tasks[:, sl], acts[:, sl], eval_actions=True)
        ad, pa = self.latest_action(ad, pa, index=time)
        acts[:, time] = pa
        pr = self.model(rtgs[:, sl], obs[:, sl], ts[:, sl], tasks[:, sl], acts[:, sl], eval_rtg=True)
        pr = self.latest_rtg(pr, index=time + 1)
        return pa, ad, pr

    # eval.py:189-220
    def run_greedy(self, env_state, pred_rtg, start_time, action_dict, obs, acts, rtgs, ts, tasks, no_ref=False, log=None):
        for time in range(start_time, self.max_timesteps + 1):
            if log is not None:
                log.append([float(action_dict[k]) for k in ("T", "sigma_d", "mu")])
            env_state, done = self.env.step(env_state, action_dict)
            ob = self.env.get_policy_ob(env_state)
            if time == self.max_timesteps or done:
                x = env_state["x"].reshape(1, self.side, self.side)
                reward = self.env.run_no_ref_reward(env_state) if no_ref else self.env.compute_reward(x, env_state["gt"])
                return reward, time, x
            obs[:, time] = ob
            rtgs[:, time] = pred_rtg
            _, action_dict, pred_rtg = self.predict_action_and_rtg(obs, acts, rtgs, ts, tasks, time)


# ------------------------------------------------------------------------------------------------------------------
# tree search (evaluation/mcts.py)
# ------------------------------------------------------------------------------------------------------------------
class Node:
    max_timesteps = 30

    def __init__(self, rtg, state, time, prob, parent, edge, action_dict, index, policy_state, task):
        self.parent, self.children = parent, []
        self.reward, self.prob, self.s_visits, self.time = 0, prob, 0, time
        self.state = state["x"].real.reshape(1, -1)                  # snapshot of x (mcts.py:15)
        self.edge, self.env_state, self.action_dict, self.index = edge, state, action_dict, index
        self.policy_rtg, self.policy_state, self.task = rtg, policy_state, task
        self.action = None

    def __repr__(self):                                              # the cache key (mcts.py:25-26)
        return f"Node(time = {self.time}, edge = {self.edge})_{self.index}"

    def backprop(self, reward):                                      # mcts.py:34-38: max, stops where no improvement
        if reward > self.reward:
            self.reward = reward
            if self.parent is not None:
                self.parent.backprop(reward)

    def build_eval(self, obs, rtgs):                                 # mcts.py:40-50
        node = self
        while True:
            t = node.time if node.time >= 1 else 0
            obs[:, t] = node.policy_state["x"].real.reshape(1, -1)
            rtgs[:, t] = node.policy_rtg
            if node.time < 1:
                return obs, rtgs
            node = node.parent

    def build_action(self, acts):                                    # mcts.py:52-58
        node = self
        while True:
            t = node.time if node.time >= 1 else 0
            acts[:, t] = node.action
            if node.time < 1:
                return acts
            node = node.parent


def sample_action_dict(action, prob):                                # mcts.py:64-70 (global CPU RNG)
    d = dist.Normal(action.item(), prob)
    a = d.sample(torch.Size([5])).abs()
    p = torch.exp(d.log_prob(a))
    p, idx = torch.sort(p, descending=True)
    return a[idx], p


def select_p_ucb(parent, children):                                  # mcts.py:74-88 (beta is computed and unused there)
    best, best_val = parent, -1000
    for node in children:
        val = (node.reward - parent.reward) + node.prob * torch.sqrt(torch.log(torch.Tensor([parent.s_visits]))) / (1 + node.s_visits)
        node.p_ucb = val
        if val > best_val:
            best, best_val = node, val
    return best


def prepare_evaluation(node, task, device="cpu", side=128):          # mcts.py:93-99
    T = node.max_timesteps
    return (task.repeat(1, T).to(device), torch.arange(0, T).reshape(1, T, 1).contiguous().to(device),
            torch.zeros((1, T, 3), device=device), torch.zeros((1, T, side * side), device=device),
            torch.zeros((1, T, 1), device=device))


def _dup_state(state):
    return OrderedDict((k, v) for k, v in state.items())             # new dict, same tensors (step re-binds, never writes)


def expand_tree(driver, node, task, env, node_list, index_tree, device="cpu", independent_children=False, log=None):
    """mcts.py:103-143.  ``independent_children=False`` keeps the reference's aliasing: the policy step and the five
    child steps all run on ``node.env_state`` itself (one dict, re-bound in place), so they CHAIN, and every child holds
    the same dict and the same (mutated) action dict."""
    tasks, ts, acts, obs, rtgs = prepare_evaluation(node, task, device, driver.side)
    obs, rtgs = node.build_eval(obs, rtgs)
    if node.parent:
        acts = node.parent.build_action(acts)
    pa, ad, pr = driver.predict_action_and_rtg(obs, acts, rtgs, ts, tasks, node.time)
    node.action = pa
    sigma_d, probs = sample_action_dict(ad["sigma_d"], 0.2)
    mu, probs = sample_action_dict(ad["mu"], 0.001)
    if log is not None:
        log.append([float(ad["T"]), float(ad["sigma_d"]), float(ad["mu"])] + [float(v) for v in sigma_d] + [float(v) for v in mu])
    parent_state = _dup_state(node.env_state) if independent_children else None
    policy_state, _ = env.step(_dup_state(parent_state) if independent_children else node.env_state, ad)
    for i in range(len(mu)):
        if independent_children:
            ad_i = OrderedDict(ad)
            ad_i["sigma_d"], ad_i["mu"] = sigma_d[i], mu[i]
            child_state, _ = env.step(_dup_state(parent_state), ad_i)
        else:
            ad_i = ad
            ad_i["sigma_d"], ad_i["mu"] = sigma_d[i], mu[i]
            child_state, _ = env.step(node.env_state, ad_i)
        child = Node(pr, child_state, node.time + 1, probs[i], node, i, ad_i, index_tree, policy_state, tasks)
        node.children.append(child)
        node_list.append(child)
    return node


def run_beam_search(node, driver, device="cpu"):                     # mcts.py:198-207
    tasks, ts, acts, obs, rtgs = prepare_evaluation(node, node.task, device, driver.side)
    obs, rtgs = node.build_eval(obs, rtgs)
    if node.parent:
        acts = node.parent.build_action(acts)
    _, ad, _ = driver.predict_action_and_rtg(obs, acts, rtgs, ts, tasks, node.time)
    reward, time, final = driver.run_greedy(node.env_state, node.policy_rtg, node.time, ad, obs, acts, rtgs, ts, tasks, True)
    return reward, final, time


def get_best_program(program_dict, state_dict, node_list, env, side=128):   # mcts.py:165-192
    best_key, best = None, -1000
    for k, r in program_dict.items():
        if r > best:
            best, best_key = r, k
    node = node_list[-1]
    for n in node_list:
        if repr(n) == best_key:
            node = n
            break
    final = state_dict[repr(node)]
    while node.parent:
        node = node.parent
    return env.compute_reward(node.env_state["gt"].reshape(1, side, side), final), best_key


def run_mcts(driver, policy_inputs, mat, task, env, device, n_iters=30, independent_children=False, log=None):
    """mcts.py:212-258.  Returns (final reward tensor, best cache key, dict of cached rewards)."""
    node_list = []
    _, rtg, _, task = policy_inputs
    states = env.reset(mat, device)
    rtg = rtg.to(device)
    # the reference's root observes its OWN state dict (mcts.py:218), which the first rollout then mutates; the fix un-aliases it
    root = Node(rtg, states, 0, 1, None, 0, None, 0, _dup_state(states) if independent_children else states, task)
    programs, finals = OrderedDict(), {}
    node_list.append(root)
    root.s_visits += 1
    for i in range(n_iters):
        node = root
        node.s_visits += 1
        while len(node.children) > 0:
            node = select_p_ucb(node, node.children)
            node.s_visits += 1
        node = expand_tree(driver, node, task, env, node_list, i, device, independent_children, log)
        key = repr(node)
        reward = programs.get(key, -100)
        if reward == -100:
            reward, final, _ = run_beam_search(node, driver, device)
            node.reward = reward
            programs[key] = reward
            finals[key] = final
        node.backprop(reward)
    reward, best_key = get_best_program(programs, finals, node_list, env, driver.side)
    return reward, best_key, programs

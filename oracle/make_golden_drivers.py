"""Pin ``oracle/ref_drivers.py`` against the REAL reference drivers and write ``tests/golden/ref_drivers.npz``.
TEST INFRASTRUCTURE ONLY.  Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden_drivers

What runs (CPU, fp32, 128 x 128, batch 1 - the reference's own shapes):
  * the reference's ``Evaluator.get_initial_policy_setup`` + ``run_greedy`` (``evaluation/eval.py:62-100,189-220``) and
    ``run_mcts`` (``evaluation/mcts.py:212-258``) on the reference's ``PnPEnv`` / ``UNetDenoiser2D`` / ``DecisionTransformer``;
  * the restated drivers on the same objects - asserted EQUAL (rewards, stop times, every action, the cached programs);
  * the restated tree search with ``independent_children=True`` (the aliasing fix) on the oracle environment: the CPU
    reference of the batched search in ``dt4image_restoration_b200/mcts.py``.
ARNIQA (the reference's no-reference reward, ``env.py:36-54``) needs the network; PSNR against ``gt`` stands in for it on
both sides.  The policy is the seeded random-init decision transformer with the stop head biased to "continue"
(``predict_action`` bias for T = -2: sigmoid -> 0.12), otherwise a random-init policy stops at a random step.
"""
from __future__ import annotations

import os
import sys
import tempfile
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dt4image_restoration_b200 import synth  # noqa: E402
from oracle import pnp_oracle as O  # noqa: E402
from oracle import ref_drivers as RD  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DT_SEED, UNET_SEED, ITEM_SEED, MCTS_SEED, T_BIAS = 1234, 0, 2, 99, -2.0


def to_t(item):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in item.items()}


def policy_inputs(item):
    """What ``EvaluationOptimalDataset.__getitem__`` + ``DataLoader(batch_size=1)`` hand over (datasets.py:184-207)."""
    x = torch.from_numpy(np.ascontiguousarray(item["x0"][..., 0])).float().reshape(1, 1, -1)
    rtg = torch.tensor([(10.0 + 1.08) / (16.6 + 1.08)]).reshape(1, 1, 1)
    return x, rtg, torch.zeros(1, 3), torch.tensor([[3]])


def bias_stop_head(model):
    with torch.no_grad():
        model.predict_action[0].bias[0] = T_BIAS          # head order in 'norm' mode: T, sigma_d, mu


class OracleEnv:
    """The oracle environment behind the reference's env interface (reset/step/get_policy_ob/compute_reward)."""

    def __init__(self, params):
        self.params = params

    def reset(self, data, device):
        return O.reset({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in data.items()})

    def step(self, states, action_dict):
        return O.step(self.params, states, action_dict)

    @staticmethod
    def get_policy_ob(state):
        return O.policy_ob(state)

    @staticmethod
    def compute_reward(x, gt):
        return O.psnr(x, gt.reshape(-1, *x.shape[1:]) if gt.numel() != x[0].numel() else gt.reshape(1, *x.shape[1:]))

    def run_no_ref_reward(self, state):
        return psnr_stand_in(self, state)


def psnr_stand_in(env, state):
    x = state["x"]
    x = x.real if x.is_complex() else x
    return float(env.compute_reward(x.reshape(1, 128, 128), state["gt"]).reshape(-1)[0])


def logged(env, log):
    """Wrap ``env.step`` so that every action the driver hands to the environment is recorded."""
    inner = env.step

    def step(states, action_dict):
        log.append([float(action_dict[k]) for k in ("T", "sigma_d", "mu")])
        return inner(states, action_dict)
    env.step = step
    return env


def main():
    ns = ref_shim.load()
    sys.path.insert(0, ref_shim.REF_ROOT)
    import evaluation.eval as reval
    import evaluation.mcts as rmcts
    from transformer.decision_transformer import DecisionTransformer as RefDT, DecisionTransformerConfig as RefCfg
    from dt4image_restoration_b200.policy import DecisionTransformer as OurDT

    params = O.init_unet_params(UNET_SEED, "default")
    item = synth.make_item(synth.phantom(128, 128, ITEM_SEED), synth.radial_mask(128, 128, 0.3), 0.0, ITEM_SEED)
    torch.manual_seed(DT_SEED)
    rdt = RefDT(RefCfg(block_size=18, n_embeds=9, mode='norm')).eval()
    bias_stop_head(rdt)
    torch.manual_seed(DT_SEED)
    odt = OurDT(block_size=18, n_embeds=9, mode='norm')
    bias_stop_head(odt)

    def ref_env(log):
        env = ns.PnPEnv(30, ref_shim.make_denoiser(ns, params), 'cpu')
        env.run_no_ref_reward = lambda state: psnr_stand_in(env, state)
        return logged(env, log)

    out = {}
    with tempfile.TemporaryDirectory() as d:
        ckpt = os.path.join(d, "dt.pt")
        torch.save(rdt.state_dict(), ckpt)

        # ------------------------------------------------ greedy: real reference
        log_ref = []
        ev = reval.Evaluator(rdt, ckpt, 3, 30, ref_env(log_ref), False, 'cpu', 18, 10)
        with torch.no_grad():
            (es, ea, er, _, ets, etk), (st, pr, _, ad) = ev.get_initial_policy_setup(policy_inputs(item), to_t(item))
            r_ref, t_ref, x_ref = ev.run_greedy(st, pr, 1, ad, es, ea, er, ets, etk)
        # ------------------------------------------------ greedy: restated drivers on the reference env + reference DT
        log_a = []
        drv = RD.GreedyDriver(rdt, ref_env(log_a), 'cpu')
        with torch.no_grad():
            (obs, acts, rtgs, ts, tasks), (st, pr, _, ad) = drv.initial_setup(policy_inputs(item), to_t(item))
            r_a, t_a, x_a = drv.run_greedy(st, pr, 1, ad, obs, acts, rtgs, ts, tasks)
        assert t_a == t_ref and torch.equal(r_a, r_ref) and torch.equal(x_a, x_ref), "restated greedy driver differs"
        assert np.array_equal(np.array(log_a), np.array(log_ref))
        # ------------------------------------------------ greedy: restated drivers on the oracle env + package DT
        log_b = []
        drv = RD.GreedyDriver(odt, logged(OracleEnv(params), log_b), 'cpu')
        with torch.no_grad():
            (obs, acts, rtgs, ts, tasks), (st, pr, _, ad) = drv.initial_setup(policy_inputs(item), to_t(item))
            r_b, t_b, x_b = drv.run_greedy(st, pr, 1, ad, obs, acts, rtgs, ts, tasks)
        assert t_b == t_ref and (r_b - r_ref).abs().max() < 1e-4 and (x_b - x_ref).abs().max() < 1e-5
        assert np.abs(np.array(log_b) - np.array(log_ref)).max() < 1e-5
        print(f"greedy: reward {float(r_ref):.4f} dB after {t_ref} steps; restated == reference")
        out.update(greedy_reward=r_ref.numpy(), greedy_time=np.int64(t_ref), greedy_x=x_ref.numpy().astype(np.float32),
                   greedy_actions=np.array(log_ref, dtype=np.float64))

        # ------------------------------------------------ tree search: real reference
        log_ref = []
        env = ref_env(log_ref)
        ev = reval.Evaluator(rdt, ckpt, 3, 30, env, False, 'cpu', 18, 10)
        torch.manual_seed(MCTS_SEED)
        rewards_ref = {}
        real_beam = rmcts.run_beam_search

        def beam_spy(node, evaluator):
            res = real_beam(node, evaluator)
            rewards_ref[repr(node)] = float(res[0])
            return res
        rmcts.run_beam_search = beam_spy
        try:
            with torch.no_grad():
                final_ref = rmcts.run_mcts(ev, policy_inputs(item), to_t(item), policy_inputs(item)[3], env, 'cpu')
        finally:
            rmcts.run_beam_search = real_beam
        # ------------------------------------------------ tree search: restated, reference env + reference DT
        log_a = []
        env = ref_env(log_a)
        drv = RD.GreedyDriver(rdt, env, 'cpu')
        torch.manual_seed(MCTS_SEED)
        with torch.no_grad():
            final_a, best_a, prog_a = RD.run_mcts(drv, policy_inputs(item), to_t(item), policy_inputs(item)[3], env, 'cpu')
        assert torch.equal(final_a, final_ref), "restated tree search differs in the final reward"
        assert list(prog_a) == list(rewards_ref) and all(float(prog_a[k]) == rewards_ref[k] for k in prog_a)
        assert np.array_equal(np.array(log_a), np.array(log_ref))
        print(f"mcts (reference aliasing): final {float(final_ref):.4f} dB, best program {best_a}, {len(log_ref)} env steps; "
              "restated == reference")
        out.update(mcts_final=final_ref.numpy(), mcts_best=np.array(best_a), mcts_keys=np.array(list(prog_a)),
                   mcts_rewards=np.array([float(v) for v in prog_a.values()]), mcts_actions=np.array(log_ref, dtype=np.float64))
        # ------------------------------------------------ tree search with independent children: oracle env + package DT
        log_c = []
        env = logged(OracleEnv(params), log_c)
        drv = RD.GreedyDriver(odt, env, 'cpu')
        torch.manual_seed(MCTS_SEED)
        with torch.no_grad():
            final_c, best_c, prog_c = RD.run_mcts(drv, policy_inputs(item), to_t(item), policy_inputs(item)[3], env, 'cpu',
                                                  independent_children=True)
        print(f"mcts (independent children): final {float(final_c):.4f} dB, best program {best_c}, {len(log_c)} env steps")
        out.update(mcts_indep_final=final_c.numpy(), mcts_indep_best=np.array(best_c), mcts_indep_keys=np.array(list(prog_c)),
                   mcts_indep_rewards=np.array([float(v) for v in prog_c.values()]),
                   mcts_indep_actions=np.array(log_c, dtype=np.float64))
    out["meta"] = np.array(f"dt_seed={DT_SEED} unet_seed={UNET_SEED} item_seed={ITEM_SEED} mcts_seed={MCTS_SEED} t_bias={T_BIAS}")
    np.savez_compressed(os.path.join(GOLD, "ref_drivers.npz"), **out)
    print("written", os.path.join(GOLD, "ref_drivers.npz"), os.path.getsize(os.path.join(GOLD, "ref_drivers.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()

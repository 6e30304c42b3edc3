"""The reference's DRIVERS against the drop-in (SURVEY.md section 4 "drop-in" tier, section 8c).

``oracle/ref_drivers.py`` restates ``Evaluator.get_initial_policy_setup / predict_action_and_rtg / run_greedy``
(``evaluation/eval.py:62-100,147-220``) and ``run_mcts`` with its tree (``evaluation/mcts.py``); ``oracle/make_golden_drivers.py``
pinned it bit-exact to the REAL reference drivers in the build container and wrote ``tests/golden/ref_drivers.npz`` (the GPU
box has no ``/root/reference``).  Here:
  * CPU: the restated drivers on the oracle environment reproduce the reference's recorded decisions;
  * GPU: the restated drivers drive ``dt4image_restoration_b200.env.PnPEnv`` on CUDA (PSNR in place of ARNIQA) and must end
    within 0.05 dB of the reference run, with the same stop time / the same programs;
  * GPU: the package's batched tree search (``mcts.BatchedMCTS``) against the restated search with independent children.
"""
import os

import numpy as np
import pytest
import torch

from dt4image_restoration_b200 import synth
from dt4image_restoration_b200.policy import DecisionTransformer
from oracle import make_golden_drivers as G
from oracle import pnp_oracle as O
from oracle import ref_drivers as RD

TOL_DB = 0.05


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_drivers.npz"))


def make_policy():
    torch.manual_seed(G.DT_SEED)
    pol = DecisionTransformer(block_size=18, n_embeds=9, mode="norm")
    G.bias_stop_head(pol)
    return pol


def make_item():
    return synth.make_item(synth.phantom(128, 128, G.ITEM_SEED), synth.radial_mask(128, 128, 0.3), 0.0, G.ITEM_SEED)


def test_restated_greedy_driver_reproduces_the_reference_run(gold):
    log = []
    item = make_item()
    drv = RD.GreedyDriver(make_policy(), G.logged(G.OracleEnv(O.init_unet_params(G.UNET_SEED, "default")), log), "cpu")
    with torch.no_grad():
        (obs, acts, rtgs, ts, tasks), (st, pr, _, ad) = drv.initial_setup(G.policy_inputs(item), G.to_t(item))
        reward, time, x = drv.run_greedy(st, pr, 1, ad, obs, acts, rtgs, ts, tasks)
    assert time == int(gold["greedy_time"])
    assert np.abs(np.array(log) - gold["greedy_actions"]).max() < 1e-5
    assert abs(float(reward) - float(gold["greedy_reward"].reshape(-1)[0])) < 1e-3
    assert np.abs(x.numpy() - gold["greedy_x"]).max() < 1e-5


def test_restated_tree_search_prefix_reproduces_the_reference_run(gold):
    """Two of the thirty iterations (the CPU suite stays short): every action handed to the environment equals the
    reference run's, in order - selection, expansion with the aliased state dict, the greedy rollouts."""
    log = []
    item = make_item()
    env = G.logged(G.OracleEnv(O.init_unet_params(G.UNET_SEED, "default")), log)
    drv = RD.GreedyDriver(make_policy(), env, "cpu")
    torch.manual_seed(G.MCTS_SEED)
    with torch.no_grad():
        _, _, programs = RD.run_mcts(drv, G.policy_inputs(item), G.to_t(item), G.policy_inputs(item)[3], env, "cpu", n_iters=2)
    n = len(log)
    assert n > 50 and np.abs(np.array(log) - gold["mcts_actions"][:n]).max() < 1e-5
    assert list(programs) == list(gold["mcts_keys"][:2])
    assert np.abs(np.array([float(v) for v in programs.values()]) - gold["mcts_rewards"][:2]).max() < 1e-3


# ------------------------------------------------------------------------------------------------------------------
def _gpu_env():
    from dt4image_restoration_b200.env import PnPEnv
    from dt4image_restoration_b200.noise import UNetDenoiser2D
    env = PnPEnv(30, UNetDenoiser2D(state_dict=O.init_unet_params(G.UNET_SEED, "default")), "cuda")
    env.no_ref_model = lambda state: G.psnr_stand_in(env, state)
    return env


@pytest.mark.gpu
def test_reference_greedy_driver_on_the_dropin(gold):
    log = []
    item = make_item()
    drv = RD.GreedyDriver(make_policy().cuda(), G.logged(_gpu_env(), log), "cuda")
    with torch.no_grad():
        (obs, acts, rtgs, ts, tasks), (st, pr, _, ad) = drv.initial_setup(G.policy_inputs(item), G.to_t(item))
        reward, time, x = drv.run_greedy(st, pr, 1, ad, obs, acts, rtgs, ts, tasks)
    assert time == int(gold["greedy_time"])
    assert reward.device.type == "cpu" and reward.shape == (1, 1)
    assert abs(float(reward) - float(gold["greedy_reward"].reshape(-1)[0])) < TOL_DB
    assert np.abs(np.array(log) - gold["greedy_actions"]).max() < 1e-3
    assert np.abs(x.cpu().numpy() - gold["greedy_x"]).max() < 1e-3


@pytest.mark.gpu
def test_reference_tree_search_on_the_dropin(gold):
    """``run_mcts`` as the reference runs it - ONE aliased state dict per expansion (mcts.py:118-136) - on the CUDA
    drop-in: 30 iterations, 1044 environment steps, same programs and best program, rewards within 0.05 dB."""
    log = []
    item = make_item()
    env = G.logged(_gpu_env(), log)
    drv = RD.GreedyDriver(make_policy().cuda(), env, "cuda")
    torch.manual_seed(G.MCTS_SEED)
    with torch.no_grad():
        final, best, programs = RD.run_mcts(drv, G.policy_inputs(item), G.to_t(item), G.policy_inputs(item)[3], env, "cuda")
    assert list(programs) == list(gold["mcts_keys"])
    assert np.abs(np.array([float(v) for v in programs.values()]) - gold["mcts_rewards"]).max() < TOL_DB
    assert best == str(gold["mcts_best"])
    assert abs(float(final) - float(gold["mcts_final"].reshape(-1)[0])) < TOL_DB
    assert len(log) == len(gold["mcts_actions"]) and np.abs(np.array(log) - gold["mcts_actions"]).max() < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("graph_policy", [True, False])
def test_batched_tree_search_matches_the_cpu_search(gold, graph_policy):
    """``mcts.BatchedMCTS`` (one batched expansion step per iteration, aliasing fixed) vs the restated reference search
    with ``independent_children=True`` on the oracle (fixture): same programs in the same order, same best program.
    ``graph_policy``: policy calls as one CUDA-graph replay on observations encoded once (default) or the eager forwards."""
    from dt4image_restoration_b200.mcts import BatchedMCTS
    from dt4image_restoration_b200.noise import UNetDenoiser2D
    item = make_item()
    den = UNetDenoiser2D(state_dict=O.init_unet_params(G.UNET_SEED, "default"))
    search = BatchedMCTS(make_policy(), den, 128, 128, width=5, n_iters=30, graph_policy=graph_policy)
    torch.manual_seed(G.MCTS_SEED)
    final, best, programs = search.search(G.to_t(item), G.policy_inputs(item)[1], G.policy_inputs(item)[3])
    assert list(programs) == list(gold["mcts_indep_keys"])
    assert np.abs(np.array([float(v) for v in programs.values()]) - gold["mcts_indep_rewards"]).max() < TOL_DB
    assert best == str(gold["mcts_indep_best"])
    assert abs(float(final) - float(gold["mcts_indep_final"].reshape(-1)[0])) < TOL_DB

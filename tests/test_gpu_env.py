"""Drop-in parity of PnPEnv / PnPEngine (CUDA, through the C-ABI) against the reference fixtures and the oracle.

Tolerances are BASELINE.json's: reconstructions within 1e-3 max-abs on [0,1] images and within 0.05 dB PSNR
per trajectory over 30 iterations (random-init = PyTorch-default-init weights); masks / indexing bit-exact.
"""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from dt4image_restoration_b200 import synth
from dt4image_restoration_b200.engine import PnPEngine
from dt4image_restoration_b200.env import PnPEnv, torch_psnr
from dt4image_restoration_b200.noise import UNetDenoiser2D
from oracle import pnp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_X = 1e-3        # max-abs on [0,1] images (BASELINE.json north_star)
TOL_DB = 0.05       # dB per trajectory


def to_t(item):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in item.items()}


def act(T, mu, sg, dev=DEV):
    return OrderedDict({"T": torch.tensor([T], dtype=torch.float32, device=dev),
                        "mu": torch.tensor([mu], dtype=torch.float32, device=dev),
                        "sigma_d": torch.tensor([sg], dtype=torch.float32, device=dev)})


@pytest.fixture(scope="module")
def env_default():
    den = UNetDenoiser2D(state_dict=O.init_unet_params(0, "default"))
    return PnPEnv(30, den, DEV)


def test_step_against_reference_fixture_128(env_default, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_env128_default.npz"))
    item = synth.make_item(synth.phantom(128, 128, 0), synth.radial_mask(128, 128, 0.3), 0.0, 0)
    st = env_default.reset(to_t(item), DEV)
    # reset contract (env.py:57-71)
    assert st["x"].dtype == torch.complex64 and st["z"].dtype == torch.complex64 and st["u"].dtype == torch.complex64
    assert st["mask"].dtype == torch.bool and st["mask"].shape == (1, 1, 128, 128) and st["T"] == 0
    assert torch.equal(st["mask"].cpu().reshape(128, 128), torch.from_numpy(item["mask"][0]).bool())   # bit-exact
    for k, (T, mu, sg) in enumerate(g["actions"]):
        old_x, old_z = st["x"], st["z"]
        old_z_copy = old_z.clone()
        st2, done = env_default.step(st, act(T, mu, sg))
        assert st2 is st and done is False                      # same dict object (env.py:100)
        assert st["x"] is not old_x and st["z"] is not old_z    # re-bound to fresh tensors (env.py:95-97)
        assert torch.equal(old_z, old_z_copy)                   # old tensors stay valid (mcts.py:15,18)
        assert st["x"].dtype == torch.float32 and st["x"].shape == (1, 1, 128, 128)
        assert np.abs(st["x"].cpu().numpy() - g["x_steps"][k]).max() < TOL_X
    assert np.abs(torch.view_as_real(st["z"]).cpu().numpy() - g["z"]).max() < TOL_X
    assert np.abs(torch.view_as_real(st["u"]).cpu().numpy() - g["u"]).max() < TOL_X
    assert abs(st["T"] - float(g["T_final"])) < 1e-9
    # early exit: untouched state, done=True (env.py:79-81)
    x_before = st["x"]
    st3, done = env_default.step(st, act(0.7, 0.5, 0.1))
    assert done is True and st3 is st and st["x"] is x_before
    # observation and reward (env.py:103-116)
    ob = PnPEnv.get_policy_ob(st)
    assert ob.shape == (1, 128 * 128) and ob.dtype == torch.float32
    r = PnPEnv.compute_reward(st["x"].reshape(1, 128, 128), st["gt"])
    assert r.shape == (1, 1) and r.device.type == "cpu"
    assert abs(r.item() - float(np.asarray(g["psnr"]).reshape(-1)[0])) < TOL_DB


def test_trajectory_30_iters_against_reference_fixture(env_default, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_traj128_default.npz"))
    item = synth.make_item(synth.phantom(128, 128, 0), synth.radial_mask(128, 128, 0.3), 0.0, 0)
    sig, mus = synth.fixed_schedule(30)
    st = env_default.reset(to_t(item), DEV)
    for k in range(30):
        st, done = env_default.step(st, act(0.0, float(mus[k]), float(sig[k])))
        p = torch_psnr(st["x"].reshape(1, 128, 128), st["gt"].reshape(1, 128, 128)).item()
        assert abs(p - g["psnr"][k]) < TOL_DB, f"iter {k}: {p} vs {g['psnr'][k]}"
    assert np.abs(st["x"].cpu().numpy() - g["x_final"]).max() < TOL_X
    assert abs(st["T"] - 1.0) < 1e-6


def test_kaiming_fixture_64(golden_dir):
    """Signal-preserving init: every conv matters.  The net is not contractive here, so bf16 rounding (2^-9 per
    activation) is amplified step to step; per-step tolerance is looser and stated."""
    g = np.load(os.path.join(golden_dir, "ref_env64_kaiming.npz"))
    den = UNetDenoiser2D(state_dict=O.init_unet_params(1, "kaiming"))
    env = PnPEnv(30, den, DEV)
    item = synth.make_item(synth.phantom(64, 64, 3), synth.cartesian_mask(64, 64, 4, 3), 10.0, 3)
    st = env.reset(to_t(item), DEV)
    v0 = (st["z"] - st["u"]).real
    _, pre = den(v0, torch.tensor([35.0 / 255], device=DEV), preclamp=True)
    ref_resid = g["preclamp0"] - v0.cpu().numpy()
    got_resid = pre.cpu().numpy() - v0.cpu().numpy()
    assert np.linalg.norm(got_resid - ref_resid) / np.linalg.norm(ref_resid) < 3e-2
    for k in range(5):
        st, _ = env.step(st, act(0.0, 0.2 + 0.15 * k, (35.0 - 6 * k) / 255))
        assert np.abs(st["x"].cpu().numpy() - g["x_steps"][k]).max() < 1e-2


def test_error_behaviour_matches_reference(env_default):
    item = synth.make_batch(2, 64, 64, "cartesian", 4, 0.0, seed0=1)
    st = env_default.reset(to_t(item), DEV)
    a = act(0.0, 0.5, 0.1)
    a["mu"] = torch.tensor([0.3, 0.4], device=DEV)               # per-image mu: RuntimeError (env.py:88 view)
    a["sigma_d"] = torch.tensor([0.1, 0.1], device=DEV)
    with pytest.raises(RuntimeError):
        env_default.step(st, a)
    a = act(0.0, 0.5, 0.1)                                        # sigma numel != B: RuntimeError (noise.py:159)
    with pytest.raises(RuntimeError):
        env_default.step(st, a)
    with pytest.raises(Exception):                                # no CPU path: fails loudly
        cpu_env_state = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in st.items()}
        a = act(0.0, 0.5, 0.1, dev="cpu")
        a["sigma_d"] = torch.tensor([0.1, 0.1])
        env_default.step(OrderedDict(cpu_env_state), a)


def test_mcts_style_aliasing(env_default):
    """expand_tree (mcts.py:118-136): six steps on ONE shared dict with a mutated action dict; nodes keep
    references to earlier x tensors, which must stay intact."""
    item = synth.make_item(synth.phantom(128, 128, 4), synth.radial_mask(128, 128, 0.2), 0.0, 4)
    st = env_default.reset(to_t(item), DEV)
    a = act(0.1, 0.5, 0.1)
    kept, copies = [], []
    for i in range(6):
        a["sigma_d"] = torch.tensor(0.05 + 0.02 * i, device=DEV)          # 0-d tensors as mcts.py:123-125
        a["mu"] = torch.tensor(0.3 + 0.1 * i, device=DEV)
        states, _ = env_default.step(st, a)
        assert states is st
        kept.append(states["x"])
        copies.append(states["x"].clone())
    for k, c in zip(kept, copies):
        assert torch.equal(k, c)
    assert abs(st["T"] - 6 / 30) < 1e-9
    assert len({t.data_ptr() for t in kept}) == 6


def test_config1_env_256_radial30_b1_30_iterations(env_default):
    """BASELINE config 1 through the drop-in: ONE 256x256 phantom, radial 30 % mask, 30 PnP-ADMM iterations with the fixed
    (sigma, mu) schedule, random-init U-Net, ``PnPEnv.reset/step`` (general-mask cluster kernel) vs the oracle."""
    H = W = 256
    item = synth.make_item(synth.phantom(H, W, 1), synth.radial_mask(H, W, 0.3), 0.0, 1)
    params = O.init_unet_params(0, "default")
    sig, mus = synth.fixed_schedule(30)
    st = env_default.reset(to_t(item), DEV)
    ref = O.reset(item)
    assert torch.equal(st["mask"].cpu(), ref["mask"])                    # masks bit-exact
    worst_x = worst_db = 0.0
    for k in range(30):
        st, done = env_default.step(st, act(0.0, float(mus[k]), float(sig[k])))
        ref, _ = O.step(params, ref, {"T": torch.zeros(1), "mu": torch.tensor([mus[k]]), "sigma_d": torch.tensor([sig[k]])})
        assert done is False
        dx = (st["x"].cpu() - ref["x"]).abs().max().item()
        db = abs(torch_psnr(st["x"].reshape(1, H, W), st["gt"].reshape(1, H, W)).item()
                 - O.psnr(ref["x"].reshape(1, H, W), ref["gt"].reshape(1, H, W)).item())
        worst_x, worst_db = max(worst_x, dx), max(worst_db, db)
        assert dx < TOL_X and db < TOL_DB, f"iteration {k}: max|dx| {dx:.2e}, dPSNR {db:.2e} dB"
    assert (st["z"].cpu() - ref["z"]).abs().max() < TOL_X and (st["u"].cpu() - ref["u"]).abs().max() < TOL_X
    assert abs(st["T"] - 1.0) < 1e-6
    print(f"config 1: worst max|dx| {worst_x:.2e}, worst dPSNR {worst_db:.2e} dB over 30 iterations")


def test_config3_candidate_slice_256_radial20():
    """BASELINE config 3 at reduced width: 16 candidates = one shared 256x256 state (radial 20 %) x 16 (sigma_d, mu)
    samples drawn as evaluation/mcts.py:64-70,114-116, one step each through the batched expander, PSNR rewards vs the
    oracle stepping every candidate on its own (the reference's scalar-mu semantics)."""
    from dt4image_restoration_b200.rollout import CandidateExpander
    H = W = 256
    K = 16
    params = O.init_unet_params(0, "default")
    item = synth.make_item(synth.phantom(H, W, 5), synth.radial_mask(H, W, 0.2), 0.0, 5)
    ref0 = O.reset(item)
    # two warm-up iterations so that z, u are not the trivial initial state
    for k in range(2):
        ref0, _ = O.step(params, ref0, {"T": torch.zeros(1), "mu": torch.tensor([0.3]), "sigma_d": torch.tensor([0.15])})
    g = torch.Generator().manual_seed(7)
    sg, mu = CandidateExpander.sample_actions(0.12, 0.4, K, generator=g)
    eng = PnPEngine(UNetDenoiser2D(state_dict=params), K, H, W, DEV)
    state = {k: ref0[k].to(DEV) for k in ("z", "u", "y0", "mask", "gt")}
    rewards = CandidateExpander(eng).expand(state, sg, mu).cpu()
    for c in range(K):
        one = OrderedDict((k, (v.clone() if torch.is_tensor(v) else v)) for k, v in ref0.items())
        one, _ = O.step(params, one, {"T": torch.zeros(1), "mu": mu[c:c + 1], "sigma_d": sg[c:c + 1]})
        assert (eng.x[c:c + 1].cpu() - one["x"]).abs().max() < TOL_X, f"candidate {c}"
        r_ref = O.psnr(one["x"].reshape(1, H, W), one["gt"].reshape(1, H, W)).item()
        assert abs(rewards[c].item() - r_ref) < TOL_DB, f"candidate {c}"
    assert int(rewards.argmax()) == int(torch.tensor(
        [O.psnr(O.step(params, OrderedDict((k, (v.clone() if torch.is_tensor(v) else v)) for k, v in ref0.items()),
                       {"T": torch.zeros(1), "mu": mu[c:c + 1], "sigma_d": sg[c:c + 1]})[0]["x"].reshape(1, H, W),
                ref0["gt"].reshape(1, H, W)).item() for c in range(K)]).argmax())


def test_kaiming_30_iterations_drift_64():
    """Signal-carrying (kaiming) weights over a whole 30-iteration trajectory.  The default-init net is nearly constant in
    its input, so the 1e-3 trajectories above mostly test the prox; here every conv matters and bf16
    rounding (2^-9 per activation) feeds back through 30 denoiser calls.  Measured on B200 (round 2): worst max|dx| 7.4e-3
    (7x north_star's 1e-3, which is stated for the default init), worst dPSNR 1.7e-3 dB (30x inside the 0.05 dB budget);
    the bound asserted is 1e-2 / 0.05 dB and the measured drift is printed."""
    H = W = 64
    params = O.init_unet_params(1, "kaiming")
    env = PnPEnv(30, UNetDenoiser2D(state_dict=params), DEV)
    item = synth.make_item(synth.phantom(H, W, 3), synth.cartesian_mask(H, W, 4, 3), 10.0, 3)
    sig, mus = synth.fixed_schedule(30)
    st = env.reset(to_t(item), DEV)
    ref = O.reset(item)
    worst_x = worst_db = 0.0
    for k in range(30):
        st, _ = env.step(st, act(0.0, float(mus[k]), float(sig[k])))
        ref, _ = O.step(params, ref, {"T": torch.zeros(1), "mu": torch.tensor([mus[k]]), "sigma_d": torch.tensor([sig[k]])})
        dx = (st["x"].cpu() - ref["x"]).abs().max().item()
        db = abs(torch_psnr(st["x"].reshape(1, H, W), st["gt"].reshape(1, H, W)).item()
                 - O.psnr(ref["x"].reshape(1, H, W), ref["gt"].reshape(1, H, W)).item())
        worst_x, worst_db = max(worst_x, dx), max(worst_db, db)
    print(f"kaiming 30 iterations at 64x64: worst max|dx| {worst_x:.2e}, worst dPSNR {worst_db:.2e} dB")
    assert worst_x < 1e-2 and worst_db < TOL_DB


@pytest.mark.parametrize("B,H,W,kind,par,sn", [(4, 256, 256, "cartesian", 4, 0.0), (2, 128, 128, "radial", 0.2, 0.0),
                                               (2, 512, 512, "cartesian", 8, 10.0), (2, 256, 256, "radial", 0.2, 0.0)])
def test_engine_batched_trajectory_vs_oracle(B, H, W, kind, par, sn):
    """BASELINE configs 2-4 at reduced batch, full 30 iterations: batched engine (per-image sigma and mu) vs the batched
    oracle."""
    n_iters = 30
    params = O.init_unet_params(0, "default")
    batch = synth.make_batch(B, H, W, kind, par, sn, seed0=20)
    sig, mus = synth.fixed_schedule(30)
    eng = PnPEngine(UNetDenoiser2D(state_dict=params), B, H, W, DEV)
    eng.reset(to_t(batch))
    st = O.reset(batch)
    scale = torch.linspace(0.8, 1.2, B)
    for k in range(n_iters):
        sg = float(sig[k]) * scale
        eng.set_actions(sg, float(mus[k]))
        eng.step()
        st, _ = O.step(params, st, {"T": torch.zeros(1), "mu": torch.tensor([mus[k]]), "sigma_d": sg})
        if k in (0, 4, n_iters - 1):
            assert (eng.x.cpu() - st["x"]).abs().max() < TOL_X, f"iter {k}"
            p_ref = O.psnr(st["x"].reshape(B, H, W), st["gt"].reshape(B, H, W)).reshape(-1)
            assert (eng.psnr().cpu() - p_ref).abs().max() < TOL_DB
    assert (eng.u.cpu() - st["u"]).abs().max() < TOL_X
    assert (eng.z.cpu() - st["z"]).abs().max() < TOL_X


def test_engine_per_image_mu_and_env_equivalence():
    params = O.init_unet_params(3, "default")
    B, H, W = 3, 64, 64
    batch = synth.make_batch(B, H, W, "radial", 0.3, 5.0, seed0=7)
    den = UNetDenoiser2D(state_dict=params)
    eng = PnPEngine(den, B, H, W, DEV)
    eng.reset(to_t(batch))
    mu = torch.tensor([0.2, 0.5, 0.9])
    sg = torch.tensor([0.15, 0.1, 0.05])
    eng.set_actions(sg, mu)
    eng.step()
    env = PnPEnv(30, den, DEV)
    for b in range(B):                                   # the reference semantics: one image, scalar mu
        one = {k: v[b:b + 1] for k, v in batch.items()}
        st = env.reset(to_t(one), DEV)
        st, _ = env.step(st, act(0.0, float(mu[b]), float(sg[b])))
        # batch 1 and batch 3 may cut the deep convs' K loop differently (split-K cluster kernel): fp32 summation order only
        assert (st["x"] - eng.x[b:b + 1]).abs().max() < 5e-5
        assert (st["u"] - eng.u[b:b + 1]).abs().max() < 5e-5


def test_engine_step_is_cuda_graph_capturable():
    params = O.init_unet_params(0, "default")
    B, H, W = 2, 128, 128
    batch = synth.make_batch(B, H, W, "cartesian", 4, 0.0, seed0=3)
    eng = PnPEngine(UNetDenoiser2D(state_dict=params), B, H, W, DEV)
    eng.reset(to_t(batch))
    eng.set_actions(0.1, 0.5)
    eng.step()                                           # warm-up (plan creation etc.)
    ref = PnPEngine(eng.denoiser, B, H, W, DEV)
    ref.reset(to_t(batch))
    ref.set_actions(0.1, 0.5)
    for _ in range(4):
        ref.step()
    eng.reset(to_t(batch))
    eng.set_actions(0.1, 0.5)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.step()
    eng.reset(to_t(batch))
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    assert (eng.x - ref.x).abs().max() < 1e-6
    assert (eng.u - ref.u).abs().max() < 1e-6


@pytest.mark.parametrize("H,kind,par", [(256, "cartesian", 4), (256, "radial", 0.3), (128, "radial", 0.3), (64, "radial", 0.3),
                                        (64, "cartesian", 4)])
def test_engine_step_active_predicate(H, kind, par):
    """``PnPEngine.step(active)``: the batched early exit (reference env.py:79-81) as a per-image predicate inside the kernels
    (last conv epilogue, every prox kernel): inactive images keep x, z, u, v bit for bit, active ones equal a plain step."""
    B = 5
    params = O.init_unet_params(0, "default")
    batch = synth.make_batch(B, H, H, kind, par, 0.0, seed0=9)
    den = UNetDenoiser2D(state_dict=params)
    a, b = PnPEngine(den, B, H, H, DEV), PnPEngine(den, B, H, H, DEV)
    for e in (a, b):
        e.reset(to_t(batch))
        e.set_actions(0.1, 0.4)
        e.step()                                   # a non-trivial state; also lets the mask-kind hint land
    torch.cuda.synchronize()
    active = torch.tensor([True, False, True, True, False], device=DEV)
    before = [t.clone() for t in (a.x, a.z, a.u, a.v)]
    for kind_known in (True, False):
        for e in (a, b):
            e.set_actions(0.07, 0.6)
        if not kind_known:
            a.probe.kind = b.probe.kind = -1       # both prox kernels launched, the device flag decides
            a.probe.event = b.probe.event = type("E", (), {"query": staticmethod(lambda: False)})()
        a.step(active)
        b.step()
        for t_a, t_b, t_0 in zip((a.x, a.z, a.u, a.v), (b.x, b.z, b.u, b.v), before):
            assert torch.equal(t_a[~active], t_0[~active])
            assert torch.equal(t_a[active], t_b[active])
        # bring b's inactive images back in line with a for the second round
        for t_a, t_b in zip((a.x, a.z, a.u, a.v), (b.x, b.z, b.u, b.v)):
            t_b.copy_(t_a)
        before = [t.clone() for t in (a.x, a.z, a.u, a.v)]


def test_engine_step_active_predicate_any_size_path():
    """Sizes on the dense-DFT prox path (no prepared constants) take the engine's select fallback for ``active``: inactive
    images keep their state bit for bit, active ones equal a plain step."""
    B, H, W = 4, 40, 48
    params = O.init_unet_params(0, "default")
    batch = synth.make_batch(B, H, W, "radial", 0.4, 0.0, seed0=3)
    den = UNetDenoiser2D(state_dict=params)
    a, b = PnPEngine(den, B, H, W, DEV), PnPEngine(den, B, H, W, DEV)
    assert not a.prepared
    for e in (a, b):
        e.reset(to_t(batch))
        e.set_actions(0.1, 0.4)
        e.step()
    active = torch.tensor([False, True, True, False], device=DEV)
    before = [t.clone() for t in (a.x, a.z, a.u, a.v)]
    a.step(active)
    b.step()
    for t_a, t_b, t_0 in zip((a.x, a.z, a.u, a.v), (b.x, b.z, b.u, b.v), before):
        assert torch.equal(t_a[~active], t_0[~active])
        assert torch.equal(t_a[active], t_b[active])


ANYSIZE_CASES = [(130, 130, "radial", 0.3, 0.0, 4), (136, 120, "cartesian", 4, 5.0, 5), (45, 51, "radial", 0.4, 0.0, 6)]


@pytest.mark.parametrize("case", ANYSIZE_CASES, ids=lambda c: f"{c[0]}x{c[1]}")
def test_step_at_non_power_of_two_sizes_against_reference_fixture(env_default, golden_dir, case):
    """The reference's ``step`` works at non-power-of-two sizes (torch.fft is mixed-radix) - so does the
    drop-in (dense-DFT prox path, ragged U-Net levels): three steps against states saved from the REAL reference
    (oracle/make_golden_anysize.py), then the fixed 30-iteration schedule against the oracle."""
    H, W, kind, par, sn, seed = case
    g = np.load(os.path.join(golden_dir, "ref_env_anysize.npz"))
    mask = synth.radial_mask(H, W, par) if kind == "radial" else synth.cartesian_mask(H, W, par, seed)
    item = synth.make_item(synth.phantom(H, W, seed), mask, sn, seed)
    st = env_default.reset(to_t(item), DEV)
    assert torch.equal(st["mask"].cpu().reshape(H, W), torch.from_numpy(item["mask"][0]).bool())
    for k, (T, mu, sg) in enumerate(g["actions"]):
        st, done = env_default.step(st, act(T, mu, sg))
        assert done is False and st["x"].shape == (1, 1, H, W)
        assert np.abs(st["x"].cpu().numpy() - g[f"x_steps_{H}x{W}"][k]).max() < TOL_X
    assert np.abs(torch.view_as_real(st["z"]).cpu().numpy() - g[f"z_{H}x{W}"]).max() < TOL_X
    assert np.abs(torch.view_as_real(st["u"]).cpu().numpy() - g[f"u_{H}x{W}"]).max() < TOL_X
    params = O.init_unet_params(0, "default")
    sig, mus = synth.fixed_schedule(30)
    st = env_default.reset(to_t(item), DEV)
    ref = O.reset(item)
    for k in range(30):
        st, _ = env_default.step(st, act(0.0, float(mus[k]), float(sig[k])))
        ref, _ = O.step(params, ref, {"T": torch.zeros(1), "mu": torch.tensor([mus[k]]), "sigma_d": torch.tensor([sig[k]])})
        assert (st["x"].cpu() - ref["x"]).abs().max().item() < TOL_X, f"iteration {k}"
    db = abs(torch_psnr(st["x"].reshape(1, H, W), st["gt"].reshape(1, H, W)).item()
             - O.psnr(ref["x"].reshape(1, H, W), ref["gt"].reshape(1, H, W)).item())
    assert db < TOL_DB and (st["u"].cpu() - ref["u"]).abs().max() < TOL_X


def test_engine_at_non_power_of_two_size_vs_oracle():
    """Batched engine (un-prepared prox path) at 136x120, B = 3, per-image masks, 6 iterations."""
    B, H, W = 3, 136, 120
    batch = synth.make_batch(B, H, W, "radial", 0.3, sigma_n=0.0, seed0=11)
    params = O.init_unet_params(0, "default")
    eng = PnPEngine(UNetDenoiser2D(state_dict=params), B, H, W, DEV)
    eng.reset(to_t(batch))
    ref = O.reset(batch)
    sig, mus = synth.fixed_schedule(30)
    for k in range(6):
        eng.set_actions(torch.full((B,), float(sig[k])), torch.full((B,), float(mus[k])))
        eng.step()
        ref, _ = O.step(params, ref, {"T": torch.zeros(1), "mu": torch.tensor([mus[k]]), "sigma_d": torch.full((B,), float(sig[k]))})
        assert (eng.x.cpu() - ref["x"]).abs().max().item() < TOL_X
    assert (eng.u.cpu() - ref["u"]).abs().max().item() < TOL_X


@pytest.mark.parametrize("H,W", [(1040, 32), (16, 2000)])
def test_unsupported_sizes_are_rejected_loudly(env_default, H, W):
    """Sizes outside 2..1024 (FFT-prox) or below 16 (four 2x2 pools of the U-Net) are REJECTED with a ``PnpError`` (a
    ``RuntimeError``) from the first step - the drop-in never computes something else."""
    from dt4image_restoration_b200._lib import PnpError
    assert issubclass(PnpError, RuntimeError)
    item = synth.make_item(synth.phantom(H, W, 0), (np.random.default_rng(0).random((H, W)) < 0.3).astype(np.uint8), 0.0, 0)
    st = env_default.reset(to_t(item), DEV)
    with pytest.raises(PnpError):
        env_default.step(st, act(0.0, 0.5, 0.1))

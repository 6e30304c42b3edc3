"""Generate ``tests/golden/*.npz`` by running the REAL reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden

For every case the reference's own code (``PnPEnv.reset/step``, ``UNetDenoiser2D``, ``fft/ifft``,
``torch_psnr``) is executed on seeded synthetic inputs and seeded random-init weights, the oracle
restatement (``oracle/pnp_oracle.py``) is asserted equal to it, and the reference outputs are saved
as small fixtures.  Inputs are NOT stored: they are regenerated from seeds by
``dt4image_restoration_b200.synth`` and ``oracle.pnp_oracle.init_unet_params``; their SHA-256 is
stored so drift in the generators is detected.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dt4image_restoration_b200 import synth  # noqa: E402
from oracle import pnp_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def params_sha(params) -> str:
    h = hashlib.sha256()
    for k, v in params.items():
        h.update(k.encode())
        h.update(v.numpy().tobytes())
    return h.hexdigest()


def to_t(item):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in item.items()}


def ref_states_any_shape(ns, item):
    """Reference reset semantics (env.py:57-71) for a shape its literal 128s reject."""
    d = to_t(item)
    H, W = d["gt"].shape[-2:]
    x = torch.view_as_complex(d["x0"].contiguous())
    return OrderedDict({"x": x, "y0": torch.view_as_complex(d["y0"].contiguous()), "z": x.clone().detach(),
                        "u": torch.zeros_like(x), "mask": d["mask"].reshape(1, 1, H, W).contiguous().to(torch.bool),
                        "gt": d["gt"], "ATy0": d["ATy0"][..., 0], "T": 0, "complex_y0": d["y0"]})


def act(T, mu, sigma):
    return OrderedDict({"T": torch.tensor([T], dtype=torch.float32), "mu": torch.tensor([mu], dtype=torch.float32),
                        "sigma_d": torch.tensor([sigma], dtype=torch.float32)})


def eq(a, b, what, tol=0.0):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    d = (a - b).abs().max().item() if a.numel() else 0.0
    assert d <= tol, f"oracle != reference for {what}: max|d|={d}"
    return d


def main():
    torch.set_num_threads(1)  # one thread -> the reference's CPU results are run-to-run reproducible
    ns = ref_shim.load()
    os.makedirs(GOLD, exist_ok=True)
    meta = {"torch": torch.__version__, "numpy": np.__version__, "cases": {}}

    # ---- case 1: the reference's native shape, its own reset + 3 steps, default init --------------
    H = W = 128
    params = O.init_unet_params(seed=0, kind="default")
    item = synth.make_item(synth.phantom(H, W, 0), synth.radial_mask(H, W, 0.3), 0.0, 0)
    den = ref_shim.make_denoiser(ns, params)
    env = ns.PnPEnv(30, den, "cpu")
    rs = env.reset(to_t(item), "cpu")
    os_ = O.reset(item)
    for k in ("x", "y0", "z", "u", "gt"):
        eq(torch.view_as_real(rs[k]) if rs[k].is_complex() else rs[k],
           torch.view_as_real(os_[k]) if os_[k].is_complex() else os_[k], f"reset.{k}")
    assert torch.equal(rs["mask"], os_["mask"])
    acts = [(0.0, 0.30, 40.0 / 255), (0.1, 0.55, 20.0 / 255), (0.2, 0.90, 8.0 / 255)]
    xs = []
    for (T, mu, sg) in acts:
        rs, rdone = env.step(rs, act(T, mu, sg))
        os_, odone = O.step(params, os_, act(T, mu, sg))
        assert rdone == odone is False
        eq(rs["x"], os_["x"], "step.x")
        eq(torch.view_as_real(rs["z"]), torch.view_as_real(os_["z"]), "step.z")
        eq(torch.view_as_real(rs["u"]), torch.view_as_real(os_["u"]), "step.u")
        xs.append(rs["x"].numpy().copy())
    # early exit (env.py:79-81)
    rs2, rdone = env.step(rs, act(0.7, 0.5, 0.1))
    assert rdone is True and rs2 is rs
    ob_r = ns.PnPEnv.get_policy_ob(rs)
    eq(ob_r, O.policy_ob(os_), "policy_ob")
    pr = ns.PnPEnv.compute_reward(rs["x"].reshape(1, 128, 128), rs["gt"])
    eq(pr, O.psnr(os_["x"].reshape(1, H, W), os_["gt"].reshape(1, H, W)), "psnr")
    np.savez_compressed(os.path.join(GOLD, "ref_env128_default.npz"), x_steps=np.stack(xs),
                        z=torch.view_as_real(rs["z"]).numpy(), u=torch.view_as_real(rs["u"]).numpy(),
                        psnr=pr.numpy(), actions=np.array(acts, dtype=np.float64), T_final=np.float64(rs["T"]))
    meta["cases"]["ref_env128_default"] = {"H": H, "W": W, "mask": "radial 0.3", "init": "default seed 0",
                                           "item_sha": {k: sha(v) for k, v in item.items()},
                                           "params_sha": params_sha(params)}

    # ---- case 2: 30-iteration trajectory at 128^2, fixed schedule, default init --------------------
    sig, mus = synth.fixed_schedule(30)
    rs = env.reset(to_t(item), "cpu")
    os_ = O.reset(item)
    psn = []
    for k in range(30):
        rs, _ = env.step(rs, act(0.0, float(mus[k]), float(sig[k])))
        os_, _ = O.step(params, os_, act(0.0, float(mus[k]), float(sig[k])))
        psn.append(ns.torch_psnr(rs["x"].reshape(1, H, W), rs["gt"].reshape(1, H, W)).item())
    eq(rs["x"], os_["x"], "traj.x")
    np.savez_compressed(os.path.join(GOLD, "ref_traj128_default.npz"), x_final=rs["x"].numpy(),
                        psnr=np.array(psn, dtype=np.float64))
    meta["cases"]["ref_traj128_default"] = {"schedule": "synth.fixed_schedule(30)", "init": "default seed 0"}

    # ---- case 3: signal-preserving init, 64^2 (reference step is shape-agnostic for B=1) ----------
    H = W = 64
    params_k = O.init_unet_params(seed=1, kind="kaiming")
    item64 = synth.make_item(synth.phantom(H, W, 3), synth.cartesian_mask(H, W, 4, 3), 10.0, 3)
    den_k = ref_shim.make_denoiser(ns, params_k)
    env_k = ns.PnPEnv(30, den_k, "cpu")
    rs = ref_states_any_shape(ns, item64)
    os_ = O.reset(item64)
    xs, pre = [], None
    for k in range(5):
        a = act(0.0, 0.2 + 0.15 * k, (35.0 - 6 * k) / 255)
        if k == 0:
            v = (rs["z"] - rs["u"]).real
            nm = torch.ones(1, 1, H, W) * a["sigma_d"].view(1, 1, 1, 1)
            pre = den_k.net(torch.cat([v, nm], dim=1))
            eq(pre, O.denoise(params_k, v, a["sigma_d"], clamp=False), "preclamp")
        rs, _ = env_k.step(rs, a)
        os_, _ = O.step(params_k, os_, a)
        eq(rs["x"], os_["x"], "k.step.x")
        eq(torch.view_as_real(rs["u"]), torch.view_as_real(os_["u"]), "k.step.u")
        xs.append(rs["x"].numpy().copy())
    np.savez_compressed(os.path.join(GOLD, "ref_env64_kaiming.npz"), x_steps=np.stack(xs), preclamp0=pre.numpy(),
                        z=torch.view_as_real(rs["z"]).numpy(), u=torch.view_as_real(rs["u"]).numpy())
    meta["cases"]["ref_env64_kaiming"] = {"H": H, "W": W, "mask": "cartesian 4x seed 3", "sigma_n": 10.0,
                                          "init": "kaiming seed 1",
                                          "item_sha": {k: sha(v) for k, v in item64.items()},
                                          "params_sha": params_sha(params_k)}

    # ---- case 4: U-Net alone, odd-ish sizes (pad path of `up`, noise.py:49-53) ---------------------
    g = torch.Generator().manual_seed(99)
    net = ns.UNet(2, 1)
    net.load_state_dict(params_k)
    net.eval()
    outs = {}
    for (h, w) in ((64, 64), (48, 80), (36, 52)):
        inp = torch.rand(2, 2, h, w, generator=g)
        r = net(inp)
        eq(r, O.unet_forward(params_k, inp), f"unet {h}x{w}", tol=0.0)
        outs[f"out_{h}x{w}"] = r.numpy()
    np.savez_compressed(os.path.join(GOLD, "ref_unet_kaiming.npz"), **outs)
    meta["cases"]["ref_unet_kaiming"] = {"input": "torch.rand(2,2,h,w, Generator(99)) drawn in order 64x64, 48x80, 36x52"}

    # ---- case 5: centred fft / ifft and psnr ------------------------------------------------------
    g = torch.Generator().manual_seed(7)
    fo = {}
    for (h, w) in ((32, 32), (64, 48), (128, 128), (30, 34)):
        zc = torch.complex(torch.randn(2, 1, h, w, generator=g), torch.randn(2, 1, h, w, generator=g))
        f_, i_ = ns.fft(zc), ns.ifft(zc)
        eq(torch.view_as_real(f_), torch.view_as_real(O.centered_fft2(zc)), "fft")
        eq(torch.view_as_real(i_), torch.view_as_real(O.centered_ifft2(zc)), "ifft")
        fo[f"fft_{h}x{w}"] = torch.view_as_real(f_).numpy()
        fo[f"ifft_{h}x{w}"] = torch.view_as_real(i_).numpy()
    a = torch.rand(3, 40, 40, generator=g) * 1.4 - 0.2
    b = torch.rand(3, 40, 40, generator=g)
    fo["psnr"] = ns.torch_psnr(a, b).numpy()
    eq(ns.torch_psnr(a, b), O.psnr(a, b), "psnr")
    np.savez_compressed(os.path.join(GOLD, "ref_fft_psnr.npz"), **fo)
    meta["cases"]["ref_fft_psnr"] = {"seed": 7}

    # ---- case 6: the policy that produces the actions (reference DecisionTransformer, seeded init) ------------
    sys.path.insert(0, ref_shim.REF_ROOT)
    from transformer.decision_transformer import DecisionTransformer as RefDT, DecisionTransformerConfig as RefCfg
    from dt4image_restoration_b200.policy import DecisionTransformer as OurDT
    torch.manual_seed(1234)
    rdt = RefDT(RefCfg(block_size=18, n_embeds=9, mode='norm')).eval()
    torch.manual_seed(1234)
    odt = OurDT(block_size=18, n_embeds=9, mode='norm')
    rsd, osd = rdt.state_dict(), odt.state_dict()
    assert list(rsd) == list(osd), "policy state_dict keys differ from the reference"
    for k in rsd:
        assert torch.equal(rsd[k], osd[k]), f"seeded init differs at {k}"
    g = torch.Generator().manual_seed(55)
    B, K = 1, 6
    rtg = torch.rand(B, K, 1, generator=g); st = torch.rand(B, K, 128 * 128, generator=g)
    ts = torch.arange(K).reshape(1, K, 1); task = torch.full((B, K), 3, dtype=torch.long)
    acts = torch.rand(B, K, 3, generator=g)
    with torch.no_grad():
        ra, rad = rdt(rtg, st, ts, task, acts, eval_actions=True)
        rr = rdt(rtg, st, ts, task, acts, eval_rtg=True)
        ra0, _ = rdt(rtg, st, ts, task, actions=None)
    oa, oad = odt(rtg, st, ts, task, acts, eval_actions=True)
    orr = odt(rtg, st, ts, task, acts, eval_rtg=True)
    oa0, _ = odt(rtg, st, ts, task, actions=None)
    eq(ra, oa, "dt actions", tol=2e-6); eq(rr, orr, "dt rtg", tol=2e-6); eq(ra0, oa0, "dt actions (no act tokens)", tol=2e-6)
    assert list(rad) == list(oad)
    np.savez_compressed(os.path.join(GOLD, "ref_dt_seed1234.npz"), actions=ra.numpy(), rtg=rr.numpy(), actions_noact=ra0.numpy())
    meta["cases"]["ref_dt_seed1234"] = {"init": "torch.manual_seed(1234) then construct", "inputs": "Generator(55)"}

    # ---- synth generators pinned by hash ----------------------------------------------------------
    meta["synth_sha"] = {
        "phantom_128_s0": sha(synth.phantom(128, 128, 0)),
        "phantom_256_s5": sha(synth.phantom(256, 256, 5)),
        "radial_256_0.3": sha(synth.radial_mask(256, 256, 0.3)),
        "radial_256_0.2": sha(synth.radial_mask(256, 256, 0.2)),
        "cartesian_256_4_s0": sha(synth.cartesian_mask(256, 256, 4, 0)),
        "cartesian_512_8_s1": sha(synth.cartesian_mask(512, 512, 8, 1)),
    }
    meta["synth_frac"] = {
        "radial_256_0.3": float(synth.radial_mask(256, 256, 0.3).mean()),
        "radial_256_0.2": float(synth.radial_mask(256, 256, 0.2).mean()),
        "cartesian_256_4_s0": float(synth.cartesian_mask(256, 256, 4, 0).mean()),
    }
    with open(os.path.join(GOLD, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("golden written to", GOLD)
    for fn in sorted(os.listdir(GOLD)):
        print(f"  {fn}: {os.path.getsize(os.path.join(GOLD, fn)) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()

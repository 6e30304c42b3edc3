"""The CPU oracle (oracle/pnp_oracle.py) against fixtures produced by the REAL reference
(oracle/make_golden.py imported /root/reference and wrote tests/golden/*.npz)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from dt4image_restoration_b200 import synth
from oracle import pnp_oracle as O

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def meta(golden_dir):
    return json.load(open(os.path.join(golden_dir, "meta.json")))


def act(T, mu, sg):
    return {"T": torch.tensor([T], dtype=torch.float32), "mu": torch.tensor([mu], dtype=torch.float32),
            "sigma_d": torch.tensor([sg], dtype=torch.float32)}


def params_sha(params):
    h = hashlib.sha256()
    for k, v in params.items():
        h.update(k.encode())
        h.update(v.numpy().tobytes())
    return h.hexdigest()


def test_synth_generators_pinned(meta):
    s = meta["synth_sha"]
    assert sha(synth.phantom(128, 128, 0)) == s["phantom_128_s0"]
    assert sha(synth.phantom(256, 256, 5)) == s["phantom_256_s5"]
    assert sha(synth.radial_mask(256, 256, 0.3)) == s["radial_256_0.3"]
    assert sha(synth.radial_mask(256, 256, 0.2)) == s["radial_256_0.2"]
    assert sha(synth.cartesian_mask(256, 256, 4, 0)) == s["cartesian_256_4_s0"]
    assert sha(synth.cartesian_mask(512, 512, 8, 1)) == s["cartesian_512_8_s1"]
    assert 0.30 <= synth.radial_mask(256, 256, 0.3).mean() < 0.31
    assert 0.20 <= synth.radial_mask(256, 256, 0.2).mean() < 0.21
    assert synth.cartesian_mask(256, 256, 4, 0).mean() == 0.25
    assert synth.cartesian_mask(512, 512, 8, 1).mean() == 0.125


def test_unet_param_inventory():
    shapes = O.unet_param_shapes()
    assert len(shapes) == 56
    assert sum(int(np.prod(s)) for s in shapes.values()) == 11773857
    assert list(shapes)[0] == "inc.conv.conv-0.conv2d.weight" and list(shapes)[-1] == "outc.conv.bias"


@pytest.mark.parametrize("seed,kind", [(0, "default"), (3, "default"), (1, "kaiming"), (5, "kaiming")])
def test_package_weight_generator_equals_oracle(seed, kind):
    """bench.py and the tools draw their synthetic weights from the package (no oracle import on the product path);
    the tests draw them from the oracle: both must be the same bits, in the same key order."""
    from dt4image_restoration_b200.noise import random_init_state_dict
    a, b = random_init_state_dict(seed, kind), O.init_unet_params(seed, kind)
    assert list(a) == list(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k


def test_env128_default_against_reference(golden_dir, meta):
    g = np.load(os.path.join(golden_dir, "ref_env128_default.npz"))
    m = meta["cases"]["ref_env128_default"]
    params = O.init_unet_params(0, "default")
    assert params_sha(params) == m["params_sha"]
    item = synth.make_item(synth.phantom(128, 128, 0), synth.radial_mask(128, 128, 0.3), 0.0, 0)
    for k, v in item.items():
        assert sha(v) == m["item_sha"][k], k
    st = O.reset(item)
    for k, (T, mu, sg) in enumerate(g["actions"]):
        st, done = O.step(params, st, act(T, mu, sg))
        assert done is False
        np.testing.assert_allclose(st["x"].numpy(), g["x_steps"][k], rtol=0, atol=2e-6)
    np.testing.assert_allclose(torch.view_as_real(st["z"]).numpy(), g["z"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(torch.view_as_real(st["u"]).numpy(), g["u"], rtol=0, atol=2e-6)
    assert abs(st["T"] - float(g["T_final"])) < 1e-12
    st2, done = O.step(params, st, act(0.7, 0.5, 0.1))          # T > 0.5: untouched, done (env.py:79-81)
    assert done is True and st2 is st
    p = O.psnr(st["x"].reshape(1, 128, 128), st["gt"].reshape(1, 128, 128))
    np.testing.assert_allclose(p.numpy(), g["psnr"], rtol=0, atol=1e-4)


def test_traj128_default_against_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_traj128_default.npz"))
    params = O.init_unet_params(0, "default")
    item = synth.make_item(synth.phantom(128, 128, 0), synth.radial_mask(128, 128, 0.3), 0.0, 0)
    sig, mus = synth.fixed_schedule(30)
    st, xs = O.run_trajectory(params, item, sig, mus, 30, keep=True)
    ps = [O.psnr(x.reshape(1, 128, 128), st["gt"].reshape(1, 128, 128)).item() for x in xs]
    np.testing.assert_allclose(ps, g["psnr"], rtol=0, atol=2e-3)
    np.testing.assert_allclose(st["x"].numpy(), g["x_final"], rtol=0, atol=5e-5)


def test_env64_kaiming_against_reference(golden_dir, meta):
    g = np.load(os.path.join(golden_dir, "ref_env64_kaiming.npz"))
    m = meta["cases"]["ref_env64_kaiming"]
    params = O.init_unet_params(1, "kaiming")
    assert params_sha(params) == m["params_sha"]
    item = synth.make_item(synth.phantom(64, 64, 3), synth.cartesian_mask(64, 64, 4, 3), 10.0, 3)
    for k, v in item.items():
        assert sha(v) == m["item_sha"][k], k
    st = O.reset(item)
    for k in range(5):
        a = act(0.0, 0.2 + 0.15 * k, (35.0 - 6 * k) / 255)
        if k == 0:
            pre = O.denoise(params, (st["z"] - st["u"]).real, a["sigma_d"], clamp=False)
            np.testing.assert_allclose(pre.numpy(), g["preclamp0"], rtol=0, atol=5e-6)
        st, _ = O.step(params, st, a)
        np.testing.assert_allclose(st["x"].numpy(), g["x_steps"][k], rtol=0, atol=1e-5)
    np.testing.assert_allclose(torch.view_as_real(st["u"]).numpy(), g["u"], rtol=0, atol=1e-5)


def test_unet_odd_sizes_against_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_unet_kaiming.npz"))
    params = O.init_unet_params(1, "kaiming")
    gen = torch.Generator().manual_seed(99)
    for (h, w) in ((64, 64), (48, 80), (36, 52)):
        inp = torch.rand(2, 2, h, w, generator=gen)
        np.testing.assert_allclose(O.unet_forward(params, inp).numpy(), g[f"out_{h}x{w}"], rtol=0, atol=1e-5)


def test_fft_psnr_against_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_fft_psnr.npz"))
    gen = torch.Generator().manual_seed(7)
    for (h, w) in ((32, 32), (64, 48), (128, 128), (30, 34)):
        zc = torch.complex(torch.randn(2, 1, h, w, generator=gen), torch.randn(2, 1, h, w, generator=gen))
        np.testing.assert_allclose(torch.view_as_real(O.centered_fft2(zc)).numpy(), g[f"fft_{h}x{w}"], atol=2e-6, rtol=0)
        np.testing.assert_allclose(torch.view_as_real(O.centered_ifft2(zc)).numpy(), g[f"ifft_{h}x{w}"], atol=2e-6, rtol=0)
    a = torch.rand(3, 40, 40, generator=gen) * 1.4 - 0.2
    b = torch.rand(3, 40, 40, generator=gen)
    np.testing.assert_allclose(O.psnr(a, b).numpy(), g["psnr"], rtol=0, atol=1e-5)


ANYSIZE_CASES = [(130, 130, "radial", 0.3, 0.0, 4), (136, 120, "cartesian", 4, 5.0, 5), (45, 51, "radial", 0.4, 0.0, 6)]


def anysize_item(H, W, kind, par, sn, seed):
    mask = synth.radial_mask(H, W, par) if kind == "radial" else synth.cartesian_mask(H, W, par, seed)
    return synth.make_item(synth.phantom(H, W, seed), mask, sn, seed)


@pytest.mark.parametrize("case", ANYSIZE_CASES, ids=lambda c: f"{c[0]}x{c[1]}")
def test_env_step_at_non_power_of_two_sizes_against_reference(golden_dir, case):
    """The reference's own ``PnPEnv.step`` at 130x130, 136x120 and 45x51 (oracle/make_golden_anysize.py): the oracle
    reproduces the saved states (generated bit-identical; a few ulp of slack for a different BLAS / thread count)."""
    g = np.load(os.path.join(golden_dir, "ref_env_anysize.npz"))
    H, W = case[:2]
    params = O.init_unet_params(0, "default")
    st = O.reset(anysize_item(*case))
    for k, (T, mu, sg) in enumerate(g["actions"]):
        st, done = O.step(params, st, act(T, mu, sg))
        assert done is False
        assert np.abs(st["x"].numpy() - g[f"x_steps_{H}x{W}"][k]).max() < 2e-5
    assert np.abs(torch.view_as_real(st["z"]).numpy() - g[f"z_{H}x{W}"]).max() < 2e-5
    assert np.abs(torch.view_as_real(st["u"]).numpy() - g[f"u_{H}x{W}"]).max() < 2e-5


def test_checkerboard_identity_used_by_the_cuda_kernels():
    """fft(w) == s * D . FFT2_ortho(D . w) with s = (-1)^((H+W)/2) (general even sizes)."""
    gen = torch.Generator().manual_seed(3)
    for (h, w) in ((32, 32), (64, 32), (30, 34), (30, 32)):
        zc = torch.complex(torch.randn(1, 1, h, w, generator=gen), torch.randn(1, 1, h, w, generator=gen))
        ii, jj = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        D = (1 - 2 * ((ii + jj) % 2)).float()
        s = -1.0 if ((h + w) // 2) % 2 else 1.0
        alt = s * D * torch.fft.fftn(D * zc, dim=(-2, -1), norm="ortho")
        assert (alt - O.centered_fft2(zc)).abs().max() < 1e-5
        alt_i = s * D * torch.fft.ifftn(D * zc, dim=(-2, -1), norm="ortho")
        assert (alt_i - O.centered_ifft2(zc)).abs().max() < 1e-5


def test_batched_step_equals_per_image_steps():
    """The oracle's batch generalisation == the reference's B=1 semantics applied image by image."""
    params = O.init_unet_params(2, "kaiming")
    batch = synth.make_batch(3, 32, 32, "cartesian", 4, 5.0, seed0=10)
    st = O.reset(batch)
    sig = torch.tensor([0.1, 0.05, 0.2])
    st, _ = O.step(params, st, {"T": torch.zeros(1), "mu": torch.tensor([0.4]), "sigma_d": sig})
    for b in range(3):
        one = {k: v[b:b + 1] for k, v in batch.items()}
        s1 = O.reset(one)
        s1, _ = O.step(params, s1, {"T": torch.zeros(1), "mu": torch.tensor([0.4]), "sigma_d": sig[b:b + 1]})
        assert (s1["x"] - st["x"][b:b + 1]).abs().max() < 1e-5
        assert (s1["u"] - st["u"][b:b + 1]).abs().max() < 1e-5


@pytest.mark.ref
def test_oracle_equals_live_reference():
    """Only where /root/reference exists: re-run the pin live (same checks as make_golden, one step)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ns = ref_shim.load()
    params = O.init_unet_params(5, "kaiming")
    item = synth.make_item(synth.phantom(128, 128, 2), synth.radial_mask(128, 128, 0.25), 5.0, 2)
    env = ns.PnPEnv(30, ref_shim.make_denoiser(ns, params), "cpu")
    rs = env.reset({k: torch.from_numpy(v.copy()) for k, v in item.items()}, "cpu")
    os_ = O.reset(item)
    a = act(0.0, 0.37, 0.12)
    rs, _ = env.step(rs, a)
    os_, _ = O.step(params, os_, a)
    assert (rs["x"] - os_["x"]).abs().max() < 1e-5
    assert (rs["u"] - os_["u"]).abs().max() < 1e-5

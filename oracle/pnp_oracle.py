"""CPU oracle for the PnP-ADMM CS-MRI environment step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module; the product path (``dt4image_restoration_b200``) never
does and has no CPU fallback.

What it is: a restatement, in plain PyTorch CPU fp32 ops, of the reference's hot path, generalised
from the reference's hard-wired batch-1 / 128x128 to any ``(B, H, W)``:

* ``centered_fft2 / centered_ifft2``  <- reference ``evaluation/utils/transformations.py:6-12, 14-19``
* ``unet_forward``                    <- reference ``evaluation/noise.py:75-98`` (ConvLayer/ConvBlock),
                                         ``:9-29`` (inconv/down), ``:32-61`` (up), ``:64-71`` (outconv),
                                         ``:119-133`` (UNet.forward)
* ``denoise``                         <- reference ``evaluation/noise.py:155-164``
* ``reset``                           <- reference ``evaluation/env.py:57-71`` (minus the literal 128s)
* ``step``                            <- reference ``evaluation/env.py:74-100``
* ``policy_ob``                       <- reference ``evaluation/env.py:103-109``
* ``psnr``                            <- reference ``evaluation/env.py:120-125``

Third-party arithmetic: everything numeric is PyTorch (``torch.fft``, ``conv2d``, ``max_pool2d``,
``interpolate``); the reference pins no version, so the oracle is pinned to this image's
torch 2.11.0 CPU kernels.

Parity pin: the reference ships NO tests, golden vectors, weights or data.
The pin is therefore "outputs of the reference itself run here": ``oracle/make_golden.py`` imports
``/root/reference`` (with the shims in ``oracle/ref_shim.py``), runs the reference's own
``PnPEnv.reset/step``, ``UNetDenoiser2D`` and ``torch_psnr`` on seeded inputs and random-init
weights, checks this restatement against them and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` re-checks the oracle against those vectors on every run.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# U-Net structure (reference evaluation/noise.py:101-113) as data: (block, cin, cout)
# --------------------------------------------------------------------------------------------
UNET_BLOCKS = (
    ("inc", 2, 32), ("down1", 32, 64), ("down2", 64, 128), ("down3", 128, 256), ("down4", 256, 512),
    ("up1", 512 + 256, 256), ("up2", 256 + 128, 128), ("up3", 128 + 64, 64), ("up4", 64 + 32, 32),
)


def _block_prefix(name: str) -> str:
    # state_dict key prefixes of the reference modules (noise.py:12, :21-24, :43)
    if name == "inc":
        return "inc.conv"
    if name.startswith("down"):
        return f"{name}.mpconv.1"
    return f"{name}.conv"


def unet_param_shapes() -> "OrderedDict[str, tuple]":
    """The 56 state_dict tensors of reference ``UNet(2, 1)``, in module registration order."""
    shapes = OrderedDict()
    for name, cin, cout in UNET_BLOCKS:
        p = _block_prefix(name)
        for i in range(3):
            ci = cin if i == 0 else cout
            shapes[f"{p}.conv-{i}.conv2d.weight"] = (cout, ci, 3, 3)
            shapes[f"{p}.conv-{i}.conv2d.bias"] = (cout,)
    shapes["outc.conv.weight"] = (1, 32, 1, 1)
    shapes["outc.conv.bias"] = (1,)
    return shapes


def init_unet_params(seed: int = 0, kind: str = "default", dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random-init weights keyed like the reference state_dict.

    ``default``  : the distribution of ``nn.Conv2d.reset_parameters`` (U(+-1/sqrt(fan_in)) for weight
                   and bias).  SURVEY.md §7.3(2b): the net is then nearly constant in its input.
    ``kaiming``  : signal-preserving N(0, 2/((1+0.2^2) fan_in)) weights, small biases, and a damped
                   1x1 output conv so the residual stays O(0.1); exercises every conv for real.
    """
    g = torch.Generator().manual_seed(int(seed) * 2654435761 % (2 ** 31) + 12345)
    out = OrderedDict()
    for key, shp in unet_param_shapes().items():
        if key.endswith("weight"):
            fan_in = shp[1] * shp[2] * shp[3]
            if kind == "default":
                b = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * b
            elif kind == "kaiming":
                std = math.sqrt(2.0 / ((1.0 + 0.2 ** 2) * fan_in))
                if key.startswith("outc"):
                    std *= 0.15
                t = torch.randn(shp, generator=g, dtype=torch.float64) * std
            else:
                raise ValueError(kind)
        else:
            w_shp = unet_param_shapes()[key[:-4] + "weight"]
            fan_in = w_shp[1] * w_shp[2] * w_shp[3]
            if kind == "default":
                b = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * b
            else:
                t = torch.randn(shp, generator=g, dtype=torch.float64) * 0.02
        out[key] = t.to(dtype)
    return out


# --------------------------------------------------------------------------------------------
# centred FFT (transformations.py:6-19)
# --------------------------------------------------------------------------------------------
def centered_fft2(img: torch.Tensor) -> torch.Tensor:
    t = torch.fft.ifftshift(img, dim=(-2, -1))
    t = torch.fft.fftn(t, dim=(-2, -1), norm="ortho")
    return torch.fft.fftshift(t, dim=(-2, -1))


def centered_ifft2(img: torch.Tensor) -> torch.Tensor:
    t = torch.fft.ifftshift(img, dim=(-2, -1))
    t = torch.fft.ifftn(t, dim=(-2, -1), norm="ortho")
    return torch.fft.fftshift(t, dim=(-2, -1))


# --------------------------------------------------------------------------------------------
# U-Net (noise.py:9-133)
# --------------------------------------------------------------------------------------------
def _conv_block(params, prefix, x, taps=None, tag=None):
    for i in range(3):
        x = F.conv2d(x, params[f"{prefix}.conv-{i}.conv2d.weight"], params[f"{prefix}.conv-{i}.conv2d.bias"],
                     stride=1, padding=1)
        x = F.leaky_relu(x, 0.2)
        if taps is not None:
            taps[f"{tag}.conv-{i}"] = x
    return x


def _up_block(params, name, x_low, x_skip, taps=None):
    x_up = F.interpolate(x_low, scale_factor=2, mode="bilinear", align_corners=True)
    dy = x_skip.shape[2] - x_up.shape[2]
    dx = x_skip.shape[3] - x_up.shape[3]
    x_up = F.pad(x_up, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    if taps is not None:
        taps[f"{name}.upsampled"] = x_up
    return _conv_block(params, _block_prefix(name), torch.cat([x_skip, x_up], dim=1), taps, name)


def unet_forward(params, inp: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """``inp`` [B,2,H,W] -> [B,1,H,W] = inp[:, :1] + residual (noise.py:119-133).

    ``taps``: optional dict that receives every post-activation feature map (for layer-wise parity).
    """
    x1 = _conv_block(params, _block_prefix("inc"), inp, taps, "inc")
    skips = [x1]
    x = x1
    for name in ("down1", "down2", "down3", "down4"):
        x = F.max_pool2d(x, 2)
        if taps is not None:
            taps[f"{name}.pooled"] = x
        x = _conv_block(params, _block_prefix(name), x, taps, name)
        skips.append(x)
    x = skips.pop()
    for name in ("up1", "up2", "up3", "up4"):
        x = _up_block(params, name, x, skips.pop(), taps)
    residual = F.conv2d(x, params["outc.conv.weight"], params["outc.conv.bias"])
    if taps is not None:
        taps["outc.residual"] = residual
    return inp[:, : residual.shape[1]] + residual


def denoise(params, x: torch.Tensor, sigma: torch.Tensor, clamp: bool = True, taps: dict | None = None) -> torch.Tensor:
    """noise.py:155-164: noise-level map concat, UNet, clamp to [0,1]."""
    N, C, H, W = x.shape
    sigma = torch.as_tensor(sigma, dtype=x.dtype).reshape(N, 1, 1, 1)
    noise_map = torch.ones(N, 1, H, W, dtype=x.dtype) * sigma
    out = unet_forward(params, torch.cat([x, noise_map], dim=1), taps)
    return torch.clamp(out, 0, 1) if clamp else out


# --------------------------------------------------------------------------------------------
# environment (env.py:57-125)
# --------------------------------------------------------------------------------------------
def reset(data: dict) -> "OrderedDict":
    """env.py:57-71 with the mask reshaped to ``[B,1,H,W]`` instead of ``[1,1,128,128]``."""
    x0 = torch.as_tensor(data["x0"], dtype=torch.float32).contiguous()
    y0 = torch.as_tensor(data["y0"], dtype=torch.float32).contiguous()
    B, _, H, W, _ = x0.shape
    x = torch.view_as_complex(x0)
    z = x.clone()
    u = torch.zeros_like(x)
    mask = torch.as_tensor(data["mask"]).reshape(-1, 1, H, W).contiguous().to(torch.bool)
    gt = torch.as_tensor(data["gt"], dtype=torch.float32)
    aty0 = torch.as_tensor(data["ATy0"])[..., 0]
    return OrderedDict({"x": x, "y0": torch.view_as_complex(y0), "z": z, "u": u, "mask": mask, "gt": gt,
                        "ATy0": aty0, "T": 0, "complex_y0": data["y0"]})


def prox_dual(x: torch.Tensor, u: torch.Tensor, y0: torch.Tensor, mask: torch.Tensor, mu) -> tuple:
    """env.py:87-93: z = ifft(blend(fft(x+u))), u' = u + x - z.  ``mu`` scalar or ``[B]``."""
    B = x.shape[0]
    z = centered_fft2(x + u)
    mu_t = torch.as_tensor(mu, dtype=torch.float32).reshape(-1)
    mu_b = mu_t.reshape(1, 1, 1, 1) if mu_t.numel() == 1 else mu_t.reshape(B, 1, 1, 1)
    temp = (mu_b * z.clone() + y0) / (1 + mu_b)
    m = mask if mask.shape[0] == B else mask.expand(B, -1, -1, -1)
    z = torch.where(m, temp, z)          # == z[mask] = temp[mask]
    z = centered_ifft2(z)
    return z, u + x - z


def step(params, states: "OrderedDict", action: dict, denoiser=None) -> tuple:
    """env.py:74-100.  Mutates and returns the same dict; ``x,z,u`` are re-bound to new tensors."""
    T, mu, sigma_d = action["T"], action["mu"], action["sigma_d"]
    if float(torch.as_tensor(T).reshape(-1)[0]) > 0.5:
        return states, True
    v = (states["z"] - states["u"]).real
    x = denoiser(v, sigma_d) if denoiser is not None else denoise(params, v, sigma_d)
    z, u = prox_dual(x, states["u"], states["y0"], states["mask"], mu)
    states["x"], states["z"], states["u"] = x, z, u
    states["T"] = states["T"] + 1 / 30
    return states, False


def policy_ob(states) -> torch.Tensor:
    """env.py:103-109, batch-generalised: ``[B, H*W]``."""
    x = states["x"].real
    return x.reshape(x.shape[0], -1)


def psnr(output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """env.py:120-125 -> ``[N,1]``."""
    N = output.shape[0]
    o = torch.clamp(output.real if output.is_complex() else output, 0, 1)
    mse = torch.mean(F.mse_loss(o.reshape(N, -1), gt.reshape(N, -1), reduction="none"), dim=1)
    return (10 * torch.log10((1 ** 2) / mse)).unsqueeze(1)


def run_trajectory(params, data: dict, sigmas, mus, n_iters: int = 30, keep: bool = False):
    """reset + ``n_iters`` steps with a fixed schedule (T forced to 0). Returns states (+ per-iter x)."""
    st = reset(data)
    xs = []
    B = st["x"].shape[0]
    for k in range(n_iters):
        act = {"T": torch.zeros(1), "mu": torch.as_tensor(mus[k]).reshape(-1),
               "sigma_d": torch.full((B,), float(sigmas[k]))}
        st, _ = step(params, st, act)
        if keep:
            xs.append(st["x"].clone())
    return (st, xs) if keep else st

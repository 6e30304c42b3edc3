"""Synthetic CS-MRI inputs (host side, NumPy, fixed seeds).

None of these generators exist in the reference: its eval items are ``.mat`` files that are not
shipped (reference ``dataset/datasets.py:148-168,184-207`` only *loads* ``x0,y0,mask,ATy0,gt``).
SURVEY.md §8d defines the generators; they run once on the host and the *same arrays* are fed to
the CPU oracle and to the CUDA path, so sampling masks and indexing are bit-exact by construction.

Conventions follow the reference item layout as it leaves ``DataLoader(batch_size=1)``
(``evaluation/eval.py:226-232``): ``x0, y0, ATy0`` are float32 ``[B,1,H,W,2]`` (real, imag last),
``mask`` is ``[B,H,W]`` (any numeric dtype, converted to bool in ``PnPEnv.reset``,
reference ``evaluation/env.py:64``) and ``gt`` is float32 ``[B,1,H,W]``.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "phantom",
    "radial_mask",
    "cartesian_mask",
    "centered_fft2",
    "centered_ifft2",
    "make_item",
    "make_batch",
    "fixed_schedule",
]

# (intensity, a, b, x0, y0, phi_deg) -- a Shepp-Logan-style head: skull, brain, ventricles, lesions.
_ELLIPSES = (
    (1.00, 0.690, 0.920, 0.000, 0.000, 0.0),
    (-0.80, 0.6624, 0.8740, 0.000, -0.0184, 0.0),
    (-0.20, 0.1100, 0.3100, 0.220, 0.000, -18.0),
    (-0.20, 0.1600, 0.4100, -0.220, 0.000, 18.0),
    (0.10, 0.2100, 0.2500, 0.000, 0.350, 0.0),
    (0.10, 0.0460, 0.0460, 0.000, 0.100, 0.0),
    (0.10, 0.0460, 0.0460, 0.000, -0.100, 0.0),
    (0.10, 0.0460, 0.0230, -0.080, -0.605, 0.0),
    (0.10, 0.0230, 0.0230, 0.000, -0.606, 0.0),
    (0.10, 0.0230, 0.0460, 0.060, -0.605, 0.0),
)


def phantom(H: int, W: int, seed: int = 0) -> np.ndarray:
    """Brain-like phantom, float32 ``[H,W]`` in [0,1]. ``seed`` perturbs every ellipse by +-5 %."""
    rng = np.random.default_rng(1000003 * int(seed) + 17)
    ys = (np.arange(H, dtype=np.float64) + 0.5) / H * 2.0 - 1.0
    xs = (np.arange(W, dtype=np.float64) + 0.5) / W * 2.0 - 1.0
    Y, X = np.meshgrid(ys, xs, indexing="ij")
    img = np.zeros((H, W), dtype=np.float64)
    for (amp, a, b, x0, y0, phi) in _ELLIPSES:
        j = 1.0 + 0.05 * rng.uniform(-1.0, 1.0, size=6)
        amp_, a_, b_ = amp * j[0], a * j[1], b * j[2]
        x0_, y0_ = x0 + 0.4 * (j[3] - 1.0), y0 + 0.4 * (j[4] - 1.0)  # centre shift within +-0.02
        th = np.deg2rad(phi * j[5])
        c, s = np.cos(th), np.sin(th)
        xr = (X - x0_) * c + (Y - y0_) * s
        yr = -(X - x0_) * s + (Y - y0_) * c
        img += amp_ * ((xr / a_) ** 2 + (yr / b_) ** 2 <= 1.0)
    img = np.clip(img, 0.0, None)
    img /= max(img.max(), 1e-12)
    return img.astype(np.float32)


def radial_mask(H: int, W: int, frac: float) -> np.ndarray:
    """Golden-angle radial lines through the k-space centre (centred convention), uint8 ``[H,W]``.

    Lines are added one at a time until the sampled fraction reaches ``frac``.
    """
    mask = np.zeros((H, W), dtype=np.uint8)
    cy, cx = H // 2, W // 2
    R = int(np.ceil(np.hypot(H, W) / 2)) + 1
    t = np.arange(-R, R + 1, dtype=np.float64)
    golden = np.pi * (np.sqrt(5.0) - 1.0) / 2.0
    target = frac * H * W
    n = 0
    while mask.sum() < target and n < 8 * max(H, W):
        ang = n * golden
        yy = np.rint(cy + t * np.sin(ang)).astype(np.int64)
        xx = np.rint(cx + t * np.cos(ang)).astype(np.int64)
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        mask[yy[ok], xx[ok]] = 1
        n += 1
    return mask


def cartesian_mask(H: int, W: int, accel: int, seed: int = 0, acs_frac: float = 0.08) -> np.ndarray:
    """Cartesian ``accel``x undersampling: full columns, centred ACS band + seeded random columns."""
    rng = np.random.default_rng(7919 * int(seed) + int(accel))
    n_keep = max(1, W // int(accel))
    n_acs = min(n_keep, max(1, int(round(acs_frac * W))))
    cols = np.zeros(W, dtype=bool)
    c0 = W // 2 - n_acs // 2
    cols[c0:c0 + n_acs] = True
    rest = np.flatnonzero(~cols)
    extra = n_keep - n_acs
    if extra > 0:
        cols[rng.choice(rest, size=extra, replace=False)] = True
    return np.repeat(cols[None, :], H, axis=0).astype(np.uint8)


def centered_fft2(x: np.ndarray) -> np.ndarray:
    """NumPy twin of reference ``fft`` (``evaluation/utils/transformations.py:6-12``)."""
    return np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(x, axes=(-2, -1)), norm="ortho"), axes=(-2, -1))


def centered_ifft2(x: np.ndarray) -> np.ndarray:
    """NumPy twin of reference ``ifft`` (``evaluation/utils/transformations.py:14-19``)."""
    return np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(x, axes=(-2, -1)), norm="ortho"), axes=(-2, -1))


def _ri(z: np.ndarray) -> np.ndarray:
    return np.stack([z.real, z.imag], axis=-1).astype(np.float32)


def make_item(gt: np.ndarray, mask: np.ndarray, sigma_n: float = 0.0, seed: int = 0) -> dict:
    """One eval item with the reference's keys/shapes for B=1 (see module docstring).

    ``y0 = mask * (fft(gt) + sigma_n/255 * (n_r + i n_i))``; ``x0 = ifft(y0)`` with real and
    imaginary parts clipped at 0 from below (reference ``dataset/datasets.py:160,199`` clips the
    whole ``x0`` array); ``ATy0 = ifft(y0)`` unclipped.
    """
    H, W = gt.shape
    rng = np.random.default_rng(104729 * int(seed) + 3)
    k = centered_fft2(gt.astype(np.float64))
    if sigma_n > 0:
        k = k + (sigma_n / 255.0) * (rng.standard_normal((H, W)) + 1j * rng.standard_normal((H, W)))
    y0 = k * mask.astype(np.float64)
    aty0 = centered_ifft2(y0)
    x0 = np.clip(_ri(aty0), 0.0, None)
    return {
        "x0": x0.reshape(1, 1, H, W, 2),
        "y0": _ri(y0).reshape(1, 1, H, W, 2),
        "ATy0": _ri(aty0).reshape(1, 1, H, W, 2),
        "mask": mask.reshape(1, H, W).copy(),
        "gt": gt.astype(np.float32).reshape(1, 1, H, W),
    }


def make_batch(B: int, H: int, W: int, mask_kind: str = "radial", mask_param: float = 0.3,
               sigma_n: float = 0.0, seed0: int = 0, shared_mask: bool = False) -> dict:
    """Batch of ``B`` items concatenated on dim 0 (image index = ``seed0 + b``)."""
    items = []
    for b in range(B):
        s = seed0 + b
        gt = phantom(H, W, s)
        if mask_kind == "radial":
            m = radial_mask(H, W, mask_param)
        elif mask_kind == "cartesian":
            m = cartesian_mask(H, W, int(mask_param), 0 if shared_mask else s)
        else:
            raise ValueError(f"unknown mask kind {mask_kind!r}")
        items.append(make_item(gt, m, sigma_n, s))
    return {k: np.concatenate([it[k] for it in items], axis=0) for k in items[0]}


def fixed_schedule(n_iters: int = 30):
    """Config-1 schedule: sigma_d geometric 50->5 (/255), mu linear 0.1->1.0."""
    k = np.arange(n_iters, dtype=np.float64) / max(n_iters - 1, 1)
    sigma = (50.0 * (5.0 / 50.0) ** k) / 255.0
    mu = 0.1 + 0.9 * k
    return sigma.astype(np.float32), mu.astype(np.float32)

"""On-disk formats of the reference, either side of the hot path.

* evaluation items: one MATLAB ``.mat`` file per image with the arrays ``x0, y0, mask, ATy0, gt`` that the
  reference ``EvaluationDataset`` / ``EvaluationOptimalDataset`` read with ``scipy.io.loadmat``
  (reference ``dataset/datasets.py:148-168,184-207``); ``x0`` is clipped at 0 from below exactly as there
  (``:160,199``) and the task ("4_10" = 4x acceleration, noise level 10) is parsed from the file name the same
  way (``extract_task``, ``:13-16``).  The loaders return the item dict that ``PnPEnv.reset`` / ``PnPEngine.reset``
  take, for one file or stacked for a whole directory, at any image size (the reference hard-codes 128x128).
* denoiser checkpoints: the ``state_dict`` of the reference ``UNet(2, 1)`` saved with ``torch.save``
  (``evaluation/noise.py:147-148``) - ``UNetDenoiser2D(ckpt_path=...)`` reads it directly; ``save_unet_checkpoint``
  writes one from a flat parameter vector or a state_dict so round trips can be tested without the (unshipped)
  pretrained file.

Host-side only (NumPy / SciPy); nothing here touches the GPU.
"""
from __future__ import annotations

import os
import re

import numpy as np

ITEM_KEYS = ("x0", "y0", "mask", "ATy0", "gt")
# task vocabulary of the reference's evaluation datasets (dataset/datasets.py:171-172)
TASKS = ["2x_5", "2x_10", "2x_15", "4x_5", "4x_10", "4x_15", "8x_5", "8x_10", "8x_15"]
TASK_TOKENIZER = {t: i for i, t in enumerate(TASKS)}
MIN_RTG, MAX_RTG = -1.08, 16.6          # dataset/datasets.py:174-175


def extract_task(filename: str) -> str:
    """'..._4_10.mat' -> '4_10' (reference ``extract_task``); raises ValueError when the name carries no task."""
    m = re.search(r"\d+_\d+", os.path.basename(filename))
    if m is None:
        raise ValueError(f"no '<accel>_<noise>' task in file name {filename!r}")
    return m.group()


def task_token(filename: str) -> int:
    """Token the decision transformer is conditioned on (reference ``datasets.py:188-191``)."""
    t = extract_task(filename)
    a, n = t.split("_")
    key = f"{a}x_{n}"
    if key not in TASK_TOKENIZER:
        raise KeyError(f"unknown task {key!r}")
    return TASK_TOKENIZER[key]


def normalised_rtg(rtg_target: float) -> float:
    """Return-to-go conditioning value (reference ``datasets.py:204``)."""
    return (rtg_target - MIN_RTG) / (MAX_RTG - MIN_RTG)


def _canon(item: dict) -> dict:
    """Bring one item to the layout ``reset`` expects: x0/y0/ATy0 [1,1,H,W,2] f32, mask [1,H,W], gt [1,1,H,W] f32."""
    gt = np.asarray(item["gt"], dtype=np.float32)
    H, W = gt.shape[-2], gt.shape[-1]
    out = {}
    for k in ("x0", "y0", "ATy0"):
        a = np.asarray(item[k])
        if np.iscomplexobj(a):
            a = np.stack([a.real, a.imag], axis=-1)
        out[k] = np.ascontiguousarray(a, dtype=np.float32).reshape(1, 1, H, W, 2)
    out["x0"] = np.clip(out["x0"], 0.0, None)                     # datasets.py:160,199
    out["mask"] = np.ascontiguousarray(np.asarray(item["mask"]).reshape(1, H, W))
    out["gt"] = gt.reshape(1, 1, H, W)
    return out


def load_eval_item(path: str) -> dict:
    """Read one reference ``.mat`` evaluation item."""
    from scipy.io import loadmat
    mat = loadmat(path)
    missing = [k for k in ITEM_KEYS if k not in mat]
    if missing:
        raise KeyError(f"{path}: missing arrays {missing}")
    return _canon({k: mat[k] for k in ITEM_KEYS})


def save_eval_item(path: str, item: dict) -> None:
    """Write an item in the reference's ``.mat`` layout (arrays squeezed to [H,W,2] / [H,W] like its data files)."""
    from scipy.io import savemat
    c = _canon(item)
    H, W = c["gt"].shape[-2:]
    savemat(path, {"x0": c["x0"].reshape(H, W, 2), "y0": c["y0"].reshape(H, W, 2), "ATy0": c["ATy0"].reshape(H, W, 2),
                   "mask": c["mask"].reshape(H, W), "gt": c["gt"].reshape(H, W)})


def list_eval_items(data_dir: str) -> list[str]:
    """Sorted ``.mat`` files of a directory (reference ``datasets.py:145-146``)."""
    return [os.path.join(data_dir, f) for f in sorted(os.listdir(data_dir)) if f.endswith(".mat")]


def load_eval_batch(paths) -> dict:
    """Stack items on dim 0 -> the batched dict of ``PnPEngine.reset`` (all items must share one image size)."""
    items = [load_eval_item(p) for p in paths]
    if not items:
        raise ValueError("no evaluation items")
    shapes = {it["gt"].shape for it in items}
    if len(shapes) != 1:
        raise ValueError(f"items of different sizes cannot be batched: {sorted(shapes)}")
    return {k: np.concatenate([it[k] for it in items], axis=0) for k in ITEM_KEYS}


def save_unet_checkpoint(path: str, params) -> None:
    """``torch.save`` a reference-format U-Net state_dict; ``params`` = state_dict or the flat fp32 vector."""
    import torch
    from . import ops
    if not isinstance(params, dict):
        params = ops.unflatten_state_dict(torch.as_tensor(params))
    torch.save({k: v.detach().cpu() for k, v in params.items()}, path)

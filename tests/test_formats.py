"""On-disk formats: reference .mat evaluation items and U-Net checkpoints. CPU only."""
import os

import numpy as np
import pytest
import torch

from dt4image_restoration_b200 import formats, ops, synth
from oracle import pnp_oracle as O


def test_mat_item_round_trip(tmp_path):
    item = synth.make_item(synth.phantom(64, 64, 3), synth.cartesian_mask(64, 64, 4, 3), 10.0, 3)
    p = os.path.join(tmp_path, "img_4_10.mat")
    formats.save_eval_item(p, item)
    back = formats.load_eval_item(p)
    for k in formats.ITEM_KEYS:
        assert back[k].shape == item[k].shape, k
        assert np.array_equal(back[k].astype(np.float64), np.asarray(item[k]).astype(np.float64)), k
    assert formats.extract_task(p) == "4_10"
    assert formats.task_token(p) == formats.TASK_TOKENIZER["4x_10"] == 4
    assert abs(formats.normalised_rtg(10) - (10 + 1.08) / (16.6 + 1.08)) < 1e-12     # reference datasets.py:204


def test_x0_is_clipped_like_the_reference(tmp_path):
    item = synth.make_item(synth.phantom(32, 32, 1), synth.radial_mask(32, 32, 0.3), 0.0, 1)
    item = dict(item)
    item["x0"] = np.asarray(item["ATy0"]).copy()          # unclipped, has negative entries
    assert (item["x0"] < 0).any()
    p = os.path.join(tmp_path, "a_2_5.mat")
    formats.save_eval_item(p, item)
    assert (formats.load_eval_item(p)["x0"] >= 0).all()    # np.clip(x0, 0, None), datasets.py:160,199


def test_directory_batch_feeds_the_oracle_reset(tmp_path):
    for i in range(3):
        it = synth.make_item(synth.phantom(32, 32, i), synth.radial_mask(32, 32, 0.3), 0.0, i)
        formats.save_eval_item(os.path.join(tmp_path, f"im{i}_4_5.mat"), it)
    open(os.path.join(tmp_path, "notes.txt"), "w").write("ignored")
    paths = formats.list_eval_items(str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["im0_4_5.mat", "im1_4_5.mat", "im2_4_5.mat"]
    batch = formats.load_eval_batch(paths)
    assert batch["x0"].shape == (3, 1, 32, 32, 2) and batch["mask"].shape == (3, 32, 32)
    st = O.reset({k: torch.from_numpy(v) for k, v in batch.items()})
    assert st["x"].shape == (3, 1, 32, 32) and st["mask"].dtype == torch.bool
    ref = synth.make_batch(3, 32, 32, "radial", 0.3)
    assert np.array_equal(batch["y0"], ref["y0"])


def test_mixed_sizes_and_missing_keys_are_rejected(tmp_path):
    formats.save_eval_item(os.path.join(tmp_path, "a_4_5.mat"),
                           synth.make_item(synth.phantom(32, 32, 0), synth.radial_mask(32, 32, 0.3)))
    formats.save_eval_item(os.path.join(tmp_path, "b_4_5.mat"),
                           synth.make_item(synth.phantom(64, 64, 0), synth.radial_mask(64, 64, 0.3)))
    with pytest.raises(ValueError):
        formats.load_eval_batch(formats.list_eval_items(str(tmp_path)))
    from scipy.io import savemat
    bad = os.path.join(tmp_path, "bad_4_5.mat")
    savemat(bad, {"x0": np.zeros((4, 4, 2))})
    with pytest.raises(KeyError):
        formats.load_eval_item(bad)
    with pytest.raises(ValueError):
        formats.extract_task("no_task_here.mat")


def test_unet_checkpoint_round_trip(tmp_path):
    sd = O.init_unet_params(1, "default")
    flat = ops.flatten_state_dict(sd)
    assert [k for k, _ in ops.unet_state_dict_shapes()] == list(sd.keys())
    p = os.path.join(tmp_path, "unet-nm.pt")
    formats.save_unet_checkpoint(p, flat)
    back = torch.load(p, map_location="cpu")
    assert list(back.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
    assert torch.equal(ops.flatten_state_dict(back), flat)

"""N>1 host logic on CPU: world_size-2 gloo processes exercise sharding + the reward all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dt4image_restoration_b200 import dist as pd


def test_shard_range_partitions_exactly():
    for n in (1, 2, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [pd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = {"x0": np.arange(n_units * 2, dtype=np.float32).reshape(n_units, 2), "T": 0}
        mine = pd.shard_batch(data, rank, world)
        lo, hi = pd.shard_range(n_units, rank, world)
        assert mine["x0"].shape[0] == hi - lo and mine["T"] == 0
        # "reward" of unit i is a known function of its global index
        local = torch.tensor([float((i * 37) % 11) - 0.01 * i for i in range(lo, hi)])
        allr = pd.gather_rewards(local, n_units)
        expect = torch.tensor([float((i * 37) % 11) - 0.01 * i for i in range(n_units)])
        assert torch.equal(allr, expect)
        idx, val = pd.global_argmax(local, n_units)
        assert idx == int(torch.argmax(expect)) and abs(val - expect.max().item()) < 1e-7
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [8, 7])
def test_reward_all_gather_world2_gloo(n_units):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

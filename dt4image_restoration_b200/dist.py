"""Multi-GPU plumbing: one process per GPU, contiguous shards of independent units, reward all-gather.

Images / trajectories / MCTS candidate expansions never interact inside ``PnPEnv.step`` (reference
``evaluation/env.py:85-93`` is per image), so the data path needs no collective.  The only exchange is the
all-gather of per-unit rewards that a global selection needs (UCB / argmax over candidates, reference
``evaluation/mcts.py:74-88,34-38``): ``B_local`` fp32 per rank, latency-bound, done with NCCL
(``torch.distributed``; ``gloo`` in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: rank r owns ``[lo, hi)``; the first ``n % world`` ranks get one extra unit."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(data: dict, rank: int, world: int) -> dict:
    """Slice every per-image entry of an item dict (``x0, y0, mask, ATy0, gt``) to this rank's shard."""
    n = data["x0"].shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: (v[lo:hi] if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in data.items()}


def gather_rewards(local: torch.Tensor, n_units: int | None = None) -> torch.Tensor:
    """All-gather per-unit rewards into global unit order.  Handles ragged shards by padding to the max shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.clone()
    world, rank = dist.get_world_size(), dist.get_rank()
    local = local.reshape(-1).contiguous()
    if n_units is None:
        n_units = local.numel() * world
    sizes = [shard_range(n_units, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    buf = torch.full((mx,), float("nan"), dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(world * mx, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def global_argmax(local: torch.Tensor, n_units: int | None = None) -> tuple[int, float]:
    """Index (global unit numbering) and value of the best reward over all ranks."""
    allr = gather_rewards(local, n_units)
    i = int(torch.argmax(allr).item())
    return i, float(allr[i].item())

"""Multi-GPU plumbing: one process per GPU, contiguous shards of independent units, reward all-gather.

Images / trajectories / MCTS candidate expansions never interact inside ``PnPEnv.step`` (reference
``evaluation/env.py:85-93`` is per image), so the data path needs no collective.  The only exchange is the
all-gather of per-unit rewards that a global selection needs (UCB / argmax over candidates, reference
``evaluation/mcts.py:74-88,34-38``): ``B_local`` fp32 per rank, latency-bound.  Two implementations:

* ``gather_rewards``: NCCL all-gather (``torch.distributed``; ``gloo`` in the CPU tests) of rewards computed before;
* ``PeerRewardGather``: reward kernel and all-gather fused - the PSNR kernel stores every reward straight into all ranks'
  copies of a symmetric buffer over NVLink peer mappings and finishes when all ranks have delivered
  (``pnp_psnr_allgather``); no separate collective launch.
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


class PeerRewardGather:
    """Per-image PSNR rewards of all ranks in one kernel (``pnp_psnr_allgather``, csrc/psnr.cu).

    ``slot``: the largest per-rank batch.  The symmetric buffer comes from ``torch.distributed._symmetric_memory``
    (cuMem allocations exchanged between the processes of one node and mapped over NVLink); ``local_only=True`` (or a
    world of one) uses an ordinary device buffer.  Every rank must call ``psnr_allgather`` the same number of times.
    The returned view is valid until THIS rank's next call: results are double-buffered by call parity so that a peer that
    runs ahead (call n + 1) cannot overwrite what this rank still reads from call n, but a peer may start call n + 2 - which
    reuses this parity - as soon as this rank has launched call n + 1.  Clone the view to keep it longer.
    A rank that does not arrive within 10 s sets a device flag (``timed_out()``); ``psnr_allgather(check=True)`` folds the
    flag into the result (all-NaN rewards) so that a stale gather cannot feed an argmax unnoticed."""

    def __init__(self, slot: int, device, group=None, local_only: bool = False):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        dev = torch.device(device)
        multi = dist.is_available() and dist.is_initialized() and not local_only
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        if self.world > 8:
            raise RuntimeError("PeerRewardGather: at most 8 ranks (one NVSwitch node)")
        self.slot = int(slot)
        n = 2 * self.world * self.slot
        self.flag_word = n
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            self.buf = symm.empty(n + 32, dtype=torch.float32, device=dev)
            self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            ptrs = [int(p) for p in self.handle.buffer_ptrs]
        else:
            self.buf = torch.empty(n + 32, dtype=torch.float32, device=dev)
            self.handle = None
            ptrs = [self.buf.data_ptr()]
        self.buf.zero_()
        torch.cuda.synchronize(dev)
        if self.world > 1:
            dist.barrier(group)                      # nobody signals before every copy is zeroed
        self.ptrs = (C.c_ulonglong * 8)(*ptrs, *([0] * (8 - len(ptrs))))
        self.local = torch.zeros(2, dtype=torch.int32, device=dev)       # [finished CTAs, timeout flag]
        self.calls = 0
        self.count = 0

    def psnr_allgather(self, x: torch.Tensor, gt: torch.Tensor, check: bool = False) -> torch.Tensor:
        """``x, gt``: fp32 ``[B,1,H,W]`` (or ``[B,H,W]``) on this rank -> view ``[world, slot]`` of all ranks' rewards
        (row r, first ``B_r`` entries)."""
        C, lib = self._C, self._lib
        B = x.shape[0]
        HW = x[0].numel()
        if not (x.is_cuda and gt.is_cuda and x.dtype == torch.float32 and gt.dtype == torch.float32
                and x.is_contiguous() and gt.is_contiguous()) or B > self.slot:
            raise RuntimeError("psnr_allgather: contiguous fp32 CUDA tensors with B <= slot expected")
        stride = HW if gt.numel() == B * HW else 0
        if stride == 0 and gt.numel() != HW:
            raise RuntimeError("psnr_allgather: gt must hold B or 1 images")
        calls, count = self.calls + 1, self.count + B      # committed only after a successful launch: a failed call must
        parity = calls & 1                                   # not leave this rank's targets out of step with its peers'
        lib.check(lib.lib().pnp_psnr_allgather(
            x.data_ptr(), gt.data_ptr(), stride, None, C.cast(self.ptrs, C.c_void_p), self.rank, self.world, self.slot,
            parity, self.flag_word, self.local.data_ptr(), C.c_uint(count & 0xFFFFFFFF),
            C.c_uint((self.world * calls) & 0xFFFFFFFF), self.local[1:].data_ptr(), B, HW, lib.stream_ptr()),
            "pnp_psnr_allgather")
        self.calls, self.count = calls, count
        w = self.world * self.slot
        out = self.buf[parity * w:(parity + 1) * w].view(self.world, self.slot)
        if check:                                            # device-side: NaN everywhere if any call timed out (no host sync)
            out = torch.where(self.local[1] != 0, torch.full_like(out, float("nan")), out)
        return out

    def timed_out(self) -> bool:
        """Host check (synchronises): did any call give up waiting for a rank?"""
        return bool(self.local[1].item())


def make_peer_gather(slot: int, device, group=None) -> "PeerRewardGather | None":
    """``PeerRewardGather`` if every rank can set up the symmetric buffer, else ``None`` on ALL ranks (callers then use the
    NCCL ``gather_rewards``).  Collective: the outcome is agreed with a MIN all-reduce."""
    pg = None
    try:
        pg = PeerRewardGather(slot, device, group)
    except Exception:                      # no peer access / symmetric memory unavailable in this process
        pg = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        ok = torch.tensor([1 if pg is not None else 0], dtype=torch.int32, device=torch.device(device))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            pg = None
    return pg


def global_argmax(local: torch.Tensor, n_units: int | None = None) -> tuple[int, float]:
    """Index (global unit numbering) and value of the best reward over all ranks."""
    allr = gather_rewards(local, n_units)
    i = int(torch.argmax(allr).item())
    return i, float(allr[i].item())

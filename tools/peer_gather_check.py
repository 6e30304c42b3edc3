#!/usr/bin/env python
"""Multi-GPU check of the fused reward all-gather (pnp_psnr_allgather) against the NCCL all-gather, plus latency.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/peer_gather_check.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops, dist as pdist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S = 64, 256
pg = pdist.PeerRewardGather(B, dev)
ok = True
for it in range(6):                       # several calls: both parities, ragged batch on the last rank
    Bl = B - 3 if (rank == world - 1 and it % 2) else B
    g = torch.Generator(device=dev).manual_seed(100 * it + rank)
    x = torch.rand(Bl, 1, S, S, device=dev, generator=g)
    gt = torch.rand(Bl, 1, S, S, device=dev, generator=g)
    allr = pg.psnr_allgather(x, gt).clone()
    ref_local = ops.psnr(x, gt).reshape(-1)
    pad = torch.full((B,), float("nan"), device=dev); pad[:Bl] = ref_local
    ref = torch.empty(world * B, device=dev)
    dist.all_gather_into_tensor(ref, pad)
    ref = ref.view(world, B)
    m = ~torch.isnan(ref)
    ok &= bool(torch.equal(allr[m], ref[m]))
ok &= not pg.timed_out()
# latency: fused kernel vs psnr kernel + NCCL all-gather
x = torch.rand(B, 1, S, S, device=dev); gt = torch.rand(B, 1, S, S, device=dev)
out = torch.empty(world * B, device=dev)
def t(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_fused = t(lambda: pg.psnr_allgather(x, gt))
t_nccl = t(lambda: dist.all_gather_into_tensor(out, ops.psnr(x, gt).reshape(-1)))
t_local = t(lambda: ops.psnr(x, gt))
ok &= not pg.timed_out()
res = torch.tensor([1.0 if ok else 0.0, t_fused, t_nccl, t_local], device=dev)
dist.all_reduce(res[:1], op=dist.ReduceOp.MIN); dist.all_reduce(res[1:], op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"peer gather world={world} B={B} {S}x{S}: match_nccl={bool(res[0].item())}  fused {res[1].item():.1f} us  "
          f"psnr+NCCL all-gather {res[2].item():.1f} us  psnr alone {res[3].item():.1f} us")
dist.destroy_process_group()
sys.exit(0 if res[0].item() == 1.0 else 1)

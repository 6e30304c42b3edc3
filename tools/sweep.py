#!/usr/bin/env python
"""BASELINE config 5: throughput sweep over the batch size (1 .. 4096 per GPU) at 128x128 and 256x256, the FFT-prox kernels
and the U-Net denoiser timed in isolation, on 1 / 2 / 4 / 8 GPUs.

    python tools/sweep.py                                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py   # N GPUs, weak scaling

The path shards by image with no exchange, so every rank runs the same per-GPU batch; a cell's time is the MAX over ranks
(CUDA events, barrier on both sides) and the table reports the aggregate over all ranks.  The denoiser's workspace is
bounded (micro-batched plans), so B = 4096 at 256x256 runs.
"""
import argparse, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200 import ops, _lib
from dt4image_restoration_b200.noise import UNetDenoiser2D, random_init_state_dict

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=""); ap.add_argument("--sizes", default="128,256")
ap.add_argument("--max-batch", type=int, default=4096)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
GF = {128: 9.684, 256: 38.734, 512: 154.938}
den = UNetDenoiser2D(state_dict=random_init_state_dict(0, "default")).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timeit(fn, it):
    for _ in range(2): fn()
    barrier(); e0.record()
    for _ in range(it): fn()
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1) / it * 1e-3], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def say(*args):
    if rank == 0:
        print(*args, flush=True)


rows = []
say(f"{world} GPU(s); per-GPU batch in the first column, rates are the aggregate over all GPUs; HBM peak {peaks['hbm_gbs']:.0f} GB/s, "
    f"bf16 peak {peaks['bf16_tflops']:.0f} (burst) / {peaks['bf16_tflops_sustained']:.0f} (sustained) TFLOP/s per GPU")
for S in [int(s) for s in a.sizes.split(",")]:
    say(f"--- {S}x{S} ---")
    say(f"{'batch':>6s} | {'prox Cartesian us':>18s} {'Mimg/s':>8s} {'HBM frac':>8s} | {'prox radial us':>15s} {'Mimg/s':>8s} {'HBM frac':>8s} | "
        f"{'U-Net ms':>9s} {'kimg/s':>8s} {'TFLOP/s':>8s} {'of burst':>8s} {'ws GB':>6s} {'micro-b':>7s}")
    B = 1
    while B <= a.max_batch:
        g = torch.Generator(device=dev).manual_seed(B)
        x = torch.rand(B, 1, S, S, device=dev, generator=g)
        u = torch.complex(torch.randn(B, 1, S, S, device=dev, generator=g), torch.randn(B, 1, S, S, device=dev, generator=g)) * 0.1
        y0 = torch.complex(torch.randn(B, 1, S, S, device=dev, generator=g), torch.randn(B, 1, S, S, device=dev, generator=g))
        mu = torch.full((B,), 0.5, device=dev)
        out = (torch.empty_like(u), torch.empty_like(u), torch.empty_like(x))
        res = {"size": S, "batch_per_gpu": B, "gpus": world}
        for kind in ("cartesian", "radial"):
            if kind == "cartesian":
                mask = (torch.rand(B, 1, 1, S, device=dev, generator=g) < 0.25).expand(B, 1, S, S).contiguous()
            else:
                mask = torch.rand(B, 1, S, S, device=dev, generator=g) < 0.25
            prep = ops.ProxPrepared(y0, mask)
            _ = prep.column_only                              # synchronise once: the host-side mask-kind hint is known
            t = timeit(lambda: prep.prox_dual(x, u, mu, out=out), 20 if B <= 256 else 5)
            res[f"prox_{kind}_us"] = t * 1e6
            res[f"prox_{kind}_hbm_frac"] = 37.0 * B * S * S / t / 1e9 / peaks["hbm_gbs"]
            del prep
        del y0, out
        l = _lib.lib()
        plan = den.plan(B, S, S)
        res["unet_ws_gb"] = l.pnp_unet_workspace_bytes(B, S, S) / 1e9
        res["unet_micro_batch"] = l.pnp_unet_micro_batch(plan.handle)
        v = torch.rand(B, 1, S, S, device=dev); sg = torch.full((B,), 0.1, device=dev)
        t = timeit(lambda: plan.forward(v, sg, out=x), 10 if B <= 64 else 3)
        res["unet_ms"] = t * 1e3
        res["unet_tflops_per_gpu"] = GF[S] * B / t / 1e3
        den._plans.clear(); del plan, v, x, u
        torch.cuda.empty_cache()
        say(f"{B:6d} | {res['prox_cartesian_us']:18.1f} {world * B / res['prox_cartesian_us']:8.3f} {res['prox_cartesian_hbm_frac']:8.3f} | "
            f"{res['prox_radial_us']:15.1f} {world * B / res['prox_radial_us']:8.3f} {res['prox_radial_hbm_frac']:8.3f} | "
            f"{res['unet_ms']:9.3f} {world * B / res['unet_ms']:8.2f} {world * res['unet_tflops_per_gpu']:8.1f} "
            f"{res['unet_tflops_per_gpu'] / peaks['bf16_tflops']:8.3f} {res['unet_ws_gb']:6.2f} {res['unet_micro_batch']:7d}")
        rows.append(res)
        B *= 4 if B >= 16 else 2
if a.out and rank == 0:
    json.dump(rows, open(a.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()

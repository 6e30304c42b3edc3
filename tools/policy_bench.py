#!/usr/bin/env python
"""Time of one policy iteration: the fused kernel (pnp_policy_step) vs the two PyTorch forwards it replaces."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dt4image_restoration_b200.policy import DecisionTransformer, FusedPolicy
B, K = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 6
dev = "cuda"
pol = DecisionTransformer().to(dev).eval()
w_rtg = torch.rand(B, K, 1, device=dev); w_emb = torch.randn(B, K, 128, device=dev); w_act = torch.rand(B, K, 3, device=dev)
w_ts = torch.randint(0, 30, (B, K, 1), device=dev); w_task = torch.full((B, K), 3, device=dev)
pos = torch.tensor([5], device=dev); ao = torch.zeros(B, 3, device=dev); ro = torch.zeros(B, 1, 1, device=dev)
fp = FusedPolicy(pol)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def two_forwards():
    pa, ad = pol.forward_tokens(w_rtg, w_emb, w_ts, w_task, w_act, eval_actions=True)
    w_act.index_copy_(1, pos, pa.index_select(1, pos))
    return pol.forward_tokens(w_rtg, w_emb, w_ts, w_task, w_act, eval_rtg=True)
print(f"B={B}: fused kernel {t(lambda: fp.step(w_rtg, w_emb, w_act, w_ts, w_task, pos, ao, ro)):.1f} us, two PyTorch forwards (eager) {t(two_forwards):.1f} us")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): two_forwards()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g): two_forwards()
print(f"       two PyTorch forwards replayed from a CUDA graph {t(g.replay):.1f} us")

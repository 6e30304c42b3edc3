"""Decision-transformer policy that PRODUCES the actions of the hot path (reference
``transformer/decision_transformer.py:106-275``).  It is a caller of the environment, tiny (1.3 M parameters) and
stays plain PyTorch; this module exists so that BASELINE config 2 ("30 iterations
driven by a random-init decision transformer policy") can run batched and at 256x256.

Parameter names and shapes equal the reference's, so its checkpoints load with ``load_state_dict``.
Differences: eval mode only (no dropout), any batch size, and observations of any ``H x W`` are resampled to
the encoder's native 128 x 128 (the reference hard-wires ``reshape(-1, 1, 128, 128)``, ``:215``); at 128 x 128
the resampling is the identity and outputs equal the reference's.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

ENC = 128   # native observation size of the state encoder (Linear(2304, .) after 8/4, 4/2, 3/1 convs)


class _Attn(nn.Module):
    def __init__(self, d, heads, block):
        super().__init__()
        self.qkv_proj = nn.Linear(d, 3 * d)
        self.o_proj = nn.Linear(d, d)
        self.heads = heads
        self.register_buffer("masking", torch.tril(torch.ones(block, block)).view(1, 1, block, block))

    def forward(self, x):
        B, T, E = x.shape
        q, k, v = self.qkv_proj(x).view(B, T, 3, self.heads, E // self.heads).permute(2, 0, 3, 1, 4)
        y = F.scaled_dot_product_attention(q, k, v, is_causal=True)        # == masked softmax(q k^T / sqrt(d)) v
        return self.o_proj(y.transpose(1, 2).reshape(B, T, E))


class _MLP(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.fc = nn.Linear(d, 4 * d)
        self.fc_proj = nn.Linear(4 * d, d)

    def forward(self, x):
        return self.fc_proj(F.gelu(self.fc(x)))


class _Block(nn.Module):
    def __init__(self, d, heads, block):
        super().__init__()
        self.ln1 = nn.LayerNorm(d)
        self.c_att = _Attn(d, heads, block)
        self.ln2 = nn.LayerNorm(d)
        self.mlp = _MLP(d)

    def forward(self, x):
        x = x + self.c_att(self.ln1(x))
        return self.mlp(self.ln2(x))          # the reference's MLP branch has no residual (:101); kept as is


class DecisionTransformer(nn.Module):
    def __init__(self, block_size: int = 18, n_embeds: int = 9, mode: str = "norm", embed_dim: int = 128,
                 n_heads: int = 4, n_blocks: int = 5, action_dim: int = 3, max_timestep: int = 30):
        super().__init__()
        d = embed_dim
        self.action_dim, self.embed_dim = action_dim, d
        self.time_embed = nn.Embedding(max_timestep, d)
        self.task_embed = nn.Embedding(n_embeds, d)
        self.embed_action = nn.Sequential(nn.Linear(action_dim, d), nn.Tanh())
        self.embed_return = nn.Sequential(nn.Linear(1, d), nn.Tanh())
        self.layer_n = nn.LayerNorm(d)
        self.state_encoder = nn.Sequential(
            nn.Conv2d(1, 8, 8, stride=4), nn.ReLU(), nn.Conv2d(8, 16, 4, stride=2), nn.ReLU(),
            nn.Conv2d(16, 16, 3, stride=1), nn.ReLU(), nn.Flatten(), nn.Linear(2304, d), nn.Tanh())
        self.transformer = nn.Sequential(*[_Block(d, n_heads, block_size) for _ in range(n_blocks)])
        self.predict_action = nn.Sequential(nn.Linear(d, action_dim), nn.Sigmoid())
        self.predict_rtg = nn.Linear(d, 1)
        # action head order / scaling (reference :138-154)
        keys = ("mu", "sigma_d", "T") if mode == "flex" else ("T", "sigma_d", "mu")
        self.action_keys = keys
        self.action_scale = {"T": 1.0, "sigma_d": 70.0 / 255.0, "mu": 1.0}
        self.apply(self._init)
        self.eval()

    @staticmethod
    def _init(m):
        # reference :156-163
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1)

    def encode_states(self, states: torch.Tensor, hw: tuple[int, int] | None = None) -> torch.Tensor:
        """``[B, K, H*W]`` (or ``[B,K,H,W]``) observations -> ``[B, K, d]``."""
        B, K = states.shape[:2]
        if states.dim() == 3:
            n = states.shape[2]
            h, w = hw if hw is not None else (int(math.isqrt(n)),) * 2
            img = states.reshape(B * K, 1, h, w)
        else:
            img = states.reshape(B * K, 1, *states.shape[2:])
        if img.shape[-2:] != (ENC, ENC):
            img = F.interpolate(img, size=(ENC, ENC), mode="area")
        return self.state_encoder(img).reshape(B, K, -1)

    @torch.no_grad()
    def forward(self, rtg, states, timesteps, task, actions=None, eval_rtg=False, eval_actions=False, hw=None):
        """Same call as the reference (:212): ``rtg [B,K,1]``, ``states [B,K,H*W]``, ``timesteps [B,K,1]``,
        ``task [B,K]``, ``actions [B,K,3] | None`` -> ``(pred_actions, action_dict)`` or ``pred_rtg``."""
        return self.forward_tokens(rtg, self.encode_states(states, hw), timesteps, task, actions, eval_rtg, eval_actions)

    @torch.no_grad()
    def forward_tokens(self, rtg, state_emb, timesteps, task, actions=None, eval_rtg=False, eval_actions=False):
        """``forward`` with the observations already encoded (``state_emb [B,K,d]`` from ``encode_states``): a rollout
        encodes every observation once instead of K times per call."""
        B, K = state_emb.shape[:2]
        r = self.embed_return(rtg)
        if r.dim() == 2:                     # a single position broadcast over the window (reference eval.py:90-95 passes
            r = r.unsqueeze(1).expand(B, K, -1)   # rtg[:, K] and actions[:, K]; its slice assignment broadcasts them)
        s = state_emb + self.task_embed(task)
        t = self.time_embed(timesteps.to(torch.int64).reshape(B, -1))
        if actions is not None:
            a = self.embed_action(actions)
            if a.dim() == 2:
                a = a.unsqueeze(1).expand(B, K, -1)
            tok = torch.stack([r, s, a], dim=2).reshape(B, 3 * K, -1)
            tt = torch.repeat_interleave(t, 3, dim=1)
        else:
            tok = torch.stack([r, s], dim=2).reshape(B, 2 * K, -1)
            tt = torch.repeat_interleave(t, 2, dim=1)
        x = self.layer_n(self.transformer(tok + tt))
        pred_rtg = None
        if actions is not None:
            pa = self.predict_action(x[:, 1::3, :])
            pred_rtg = self.predict_rtg(x[:, 2::3, :])
        else:
            pa = self.predict_action(x[:, 1::2, :])
        parts = torch.split(pa, pa.shape[-1] // self.action_dim, dim=-1)
        action_dict = OrderedDict((k, parts[i] * self.action_scale[k]) for i, k in enumerate(self.action_keys))
        pred_actions = torch.cat(list(action_dict.values()), dim=-1)
        if eval_rtg:
            return pred_rtg
        if eval_actions or actions is None:
            return pred_actions, action_dict
        return torch.cat([pred_actions, pred_rtg], dim=-1), action_dict


class FusedPolicy:
    """One-kernel rollout step of ``DecisionTransformer`` (``pnp_policy_step``, csrc/policy.cu): the action head at the
    newest observation and the return head at the new action in a single launch instead of two PyTorch forwards (~180 small
    kernels).  ``FusedPolicy(policy)`` packs the weights once (re-pack after loading a checkpoint)."""

    def __init__(self, policy: "DecisionTransformer"):
        from . import _lib
        self._lib = _lib
        self.policy = policy
        dev = next(policy.parameters()).device
        if dev.type != "cuda":
            raise _lib.PnpError("FusedPolicy needs the policy on a CUDA device")
        d = policy.embed_dim
        if d != 128 or policy.action_dim != 3 or len(policy.transformer) != 5:
            raise _lib.PnpError("FusedPolicy is built for the reference configuration (d=128, 5 blocks, 3 actions)")
        self.n_time, self.n_task = policy.time_embed.num_embeddings, policy.task_embed.num_embeddings
        sd = {k: v.detach().float() for k, v in policy.state_dict().items()}
        parts = [sd["embed_return.0.weight"].reshape(-1), sd["embed_return.0.bias"],
                 sd["embed_action.0.weight"].t().contiguous().reshape(-1), sd["embed_action.0.bias"],
                 sd["time_embed.weight"].reshape(-1), sd["task_embed.weight"].reshape(-1),
                 sd["layer_n.weight"], sd["layer_n.bias"],
                 sd["predict_action.0.weight"].reshape(-1), torch.cat([sd["predict_action.0.bias"], sd["predict_action.0.bias"].new_zeros(1)]),
                 sd["predict_rtg.weight"].reshape(-1), torch.cat([sd["predict_rtg.bias"], sd["predict_rtg.bias"].new_zeros(3)])]
        # Per block; the N-split GEMMs (qkv by heads, fc by hidden units) are stored per CTA of the kernel's two-CTA
        # cluster: [2][in][out/2] with CTA r's columns contiguous (csrc/policy.cu, policy_step_cl_kernel).
        dh, heads = d // 4, 4
        cols = [torch.tensor([w * d + (2 * r + hh) * dh + i for w in range(3) for hh in range(2) for i in range(dh)]) for r in range(2)]
        for i in range(5):
            b = f"transformer.{i}."
            qkv_w = sd[b + "c_att.qkv_proj.weight"].t().contiguous()           # [128][384]
            qkv_b = sd[b + "c_att.qkv_proj.bias"]
            fc_w = sd[b + "mlp.fc.weight"].t().contiguous()                    # [128][512]
            fc_b = sd[b + "mlp.fc.bias"]
            h2 = fc_w.shape[1] // 2
            parts += [sd[b + "ln1.weight"], sd[b + "ln1.bias"],
                      torch.cat([qkv_w[:, cols[r].to(qkv_w.device)].contiguous().reshape(-1) for r in range(2)]),
                      torch.cat([qkv_b[cols[r].to(qkv_b.device)] for r in range(2)]),
                      sd[b + "c_att.o_proj.weight"].t().contiguous().reshape(-1), sd[b + "c_att.o_proj.bias"],
                      sd[b + "ln2.weight"], sd[b + "ln2.bias"],
                      torch.cat([fc_w[:, r * h2:(r + 1) * h2].contiguous().reshape(-1) for r in range(2)]),
                      torch.cat([fc_b[r * h2:(r + 1) * h2] for r in range(2)]),
                      sd[b + "mlp.fc_proj.weight"].t().contiguous().reshape(-1), sd[b + "mlp.fc_proj.bias"]]
        parts.append(torch.zeros(d))                                            # bias of the second K slice of a split GEMM
        self.packed = torch.cat([t.reshape(-1).to(dev) for t in parts]).contiguous()
        n = _lib.lib().pnp_policy_packed_floats(self.n_time, self.n_task)
        if self.packed.numel() != n:
            raise _lib.PnpError(f"policy packing mismatch: {self.packed.numel()} floats, the kernel expects {n}")
        self.scales = [float(policy.action_scale[k]) for k in policy.action_keys]
        # state encoder for pnp_policy_observe: conv weights as [.. taps ..][co] (broadcast reads), Linear as [k][o]
        enc = [sd["state_encoder.0.weight"].permute(1, 2, 3, 0).reshape(-1), sd["state_encoder.0.bias"],
               sd["state_encoder.2.weight"].permute(1, 2, 3, 0).reshape(-1), sd["state_encoder.2.bias"],
               sd["state_encoder.4.weight"].permute(1, 2, 3, 0).reshape(-1), sd["state_encoder.4.bias"],
               sd["state_encoder.7.weight"].t().reshape(-1), sd["state_encoder.7.bias"]]
        self.enc_packed = torch.cat([t.contiguous().reshape(-1).to(dev) for t in enc]).contiguous()
        if self.enc_packed.numel() != _lib.lib().pnp_policy_encoder_packed_floats():
            raise _lib.PnpError("state-encoder packing mismatch")

    def step(self, w_rtg, w_emb, w_act, w_ts, w_task, pos, act_out, rtg_out):
        """All tensors fp32 / int64 CUDA, contiguous: the rollout's static context window (``rollout.BatchedRollout``).
        ``w_emb`` holds the state-encoder outputs (without the task embedding).  Writes ``act_out [B,3]``, ``rtg_out [B,1]``
        and the new action into ``w_act[:, pos]``."""
        B, K = w_emb.shape[:2]
        self._lib.check(self._lib.lib().pnp_policy_step(
            self.packed.data_ptr(), w_rtg.data_ptr(), w_emb.data_ptr(), w_act.data_ptr(), w_ts.data_ptr(), w_task.data_ptr(),
            pos.data_ptr(), act_out.data_ptr(), rtg_out.data_ptr(), self.scales[0], self.scales[1], self.scales[2], B, K,
            self.n_time, self.n_task, self._lib.stream_ptr()), "pnp_policy_step")

    @staticmethod
    def observe_supported(H: int, W: int) -> bool:
        return H == W and H >= ENC and H % ENC == 0

    def observe(self, x, next_rtg, w_rtg, w_emb, w_act, w_ts, pos, t_dev):
        """``x [B,1,H,W]`` fp32 CUDA (the new reconstructions): encode them and append the new entry (``next_rtg [B]``, the
        encoding, an empty action, time step ``t_dev + 1``) to every trajectory's window, shifting a full window first
        (``pnp_policy_observe``).  ``pos`` / ``t_dev`` are read, not advanced."""
        B, K = w_emb.shape[:2]
        H, W = x.shape[-2:]
        self._lib.check(self._lib.lib().pnp_policy_observe(
            self.enc_packed.data_ptr(), x.data_ptr(), H, W, next_rtg.data_ptr(), w_rtg.data_ptr(), w_emb.data_ptr(),
            w_act.data_ptr(), w_ts.data_ptr(), pos.data_ptr(), t_dev.data_ptr(), B, K, self.n_time,
            self._lib.stream_ptr()), "pnp_policy_observe")

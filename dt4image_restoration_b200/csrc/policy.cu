// Decision-transformer policy step as ONE kernel (reference transformer/decision_transformer.py:106-275; the caller side of
// the hot path, evaluation/eval.py:147-186): per rollout iteration the reference runs two full forwards - the action head at
// the newest observation token, then, with that action written into the context, the return head at the new action token.
// The second forward differs from the first in ONE token (attention is causal), so both are computed here in one pass:
//   tokens 0 .. 3 p + 1 (p = index of the newest context entry) go through the five blocks once, their keys / values stay in
//   shared memory, the action head reads token 3 p + 1; the action token 3 p + 2 is then embedded and pushed through the
//   blocks alone against the cached keys / values, and the return head reads it.
// One CTA per trajectory, fp32 CUDA-core math (1.3 M parameters, 18 tokens: launch-latency, not FLOPs, is what the ~180 small
// PyTorch kernels of the two forwards cost - 0.5 ms per iteration at batch 64, profiles/r01_rollout_static_window.txt).
// Weights arrive as ONE flat fp32 buffer packed by policy.FusedPolicy (GEMM weights transposed to [in][out] so that
// consecutive threads read consecutive words); layout = struct PolicyOffsets below.
#include "common.cuh"
#include "pnp_internal.h"

namespace pnp {

constexpr int kPD = 128, kPHeads = 4, kPDh = 32, kPBlocks = 5, kPFF = 512, kPA = 3, kPMaxTok = 18, kPThreads = 256;
constexpr int kPKP = 129;   // pitch of the cached keys
constexpr int kPWringOff = (2 * kPMaxTok * kPD + kPMaxTok * kPFF + kPBlocks * kPMaxTok * (kPKP + kPD) + 3) / 4 * 4;

struct PolicyOffsets {           // float offsets into the packed buffer
  int er_w, er_b, ea_w, ea_b, time, task, lnf_g, lnf_b, pa_w, pa_b, pr_w, pr_b, blocks, block_stride;
  // inside a block: ln1_g, ln1_b, qkv_wt [128][384], qkv_b, o_wt [128][128], o_b, ln2_g, ln2_b, fc_wt [128][512], fc_b,
  // pj_wt [512][128], pj_b
  int ln1_g, ln1_b, qkv_w, qkv_b, o_w, o_b, ln2_g, ln2_b, fc_w, fc_b, pj_w, pj_b;
};

__host__ __device__ inline PolicyOffsets policy_offsets(int n_time, int n_task) {
  PolicyOffsets o{};
  int p = 0;
  o.er_w = p; p += kPD; o.er_b = p; p += kPD;
  o.ea_w = p; p += kPD * kPA; o.ea_b = p; p += kPD;         // ea_w stored [3][128]
  o.time = p; p += n_time * kPD;
  o.task = p; p += n_task * kPD;
  o.lnf_g = p; p += kPD; o.lnf_b = p; p += kPD;
  o.pa_w = p; p += kPA * kPD; o.pa_b = p; p += 4;           // pa_w [3][128]
  o.pr_w = p; p += kPD; o.pr_b = p; p += 4;
  o.blocks = p;
  int q = 0;
  o.ln1_g = q; q += kPD; o.ln1_b = q; q += kPD;
  o.qkv_w = q; q += kPD * 3 * kPD; o.qkv_b = q; q += 3 * kPD;
  o.o_w = q; q += kPD * kPD; o.o_b = q; q += kPD;
  o.ln2_g = q; q += kPD; o.ln2_b = q; q += kPD;
  o.fc_w = q; q += kPD * kPFF; o.fc_b = q; q += kPFF;
  o.pj_w = q; q += kPFF * kPD; o.pj_b = q; q += kPD;
  o.block_stride = q;
  return o;
}

size_t policy_packed_floats(int n_time, int n_task) {
  const PolicyOffsets o = policy_offsets(n_time, n_task);
  return size_t(o.blocks) + size_t(kPBlocks) * o.block_stride;
}

struct PolicyParams {
  const float* w;               // packed weights
  const float* rtg;             // [B][K][1]
  const float* emb;             // [B][K][128] encoded observations (state encoder output, before the task embedding)
  float* act;                   // [B][K][3]  in: actions of the older entries; out: entry `pos` receives the new action
  const long long* ts;          // [B][K]
  const long long* task;        // [B][K]
  const long long* pos;         // [1]: index of the newest context entry
  float* act_out;               // [B][3]  predicted (scaled) action = the action dict values in head order
  float* rtg_out;               // [B][1]  predicted return-to-go at the new action token
  float scale0, scale1, scale2; // action scaling in head order (reference :138-154)
  int K, n_time, n_task;
};

// ---- weight stream: every GEMM weight matrix passes through a two-stage ring of shared memory, filled by bulk-async
// copies (cp.async.bulk, one elected thread) that run AHEAD of the arithmetic, across GEMM boundaries: a GEMM's last two
// refills fetch the first two tiles of the next one.  (Measured before this: weights read with plain loads inside the
// FMA loop - 54 % of the kernel's stall samples were FFMAs waiting for them, 584 us per step; profiles/r02_policy_steps.txt.)
constexpr int kPKT = 16;                                      // k rows per tile
constexpr int kPStageFloats = kPKT * kPFF;                    // largest tile: 16 x 512 floats = 32 KB
struct WStream {
  float* stage[2];
  uint64_t* full;                                             // [2]
  uint32_t g;                                                 // tiles consumed so far (same value in every thread)
};
__device__ __forceinline__ void ws_issue(const WStream& ws, uint32_t tile_g, const float* src, uint32_t floats) {
  const uint32_t st = tile_g & 1u;
  mbar_arrive_expect_tx(&ws.full[st], floats * 4u);
  bulk_load_1d(ws.stage[st], src, floats * 4u, &ws.full[st]);
}

// out[t][o] = sum_k in[t][k] * Wt[k][o] + b[o] for t < ntok (NTOK1: one token, `in` / `out` are its rows), o < N.
// Thread (lane, warp): output columns lane + 32 c.  Many tokens: tokens warp, warp + 8, warp + 16, all k.  One token: the
// k rows of a tile are split over the warps (two each) and the partial sums meet in `red` ([8][N]).
// next_w / next_n: the weight matrix the FOLLOWING call consumes (its first two tiles are prefetched here), or null.
template <int KD, int N, bool GELU, bool NTOK1>
__device__ __forceinline__ void policy_gemm(const float* __restrict__ in, const float* __restrict__ Wt,
                                            const float* __restrict__ bias, float* __restrict__ out, int out_pitch, int ntok,
                                            WStream& ws, const float* next_w, int next_n, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int CPT = N / 32, TPT = NTOK1 ? 1 : 3, NT = KD / kPKT;
  float acc[TPT][CPT];
#pragma unroll
  for (int i = 0; i < TPT; ++i)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;
#pragma unroll 1
  for (int it = 0; it < NT; ++it) {
    const uint32_t st = ws.g & 1u;
    mbar_wait(&ws.full[st], (ws.g >> 1) & 1u);
    const float* wt = ws.stage[st] + lane;
    if constexpr (NTOK1) {
#pragma unroll
      for (int kk = 0; kk < kPKT / 8; ++kk) {
        const int k = warp * (kPKT / 8) + kk;
        const float xv = in[it * kPKT + k];
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[0][c] = fmaf(xv, wt[k * N + 32 * c], acc[0][c]);
      }
    } else {
#pragma unroll
      for (int k4 = 0; k4 < kPKT; k4 += 4) {
        float4 x[TPT];
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
          const int t = warp + 8 * i;
          x[i] = (t < ntok) ? *reinterpret_cast<const float4*>(in + t * KD + it * kPKT + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float w[CPT];
#pragma unroll
          for (int c = 0; c < CPT; ++c) w[c] = wt[(k4 + j) * N + 32 * c];
#pragma unroll
          for (int i = 0; i < TPT; ++i) {
            const float xv = j == 0 ? x[i].x : (j == 1 ? x[i].y : (j == 2 ? x[i].z : x[i].w));
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(xv, w[c], acc[i][c]);
          }
        }
      }
    }
    __syncthreads();                                         // every thread is done with this stage
    if (threadIdx.x == 0) {
      const int nx = it + 2;
      if (nx < NT) ws_issue(ws, ws.g + 2, Wt + size_t(nx) * kPKT * N, kPKT * N);
      else if (next_w) ws_issue(ws, ws.g + 2, next_w + size_t(nx - NT) * kPKT * next_n, kPKT * next_n);
    }
    ++ws.g;
  }
  if constexpr (NTOK1) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) red[warp * N + lane + 32 * c] = acc[0][c];
    __syncthreads();
    for (int o = threadIdx.x; o < N; o += kPThreads) {
      float v = __ldg(bias + o);
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) v += red[w8 * N + o];
      if (GELU) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
      out[o] = v;
    }
  } else {
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float b = __ldg(bias + lane + 32 * c);
#pragma unroll
      for (int i = 0; i < TPT; ++i) {
        const int t = warp + 8 * i;
        if (t < ntok) {
          float v = acc[i][c] + b;
          if (GELU) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
          out[t * out_pitch + lane + 32 * c] = v;
        }
      }
    }
  }
}

// LayerNorm (eps 1e-5, biased variance) of rows [t0, t1) of x [.][128] into y; a warp per row
__device__ __forceinline__ void policy_layernorm(const float* x, float* y, const float* g, const float* b, int t0, int t1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = t0 + warp; t < t1; t += kPThreads / 32) {
    float v[4], s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = x[t * kPD + lane + 32 * i]; s += v[i]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / kPD);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float d = v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / kPD) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      y[t * kPD + c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
    }
  }
}

// causal attention for queries [q0, q1): qkv of the queries in `qkv` ([.][384], rows indexed by token - q0 when `rel`), keys /
// values of all tokens in kc / vc ([tok][128]); out [.][128].  A warp per (query, head): lane = key index, then lane = dim.
__device__ __forceinline__ void policy_attention(const float* qkv, int qrow0, const float* kc, const float* vc, float* out,
                                                 int q0, int q1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = 0.17677669529663687f;                  // 1 / sqrt(32)
  for (int job = warp; job < (q1 - q0) * kPHeads; job += kPThreads / 32) {
    const int t = q0 + job / kPHeads, h = job % kPHeads;
    const float* q = qkv + (t - qrow0) * 3 * kPD + h * kPDh;
    float s = -INFINITY;
    if (lane <= t) {
      const float* k = kc + lane * kPKP + h * kPDh;
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < kPDh; ++i) d = fmaf(q[i], k[i], d);
      s = d * scale;
    }
    float mx = s;
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e = (lane <= t) ? __expf(s - mx) : 0.f;
    float sum = e;
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float p = e / sum;
    float acc = 0.f;                                          // lane = output dim
    for (int j = 0; j <= t; ++j) acc = fmaf(__shfl_sync(0xffffffffu, p, j), vc[j * kPD + h * kPDh + lane], acc);
    out[(t - qrow0) * kPD + h * kPDh + lane] = acc;
  }
}

__global__ void __launch_bounds__(kPThreads) policy_step_kernel(const PolicyParams p) {
  extern __shared__ __align__(16) float psm[];
  float* x = psm;                                   // [18][128] residual stream
  float* hbuf = x + kPMaxTok * kPD;                 // [18][128] LayerNorm output / attention output
  float* big = hbuf + kPMaxTok * kPD;               // [18][512] qkv (384 used) / MLP hidden
  float* kcache = big + kPMaxTok * kPFF;            // [5][18][129]: odd pitch, the score loop walks it with lanes along tokens
  float* vcache = kcache + kPBlocks * kPMaxTok * kPKP;
  float* wring = psm + kPWringOff;                          // 2 x [16][512] weight tiles, 16-byte aligned
  __shared__ __align__(8) uint64_t ws_full[2];
  WStream ws{{wring, wring + kPStageFloats}, ws_full, 0u};
  const PolicyOffsets O = policy_offsets(p.n_time, p.n_task);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int pos = int(*p.pos);
  const int n1 = 3 * pos + 2;                       // tokens 0 .. 3 pos + 1 take part in the first pass
  const int K = p.K;
  const float* W = p.w;
  const float* Wb0 = W + O.blocks;
  if (tid == 0) {
    mbar_init(&ws_full[0], 1);
    mbar_init(&ws_full[1], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
    ws_issue(ws, 0, Wb0 + O.qkv_w, kPKT * 3 * kPD);           // the first GEMM's first two tiles
    ws_issue(ws, 1, Wb0 + O.qkv_w + kPKT * 3 * kPD, kPKT * 3 * kPD);
  }

  // ---- token embeddings (reference :212-240): (return, observation + task, action) per entry, + time embedding ----
  for (int e = tid; e < (pos + 1) * 3 * kPD; e += kPThreads) {
    const int tok = e / kPD, c = e % kPD, ent = tok / 3, kind = tok % 3;
    const size_t bi = size_t(b) * K + ent;
    float v;
    if (kind == 0) v = tanhf(fmaf(__ldg(W + O.er_w + c), p.rtg[bi], __ldg(W + O.er_b + c)));
    else if (kind == 1) v = p.emb[bi * kPD + c] + __ldg(W + O.task + size_t(p.task[bi]) * kPD + c);
    else {
      const float* a = p.act + bi * kPA;
      v = tanhf(fmaf(__ldg(W + O.ea_w + c), a[0], fmaf(__ldg(W + O.ea_w + kPD + c), a[1],
                fmaf(__ldg(W + O.ea_w + 2 * kPD + c), a[2], __ldg(W + O.ea_b + c)))));
    }
    x[tok * kPD + c] = v + __ldg(W + O.time + size_t(p.ts[bi]) * kPD + c);
  }
  __syncthreads();

  // ---- pass 1: tokens [0, n1) through the blocks; keys / values cached ----
  for (int l = 0; l < kPBlocks; ++l) {
    const float* Wb = W + O.blocks + size_t(l) * O.block_stride;
    policy_layernorm(x, hbuf, Wb + O.ln1_g, Wb + O.ln1_b, 0, n1);
    __syncthreads();
    policy_gemm<kPD, 3 * kPD, false, false>(hbuf, Wb + O.qkv_w, Wb + O.qkv_b, big, 3 * kPD, n1, ws, Wb + O.o_w, kPD, nullptr);
    __syncthreads();
    // torch: qkv.view(B, T, 3, heads, dh): column = which * 128 + head * 32 + dim
    for (int e = tid; e < n1 * kPD; e += kPThreads) {
      const int t = e / kPD, c = e % kPD;
      kcache[(l * kPMaxTok + t) * kPKP + c] = big[t * 3 * kPD + kPD + c];
      vcache[(l * kPMaxTok + t) * kPD + c] = big[t * 3 * kPD + 2 * kPD + c];
    }
    __syncthreads();
    if (l == kPBlocks - 1) {
      // Last block: only token n1 - 1 (the action head's) is read after it - the keys / values of ALL tokens are already
      // cached for pass 2 - so attention, o_proj and the MLP run for that one token (one-token GEMMs, same weight stream).
      const int tq = n1 - 1;
      float* xq = x + tq * kPD;
      float* hq = hbuf + tq * kPD;
      float* red1 = big + 4096;
      policy_attention(big, 0, kcache + l * kPMaxTok * kPKP, vcache + l * kPMaxTok * kPD, hbuf, tq, n1);
      __syncthreads();
      policy_gemm<kPD, kPD, false, true>(hq, Wb + O.o_w, Wb + O.o_b, big + 1024, 0, 1, ws, Wb + O.fc_w, kPFF, red1);
      __syncthreads();
      for (int c = tid; c < kPD; c += kPThreads) xq[c] += big[1024 + c];
      __syncthreads();
      policy_layernorm(x, hbuf, Wb + O.ln2_g, Wb + O.ln2_b, tq, n1);
      __syncthreads();
      policy_gemm<kPD, kPFF, true, true>(hq, Wb + O.fc_w, Wb + O.fc_b, big, 0, 1, ws, Wb + O.pj_w, kPD, red1);
      __syncthreads();
      policy_gemm<kPFF, kPD, false, true>(big, Wb + O.pj_w, Wb + O.pj_b, xq, 0, 1, ws, Wb0 + O.qkv_w, 3 * kPD, red1);
      __syncthreads();
      break;
    }
    policy_attention(big, 0, kcache + l * kPMaxTok * kPKP, vcache + l * kPMaxTok * kPD, hbuf, 0, n1);
    __syncthreads();
    // x += o_proj(att)
    {
      float* tmp = big;                                       // [18][128] (qkv no longer needed)
      policy_gemm<kPD, kPD, false, false>(hbuf, Wb + O.o_w, Wb + O.o_b, tmp, kPD, n1, ws, Wb + O.fc_w, kPFF, nullptr);
      __syncthreads();
      for (int e = tid; e < n1 * kPD; e += kPThreads) x[e] += tmp[e];
      __syncthreads();
    }
    policy_layernorm(x, hbuf, Wb + O.ln2_g, Wb + O.ln2_b, 0, n1);
    __syncthreads();
    policy_gemm<kPD, kPFF, true, false>(hbuf, Wb + O.fc_w, Wb + O.fc_b, big, kPFF, n1, ws, Wb + O.pj_w, kPD, nullptr);
    __syncthreads();
    // no residual (reference :101); the stream continues with the next block's qkv, or with block 0 again for pass 2
    policy_gemm<kPFF, kPD, false, false>(big, Wb + O.pj_w, Wb + O.pj_b, x, kPD, n1, ws,
                                         (l + 1 < kPBlocks ? Wb + O.block_stride : Wb0) + O.qkv_w, 3 * kPD, nullptr);
    __syncthreads();
  }
  // ---- action head at token 3 pos + 1 ----
  policy_layernorm(x, hbuf, W + O.lnf_g, W + O.lnf_b, n1 - 1, n1);
  __syncthreads();
  __shared__ float s_act[4];
  if (tid < 96) {
    const int a = tid >> 5, lane = tid & 31;
    float d = 0.f;
    for (int c = lane; c < kPD; c += 32) d = fmaf(hbuf[(n1 - 1) * kPD + c], __ldg(W + O.pa_w + a * kPD + c), d);
#pragma unroll
    for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) {
      const float sg = 1.f / (1.f + expf(-(d + __ldg(W + O.pa_b + a))));
      const float v = sg * (a == 0 ? p.scale0 : (a == 1 ? p.scale1 : p.scale2));
      s_act[a] = v;
      p.act_out[b * kPA + a] = v;
      p.act[(size_t(b) * K + pos) * kPA + a] = v;             // the context entry receives its action (eval.py:166)
    }
  }
  __syncthreads();

  // ---- pass 2: the action token 3 pos + 2 alone, against the cached keys / values ----
  const int tn = n1;                                          // its token index
  float* xn = x + tn * kPD;
  float* hn = hbuf + tn * kPD;
  float* red = big + 4096;                                  // [8][512] partial sums of the one-token GEMMs
  for (int c = tid; c < kPD; c += kPThreads) {
    const size_t bi = size_t(b) * K + pos;
    xn[c] = tanhf(fmaf(__ldg(W + O.ea_w + c), s_act[0], fmaf(__ldg(W + O.ea_w + kPD + c), s_act[1],
                  fmaf(__ldg(W + O.ea_w + 2 * kPD + c), s_act[2], __ldg(W + O.ea_b + c))))) +
            __ldg(W + O.time + size_t(p.ts[bi]) * kPD + c);
  }
  __syncthreads();
  for (int l = 0; l < kPBlocks; ++l) {
    const float* Wb = W + O.blocks + size_t(l) * O.block_stride;
    policy_layernorm(x, hbuf, Wb + O.ln1_g, Wb + O.ln1_b, tn, tn + 1);
    __syncthreads();
    policy_gemm<kPD, 3 * kPD, false, true>(hn, Wb + O.qkv_w, Wb + O.qkv_b, big, 0, 1, ws, Wb + O.o_w, kPD, red);
    __syncthreads();
    for (int c = tid; c < kPD; c += kPThreads) {
      kcache[(l * kPMaxTok + tn) * kPKP + c] = big[kPD + c];
      vcache[(l * kPMaxTok + tn) * kPD + c] = big[2 * kPD + c];
    }
    __syncthreads();
    policy_attention(big, tn, kcache + l * kPMaxTok * kPKP, vcache + l * kPMaxTok * kPD, hn, tn, tn + 1);
    __syncthreads();
    policy_gemm<kPD, kPD, false, true>(hn, Wb + O.o_w, Wb + O.o_b, big + 1024, 0, 1, ws, Wb + O.fc_w, kPFF, red);
    __syncthreads();
    for (int c = tid; c < kPD; c += kPThreads) xn[c] += big[1024 + c];
    __syncthreads();
    policy_layernorm(x, hbuf, Wb + O.ln2_g, Wb + O.ln2_b, tn, tn + 1);
    __syncthreads();
    policy_gemm<kPD, kPFF, true, true>(hn, Wb + O.fc_w, Wb + O.fc_b, big, 0, 1, ws, Wb + O.pj_w, kPD, red);
    __syncthreads();
    policy_gemm<kPFF, kPD, false, true>(big, Wb + O.pj_w, Wb + O.pj_b, xn, 0, 1, ws,
                                        l + 1 < kPBlocks ? Wb + O.block_stride + O.qkv_w : nullptr, 3 * kPD, red);
    __syncthreads();
  }
  policy_layernorm(x, hbuf, W + O.lnf_g, W + O.lnf_b, tn, tn + 1);
  __syncthreads();
  if (tid < 32) {
    float d = 0.f;
    for (int c = tid; c < kPD; c += 32) d = fmaf(hn[c], __ldg(W + O.pr_w + c), d);
#pragma unroll
    for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (tid == 0) p.rtg_out[b] = d + __ldg(W + O.pr_b);
  }
}

constexpr size_t kPolicySmem = sizeof(float) * (size_t(kPWringOff) + 2 * size_t(kPStageFloats));

int policy_step_launch(const float* w, const float* rtg, const float* emb, float* act, const long long* ts,
                       const long long* task, const long long* pos, float* act_out, float* rtg_out, float s0, float s1,
                       float s2, int B, int K, int n_time, int n_task, cudaStream_t st) {
  if (K < 1 || 3 * K > kPMaxTok || B < 1) return -1;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(policy_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kPolicySmem));
    if (e != cudaSuccess) return int(e);
    attr_done = true;
  }
  PolicyParams p{w, rtg, emb, act, ts, task, pos, act_out, rtg_out, s0, s1, s2, K, n_time, n_task};
  policy_step_kernel<<<B, kPThreads, kPolicySmem, st>>>(p);
  return int(cudaGetLastError());
}

}  // namespace pnp

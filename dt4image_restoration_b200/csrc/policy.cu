// Decision-transformer policy step as ONE kernel (reference transformer/decision_transformer.py:106-275; the caller side of
// the hot path, evaluation/eval.py:147-186): per rollout iteration the reference runs two full forwards - the action head at
// the newest observation token, then, with that action written into the context, the return head at the new action token.
// The second forward differs from the first in ONE token (attention is causal), so both are computed here in one pass:
//   tokens 0 .. 3 p + 1 (p = index of the newest context entry) go through the five blocks once, their keys / values stay in
//   shared memory, the action head reads token 3 p + 1; the action token 3 p + 2 is then embedded and pushed through the
//   blocks alone against the cached keys / values, and the return head reads it.
// A cluster of two CTAs per trajectory (see policy_step_cl_kernel), fp32 CUDA-core math (1.3 M parameters, 18 tokens:
// launch-latency, not FLOPs, is what the ~180 small PyTorch kernels of the two forwards cost - 0.5 ms per iteration at batch
// 64, profiles/r01_rollout_static_window.txt).
// Weights arrive as ONE flat fp32 buffer packed by policy.FusedPolicy (GEMM weights transposed to [in][out] so that
// consecutive threads read consecutive words; qkv and fc stored per CTA, [2][in][out/2]); layout = struct PolicyOffsets below.
#include "common.cuh"
#include "pnp_internal.h"

namespace pnp {

constexpr int kPD = 128, kPHeads = 4, kPDh = 32, kPBlocks = 5, kPFF = 512, kPA = 3, kPMaxTok = 18, kPThreads = 256;

struct PolicyOffsets {           // float offsets into the packed buffer
  int er_w, er_b, ea_w, ea_b, time, task, lnf_g, lnf_b, pa_w, pa_b, pr_w, pr_b, blocks, block_stride;
  // inside a block: ln1_g, ln1_b, qkv_wt [128][384], qkv_b, o_wt [128][128], o_b, ln2_g, ln2_b, fc_wt [128][512], fc_b,
  // pj_wt [512][128], pj_b
  int ln1_g, ln1_b, qkv_w, qkv_b, o_w, o_b, ln2_g, ln2_b, fc_w, fc_b, pj_w, pj_b;
  int zeros;                     // 128 zero floats behind the blocks (the bias of the second K-slice of a split GEMM)
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
  o.zeros = o.blocks + kPBlocks * q;
  return o;
}

size_t policy_packed_floats(int n_time, int n_task) {
  const PolicyOffsets o = policy_offsets(n_time, n_task);
  return size_t(o.zeros) + kPD;
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

// =====================================================================================================================
// Two-CTA cluster version: the FFMA floor of one SM per trajectory is ~290 us (71 M FMA per pass), so a trajectory gets a
// CLUSTER of two CTAs that split every GEMM and keep the residual stream replicated:
//   qkv  (128 -> 384): split along N by HEADS - CTA r computes q, k, v of heads 2r, 2r+1 (its own 192 columns, stored
//                      contiguously per CTA by policy.FusedPolicy), so attention is local and needs no exchange;
//   o    (128 -> 128): split along K - CTA r multiplies its 64 attention columns with rows 64r.. of the weight; the two
//                      partial results are exchanged through distributed shared memory and summed in rank order;
//   fc   (128 -> 512): split along N - CTA r computes hidden units 256r.. (+ GELU);
//   proj (512 -> 128): split along K over those hidden units, exchanged and summed like o.
// Two exchanges per block (n1 x 128 floats, st.shared::cluster + one cluster barrier each); LayerNorms, embeddings and the
// heads are computed by both CTAs on identical data.  Every CTA streams half of every weight matrix.
constexpr int kCH = kPHeads / 2, kCQ = 3 * kCH * kPDh, kCA = kCH * kPDh, kCF = kPFF / 2, kCKP = kCA + 1;
constexpr int kCStageFloats = kPKT * kCF;                    // largest local tile: 16 x 256
constexpr int kCx = 0, kCh = kCx + kPMaxTok * kPD, kCbig = kCh + kPMaxTok * kPD, kCpart = kCbig + kPMaxTok * kCF,
              kCrecv = kCpart + kPMaxTok * kPD, kCkc = kCrecv + 2 * kPMaxTok * kPD, kCvc = kCkc + kPBlocks * kPMaxTok * kCKP,
              kCred = (kCvc + kPBlocks * kPMaxTok * kCA + 3) / 4 * 4, kCring = kCred + 8 * kCF,
              kCTotal = kCring + 2 * kCStageFloats;

template <int NH, int QP, int KP, int VP, int OP>
__device__ __forceinline__ void policy_attention_t(const float* qkv, int qrow0, const float* kc, const float* vc, float* out,
                                                   int q0, int q1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = 0.17677669529663687f;                  // 1 / sqrt(32)
  for (int job = warp; job < (q1 - q0) * NH; job += kPThreads / 32) {
    const int t = q0 + job / NH, h = job % NH;
    const float* q = qkv + (t - qrow0) * QP + h * kPDh;
    float s = -INFINITY;
    if (lane <= t) {
      const float* k = kc + lane * KP + h * kPDh;
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
    const float pr = e / sum;
    float acc = 0.f;                                          // lane = output dim
    for (int j = 0; j <= t; ++j) acc = fmaf(__shfl_sync(0xffffffffu, pr, j), vc[j * VP + h * kPDh + lane], acc);
    out[(t - qrow0) * OP + h * kPDh + lane] = acc;
  }
}

// rows [r0, r1) of `part` ([.][128]) go to the peer's receive buffer; returns after the peer's rows have arrived here
__device__ __forceinline__ void policy_exchange(const float* part, uint32_t peer_recv, int r0, int r1) {
  const int n4 = (r1 - r0) * (kPD / 4);
  for (int i = threadIdx.x; i < n4; i += kPThreads) {
    const float4 v = *reinterpret_cast<const float4*>(part + r0 * kPD + 4 * i);
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(peer_recv + uint32_t(r0 * kPD + 4 * i) * 4u), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}

__global__ void __launch_bounds__(kPThreads) policy_step_cl_kernel(const PolicyParams p) {
  extern __shared__ __align__(16) float psm[];
  float* x = psm + kCx;                             // [18][128] residual stream (identical in both CTAs)
  float* hbuf = psm + kCh;                          // [18][128] LayerNorm output; [18][64] attention output
  float* big = psm + kCbig;                         // [18][192] local qkv / [18][256] local MLP hidden
  float* part = psm + kCpart;                       // [18][128] this CTA's partial sum of a K-split GEMM
  float* recv = psm + kCrecv;                       // [2][18][128] the peer's partial sums (alternating buffers)
  float* kcache = psm + kCkc;                       // [5][18][65] keys of the local heads
  float* vcache = psm + kCvc;                       // [5][18][64]
  float* red = psm + kCred;                         // [8][256] one-token GEMM partial sums
  float* wring = psm + kCring;
  __shared__ __align__(8) uint64_t ws_full[2];
  WStream ws{{wring, wring + kCStageFloats}, ws_full, 0u};
  const PolicyOffsets O = policy_offsets(p.n_time, p.n_task);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int b = blockIdx.x >> 1, tid = threadIdx.x;
  const int pos = int(*p.pos);
  const int n1 = 3 * pos + 2;
  const int K = p.K;
  const float* W = p.w;
  const float* Wb0 = W + O.blocks;
  const float* zeros = W + O.zeros;
  auto qkv_w = [&](const float* Wb) { return Wb + O.qkv_w + rank * (kPD * kCQ); };
  if (tid == 0) {
    mbar_init(&ws_full[0], 1);
    mbar_init(&ws_full[1], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
    ws_issue(ws, 0, qkv_w(Wb0), kPKT * kCQ);
    ws_issue(ws, 1, qkv_w(Wb0) + kPKT * kCQ, kPKT * kCQ);
  }
  uint32_t peer_recv;
  {
    const uint32_t local = smem_u32(recv);
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_recv) : "r"(local), "r"(rank ^ 1u));
  }
  int xch = 0;                                      // exchanges done (selects the receive buffer)
  // sum of the two partial results in RANK order (bitwise the same in both CTAs)
  auto combined = [&](int e) {
    const float mine = part[e], theirs = recv[(xch & 1) * kPMaxTok * kPD + e];
    return rank == 0 ? mine + theirs : theirs + mine;
  };

  // ---- token embeddings (reference :212-240), replicated ----
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
  // both CTAs are running (their shared memory may be written remotely from here on)
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");

  // one transformer block for token rows [t0, t1) (MANY = false: the single row t0 with the one-token GEMMs); the qkv of
  // the rows lands in `big` (row-relative), keys / values in the cache
  auto block = [&](int l, int t0, int t1, bool many) {
    const float* Wb = W + O.blocks + size_t(l) * O.block_stride;
    const float* o_w = Wb + O.o_w + rank * (kCA * kPD);
    const float* qkv_b = Wb + O.qkv_b + rank * kCQ;
    policy_layernorm(x, hbuf, Wb + O.ln1_g, Wb + O.ln1_b, t0, t1);
    __syncthreads();
    if (many) policy_gemm<kPD, kCQ, false, false>(hbuf, qkv_w(Wb), qkv_b, big, kCQ, t1, ws, o_w, kPD, nullptr);
    else policy_gemm<kPD, kCQ, false, true>(hbuf + t0 * kPD, qkv_w(Wb), qkv_b, big, 0, 1, ws, o_w, kPD, red);
    __syncthreads();
    for (int e = tid; e < (t1 - t0) * kCA; e += kPThreads) {
      const int r = e / kCA, c = e % kCA, t = t0 + r;
      kcache[(l * kPMaxTok + t) * kCKP + c] = big[r * kCQ + kCA + c];
      vcache[(l * kPMaxTok + t) * kCA + c] = big[r * kCQ + 2 * kCA + c];
    }
    __syncthreads();
    return Wb;
  };
  // attention + o_proj + MLP of block l for rows [t0, t1) whose qkv sits in `big` starting at row `qrow0`
  auto finish = [&](int l, int t0, int t1, int qrow0, bool many, const float* next_qkv) {
    const float* Wb = W + O.blocks + size_t(l) * O.block_stride;
    const float* o_w = Wb + O.o_w + rank * (kCA * kPD);
    const float* fc_w = Wb + O.fc_w + rank * (kPD * kCF);
    const float* fc_b = Wb + O.fc_b + rank * kCF;
    const float* pj_w = Wb + O.pj_w + rank * (kCF * kPD);
    const float* o_b = rank == 0 ? Wb + O.o_b : zeros;
    const float* pj_b = rank == 0 ? Wb + O.pj_b : zeros;
    // attention of the local heads: rows land in hbuf as [row][64] (row-relative to qrow0)
    policy_attention_t<kCH, kCQ, kCKP, kCA, kCA>(big, qrow0, kcache + l * kPMaxTok * kCKP, vcache + l * kPMaxTok * kCA,
                                                 hbuf, t0, t1);
    __syncthreads();
    const int a0 = t0 - qrow0;                      // first attention row in hbuf
    if (many) policy_gemm<kCA, kPD, false, false>(hbuf, o_w, o_b, part, kPD, t1, ws, fc_w, kCF, nullptr);
    else policy_gemm<kCA, kPD, false, true>(hbuf + a0 * kCA, o_w, o_b, part + t0 * kPD, 0, 1, ws, fc_w, kCF, red);
    __syncthreads();
    policy_exchange(part, peer_recv + uint32_t((xch & 1) * kPMaxTok * kPD) * 4u, t0, t1);
    for (int e = t0 * kPD + tid; e < t1 * kPD; e += kPThreads) x[e] += combined(e);
    ++xch;
    __syncthreads();
    policy_layernorm(x, hbuf, Wb + O.ln2_g, Wb + O.ln2_b, t0, t1);
    __syncthreads();
    if (many) policy_gemm<kPD, kCF, true, false>(hbuf, fc_w, fc_b, big, kCF, t1, ws, pj_w, kPD, nullptr);
    else policy_gemm<kPD, kCF, true, true>(hbuf + t0 * kPD, fc_w, fc_b, big, 0, 1, ws, pj_w, kPD, red);
    __syncthreads();
    if (many) policy_gemm<kCF, kPD, false, false>(big, pj_w, pj_b, part, kPD, t1, ws, next_qkv, kCQ, nullptr);
    else policy_gemm<kCF, kPD, false, true>(big, pj_w, pj_b, part + t0 * kPD, 0, 1, ws, next_qkv, kCQ, red);
    __syncthreads();
    policy_exchange(part, peer_recv + uint32_t((xch & 1) * kPMaxTok * kPD) * 4u, t0, t1);
    for (int e = t0 * kPD + tid; e < t1 * kPD; e += kPThreads) x[e] = combined(e);      // no residual (reference :101)
    ++xch;
    __syncthreads();
  };

  // ---- pass 1: tokens [0, n1); in the last block only the action head's token goes past the key / value cache ----
  for (int l = 0; l < kPBlocks; ++l) {
    const float* Wb = block(l, 0, n1, true);
    const bool last = l == kPBlocks - 1;
    const float* next = qkv_w(last ? Wb0 : Wb + O.block_stride);
    if (!last) finish(l, 0, n1, 0, true, next);
    else finish(l, n1 - 1, n1, 0, false, next);
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
      if (rank == 0) {
        p.act_out[b * kPA + a] = v;
        p.act[(size_t(b) * K + pos) * kPA + a] = v;             // the context entry receives its action (eval.py:166)
      }
    }
  }
  __syncthreads();
  // ---- pass 2: the action token 3 pos + 2 alone, against the cached keys / values ----
  const int tn = n1;
  for (int c = tid; c < kPD; c += kPThreads) {
    const size_t bi = size_t(b) * K + pos;
    x[tn * kPD + c] = tanhf(fmaf(__ldg(W + O.ea_w + c), s_act[0], fmaf(__ldg(W + O.ea_w + kPD + c), s_act[1],
                            fmaf(__ldg(W + O.ea_w + 2 * kPD + c), s_act[2], __ldg(W + O.ea_b + c))))) +
                      __ldg(W + O.time + size_t(p.ts[bi]) * kPD + c);
  }
  __syncthreads();
  for (int l = 0; l < kPBlocks; ++l) {
    const float* Wb = block(l, tn, tn + 1, false);
    finish(l, tn, tn + 1, tn, false, l + 1 < kPBlocks ? qkv_w(Wb + O.block_stride) : nullptr);
  }
  policy_layernorm(x, hbuf, W + O.lnf_g, W + O.lnf_b, tn, tn + 1);
  __syncthreads();
  if (tid < 32 && rank == 0) {
    float d = 0.f;
    for (int c = tid; c < kPD; c += 32) d = fmaf(hbuf[tn * kPD + c], __ldg(W + O.pr_w + c), d);
#pragma unroll
    for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (tid == 0) p.rtg_out[b] = d + __ldg(W + O.pr_b);
  }
}

constexpr size_t kPolicyClSmem = sizeof(float) * size_t(kCTotal);


int policy_step_launch(const float* w, const float* rtg, const float* emb, float* act, const long long* ts,
                       const long long* task, const long long* pos, float* act_out, float* rtg_out, float s0, float s1,
                       float s2, int B, int K, int n_time, int n_task, cudaStream_t st) {
  if (K < 1 || 3 * K > kPMaxTok || B < 1) return -1;
  PolicyParams p{w, rtg, emb, act, ts, task, pos, act_out, rtg_out, s0, s1, s2, K, n_time, n_task};
  {
    static bool cl_attr_done = false;
    if (!cl_attr_done) {
      cudaError_t e = cudaFuncSetAttribute(policy_step_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kPolicyClSmem));
      if (e != cudaSuccess) return int(e);
      cl_attr_done = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * B); cfg.blockDim = dim3(kPThreads); cfg.dynamicSmemBytes = kPolicyClSmem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return int(cudaLaunchKernelEx(&cfg, policy_step_cl_kernel, p));
  }
  return int(cudaGetLastError());
}

}  // namespace pnp

// Observation side of a decision-transformer rollout iteration as ONE kernel (the caller side of the hot path,
// reference evaluation/eval.py:204-217 + transformer/decision_transformer.py:128-132,215): the new reconstruction x of every
// trajectory is reduced to the encoder's 128x128 input (area mean, what F.interpolate(mode='area') does for an integer
// factor), pushed through the state encoder (Conv 1->8 k8 s4, ReLU, Conv 8->16 k4 s2, ReLU, Conv 16->16 k3 s1, ReLU,
// Flatten, Linear 2304->128, Tanh) and appended to the trajectory's context window (a full window first moves one entry to
// the left; the new entry gets the predicted return-to-go, an empty action and the next time step).
// Replaces ~38 small PyTorch / cuDNN launches per iteration (adaptive pool, two layout conversions, three convs, a split-K
// GEMM, index_select / index_copy per window tensor: 0.25 ms at batch 64, profiles/r02_rollout_iteration_kernels.txt).
// One CTA per trajectory, fp32, everything after the input read stays in shared memory.
#include "common.cuh"
#include "pnp_internal.h"

namespace pnp {

constexpr int kOE = 128;                 // encoder input edge
constexpr int kO1 = 31, kO2 = 14, kO3 = 12, kOD = 128, kOFlat = 16 * kO3 * kO3;   // 2304
constexpr int kOThreads = 1024;          // the Linear layer and the input read want many loads in flight
constexpr int kOSlices = kOThreads / kOD; // k slices of the Linear layer
// packed encoder weights (floats): c1w [8][8][8] (ky, kx, co) | c1b [8] | c2w [8][4][4][16] (ci, ky, kx, co) | c2b [16] |
// c3w [16][3][3][16] | c3b [16] | lwT [2304][128] (k, o) | lb [128]
constexpr int kOc1w = 0, kOc1b = kOc1w + 512, kOc2w = kOc1b + 8, kOc2b = kOc2w + 2048, kOc3w = kOc2b + 16,
              kOc3b = kOc3w + 2304, kOlw = kOc3b + 16, kOlb = kOlw + kOFlat * kOD, kOTotal = kOlb + kOD;
constexpr int kOConvW = kOlw;            // the conv weights and biases (4904 floats) are staged in shared memory

size_t policy_encoder_packed_floats() { return size_t(kOTotal); }

struct ObserveParams {
  const float* w;          // packed encoder weights
  const float* x;          // [B][H][W] reconstructions
  int H, W, f;             // f = H / 128 = W / 128
  const float* nxt_rtg;    // [B] return-to-go of the new entry
  float* w_rtg;            // [B][K][1]
  float* w_emb;            // [B][K][128]
  float* w_act;            // [B][K][3]
  long long* w_ts;         // [B][K][1]
  const long long* pos;    // [1] newest entry BEFORE this call
  const long long* t_dev;  // [1] time step of that entry
  int K, n_time;
};

__global__ void __launch_bounds__(kOThreads) policy_observe_kernel(const ObserveParams p) {
  extern __shared__ __align__(16) float osm[];
  float* img = osm;                           // [4][128][32]: pixel (y, x) at [x & 3][y][x >> 2] (stride-4 reads are contiguous)
  float* c1 = img + kOE * kOE;                // [8][31][31]
  float* c2 = c1 + 8 * kO1 * kO1;             // [16][14][14]
  float* c3 = c2 + 16 * kO2 * kO2;            // [2304] in Flatten order (c, y, x)
  float* wsm = c3 + kOFlat;                   // conv weights + biases
  float* red = wsm + kOConvW;                 // [kOSlices][128] partial sums of the linear layer
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* W = p.w;

  for (int i = tid; i < kOConvW; i += kOThreads) wsm[i] = __ldg(W + i);
  // ---- area mean to 128 x 128 ----
  {
    const float* xb = p.x + size_t(b) * p.H * p.W;
    const int f = p.f;
    const float inv = 1.f / float(f * f);
    if (f == 2) {
      // a thread per PAIR of pooled pixels: two 16-byte loads, eight pairs (16 loads) in flight
      constexpr int kPairs = kOE * kOE / 2, kPer = kPairs / kOThreads;      // 8
      float4 r0[kPer], r1[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tid + j * kOThreads, y = i >> 6, xp = i & 63;
        r0[j] = __ldg(reinterpret_cast<const float4*>(xb + size_t(2 * y) * p.W + 4 * xp));
        r1[j] = __ldg(reinterpret_cast<const float4*>(xb + size_t(2 * y + 1) * p.W + 4 * xp));
      }
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tid + j * kOThreads, y = i >> 6, x = 2 * (i & 63);
        img[((x & 3) * kOE + y) * 32 + (x >> 2)] = ((r0[j].x + r0[j].y) + (r1[j].x + r1[j].y)) * inv;
        img[(((x + 1) & 3) * kOE + y) * 32 + ((x + 1) >> 2)] = ((r0[j].z + r0[j].w) + (r1[j].z + r1[j].w)) * inv;
      }
    } else {
      for (int i = tid; i < kOE * kOE; i += kOThreads) {
        const int y = i >> 7, x = i & 127;
        float s = 0.f;
        for (int dy = 0; dy < f; ++dy)
          for (int dx = 0; dx < f; ++dx) s += __ldg(xb + size_t(f * y + dy) * p.W + f * x + dx);
        img[((x & 3) * kOE + y) * 32 + (x >> 2)] = s * inv;
      }
    }
  }
  __syncthreads();
  // ---- conv1: 1 -> 8, 8x8, stride 4 -> [8][31][31]; a thread per output position, all 8 channels ----
  for (int pos = tid; pos < kO1 * kO1; pos += kOThreads) {
    const int oy = pos / kO1, ox = pos % kO1;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = wsm[kOc1b + c];
#pragma unroll 1
    for (int ky = 0; ky < 8; ++ky) {
      const int y = 4 * oy + ky;
#pragma unroll
      for (int kx = 0; kx < 8; ++kx) {
        const float v = img[((kx & 3) * kOE + y) * 32 + ox + (kx >> 2)];
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + kOc1w + (ky * 8 + kx) * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + kOc1w + (ky * 8 + kx) * 8 + 4);
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
        acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) c1[c * kO1 * kO1 + pos] = fmaxf(acc[c], 0.f);
  }
  __syncthreads();
  // ---- conv2: 8 -> 16, 4x4, stride 2 -> [16][14][14]; a thread per (position, quarter of the channels) ----
  for (int item = tid; item < 4 * kO2 * kO2; item += kOThreads) {
    const int q = item / (kO2 * kO2), pos = item % (kO2 * kO2);
    const int oy = pos / kO2, ox = pos % kO2;
    float acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = wsm[kOc2b + q * 4 + c];
#pragma unroll 1
    for (int ci = 0; ci < 8; ++ci) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int ky = k >> 2, kx = k & 3;
        const float v = c1[ci * kO1 * kO1 + (2 * oy + ky) * kO1 + 2 * ox + kx];
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + kOc2w + ((ci * 16 + k) * 16) + q * 4);
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) c2[(q * 4 + c) * kO2 * kO2 + pos] = fmaxf(acc[c], 0.f);
  }
  __syncthreads();
  // ---- conv3: 16 -> 16, 3x3, stride 1 -> [16][12][12] ----
  for (int item = tid; item < 4 * kO3 * kO3; item += kOThreads) {
    const int q = item / (kO3 * kO3), pos = item % (kO3 * kO3);
    const int oy = pos / kO3, ox = pos % kO3;
    float acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = wsm[kOc3b + q * 4 + c];
#pragma unroll 1
    for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int ky = k / 3, kx = k % 3;
        const float v = c2[ci * kO2 * kO2 + (oy + ky) * kO2 + ox + kx];
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + kOc3w + ((ci * 9 + k) * 16) + q * 4);
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) c3[(q * 4 + c) * kO3 * kO3 + pos] = fmaxf(acc[c], 0.f);
  }
  __syncthreads();
  // ---- Linear 2304 -> 128 + Tanh: thread (output o, k slice); weights [k][o] read coalesced from L2, 16 loads in flight ----
  {
    constexpr int kPerSlice = kOFlat / kOSlices;            // 288
    const int o = tid & (kOD - 1), sl = tid >> 7;
    const float* wl = W + kOlw + size_t(sl) * kPerSlice * kOD + o;
    const float* in = c3 + sl * kPerSlice;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
    for (int k = 0; k < kPerSlice; k += 16) {
      float w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = __ldg(wl + size_t(k + j) * kOD);
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        a0 = fmaf(in[k + j], w[j], a0); a1 = fmaf(in[k + j + 1], w[j + 1], a1);
        a2 = fmaf(in[k + j + 2], w[j + 2], a2); a3 = fmaf(in[k + j + 3], w[j + 3], a3);
      }
    }
    red[sl * kOD + o] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  // ---- context window of trajectory b: shift when full, append the new entry ----
  const int K = p.K;
  const int pos = int(*p.pos);
  const bool full = pos == K - 1;
  const int npos = full ? K - 1 : pos + 1;
  float* rtg = p.w_rtg + size_t(b) * K;
  float* emb = p.w_emb + size_t(b) * K * kOD;
  float* act = p.w_act + size_t(b) * K * 3;
  long long* ts = p.w_ts + size_t(b) * K;
  if (full) {
    // entry i <- entry i + 1 for i < K - 1; every element is read before anything is written
    const float e = tid < (K - 1) * kOD ? emb[kOD + tid] : 0.f;      // K <= 6: (K - 1) * 128 <= 640 <= kOThreads
    float r = 0.f, a = 0.f; long long t = 0;
    if (tid < K - 1) { r = rtg[tid + 1]; t = ts[tid + 1]; }
    if (tid < (K - 1) * 3) a = act[3 + tid];
    __syncthreads();
    if (tid < (K - 1) * kOD) emb[tid] = e;
    if (tid < K - 1) { rtg[tid] = r; ts[tid] = t; }
    if (tid < (K - 1) * 3) act[tid] = a;
  }
  if (tid < kOD) {
    float v = __ldg(W + kOlb + tid);
#pragma unroll
    for (int j = 0; j < kOSlices; ++j) v += red[j * kOD + tid];
    emb[npos * kOD + tid] = tanhf(v);
  }
  if (tid < 3) act[npos * 3 + tid] = 0.f;
  if (tid == 0) {
    rtg[npos] = p.nxt_rtg[b];
    ts[npos] = (*p.t_dev + 1) % p.n_time;
  }
}

constexpr size_t kObserveSmem = sizeof(float) * (size_t(kOE) * kOE + 8 * kO1 * kO1 + 16 * kO2 * kO2 + kOFlat + kOConvW + kOSlices * kOD);

int policy_observe_launch(const float* w, const float* x, int H, int W, const float* nxt_rtg, float* w_rtg, float* w_emb,
                          float* w_act, long long* w_ts, const long long* pos, const long long* t_dev, int B, int K,
                          int n_time, cudaStream_t st) {
  if (B < 1 || K < 1 || K > 6 || H != W || H % kOE != 0 || H / kOE < 1 || n_time < 1) return -1;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(policy_observe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kObserveSmem));
    if (e != cudaSuccess) return int(e);
    attr_done = true;
  }
  ObserveParams p{w, x, H, W, H / kOE, nxt_rtg, w_rtg, w_emb, w_act, w_ts, pos, t_dev, K, n_time};
  policy_observe_kernel<<<B, kOThreads, kObserveSmem, st>>>(p);
  return int(cudaGetLastError());
}

}  // namespace pnp

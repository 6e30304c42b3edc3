// Register-resident radix-16 x radix-16 256-point FFT shared by the FFT-prox kernels, and the optional per-phase cycle
// accounting of the cluster kernel (-DPNP_PROX_PHASE_TIMING, tools/prox_phases.py; the default build contains none of it).
//   * a half-warp owns one row (lane j, register r <-> element j + 16 r); pass A is a 16-point DFT in registers, ONE
//     swizzled shared-memory round trip re-distributes the data, pass B (twiddles + 16-point DFT) leaves element
//     16 r + j in (lane j, register r) - the pattern pass A consumes - so global loads feed pass A directly, the k-space
//     blend and the following inverse transform run register to register, and the last pass stores straight to global
//     memory.  All global accesses are 128-byte coalesced.
#pragma once
#include "common.cuh"
#include "fft_core.cuh"

namespace pnp {

constexpr int kF2N = 256;

#ifdef PNP_PROX_PHASE_TIMING
__device__ unsigned long long g_f2_phase[16];
__device__ __forceinline__ unsigned long long f2_clock() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
#define F2_PHASE_BEGIN() unsigned long long f2_t_prev = f2_clock()
#define F2_PHASE(slot)                                                                 \
  do {                                                                                 \
    const unsigned long long f2_t_now = f2_clock();                                    \
    if (threadIdx.x == 0) atomicAdd(&g_f2_phase[slot], f2_t_now - f2_t_prev);          \
    f2_t_prev = f2_t_now;                                                              \
  } while (0)
#else
#define F2_PHASE_BEGIN() do { } while (0)
#define F2_PHASE(slot) do { } while (0)
#endif

// 16-point forward DFT in registers (4 x 4), natural order in and out.
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  const float c8 = 0.92387953251128675613f, s8 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  float2 t[4][4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    float2 a[4] = {v[b], v[b + 4], v[b + 8], v[b + 12]};
    dft4(a);
#pragma unroll
    for (int q = 0; q < 4; ++q) t[b][q] = a[q];
  }
  // twiddles w16^(b q)
  t[1][1] = cmul(t[1][1], make_float2(c8, -s8));
  t[1][2] = cmul(t[1][2], make_float2(h, -h));
  t[1][3] = cmul(t[1][3], make_float2(s8, -c8));
  t[2][1] = cmul(t[2][1], make_float2(h, -h));
  t[2][2] = make_float2(t[2][2].y, -t[2][2].x);                 // * (-i)
  t[2][3] = cmul(t[2][3], make_float2(-h, -h));
  t[3][1] = cmul(t[3][1], make_float2(s8, -c8));
  t[3][2] = cmul(t[3][2], make_float2(-h, -h));
  t[3][3] = cmul(t[3][3], make_float2(-c8, s8));
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float2 a[4] = {t[0][q], t[1][q], t[2][q], t[3][q]};
    dft4(a);
#pragma unroll
    for (int p = 0; p < 4; ++p) v[q + 4 * p] = a[p];
  }
}

// 256-point forward FFT of one row held by a half-warp.  in: v[r] = x[j + 16 r]; out: v[r] = X[16 r + j].
// `row` = this half-warp's 256-float2 scratch row in shared memory (contents destroyed);
// wtab[t][j] = exp(-2 pi i j m_t / 256) for m_t in {1,2,3,4,8,12} (lanes read consecutive words: conflict free).
__device__ __forceinline__ void fft256_halfwarp(float2 (&v)[16], float2* row, const float2* wtab, int j) {
  dft16(v);
  // pass A store: y[16 j + q] = V[q], 16-byte vectors, chunk m -> m ^ (j & 7)
  {
    float4* dst = reinterpret_cast<float4*>(row + 16 * j);
#pragma unroll
    for (int m = 0; m < 8; ++m) dst[m ^ (j & 7)] = make_float4(v[2 * m].x, v[2 * m].y, v[2 * m + 1].x, v[2 * m + 1].y);
  }
  __syncwarp();
  // pass B load: u[r] = y[16 r + j]
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = row[16 * r + (j ^ ((r & 7) << 1))];
  __syncwarp();
  // twiddles w256^(j r): six table look-ups, nine products
  {
    const float2 w1 = wtab[j], w2 = wtab[16 + j], w3 = wtab[32 + j];
    const float2 w4 = wtab[48 + j], w8 = wtab[64 + j], w12 = wtab[80 + j];
    v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3); v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], cmul(w4, w1)); v[6] = cmul(v[6], cmul(w4, w2)); v[7] = cmul(v[7], cmul(w4, w3));
    v[8] = cmul(v[8], w8);
    v[9] = cmul(v[9], cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
    v[12] = cmul(v[12], w12);
    v[13] = cmul(v[13], cmul(w12, w1)); v[14] = cmul(v[14], cmul(w12, w2)); v[15] = cmul(v[15], cmul(w12, w3));
  }
  dft16(v);
}

}  // namespace pnp

// FFT-prox + dual update for sampling masks that depend on the column index only (Cartesian undersampling with fully
// sampled k-space columns - BASELINE.json configs 2 and 4): single launch, no inter-CTA communication.  256x256 has the
// specialised kernel below, the other power-of-two sizes fftprox_rows_generic_kernel at the end of this file.
//
// With orthonormal 1-D transforms Fr (along a row) and Fc (along a column), F = Fc Fr, and a mask m(j) that does not
// depend on the row index, m commutes with Fc, so the reference step (evaluation/env.py:87-93)
//     z = F^-1 [ blend(F w) ],   blend(G)[m] = (mu G + y0)[m] / (1 + mu),   w = x + u
// is identical to
//     z = Fr^-1 [ blend_row(Fr w) ],   blend_row(G)[i, j] = m(j) ? (mu G[i, j] + Yt[i, j]) / (1 + mu) : G[i, j],
//     Yt = Fc^-1 y0                                  (one column transform per TRAJECTORY, done in pnp_prox_prepare)
// i.e. every image row is independent: load x, u -> 256-point FFT -> blend -> inverse FFT -> z, u', v', with the
// reference's centring folded into sign flips exactly as in the general kernels (D = (-1)^(i+j) on load and store,
// s*D folded into Yt).  Half the arithmetic of the 2-D path, no transposes, no cluster: the kernel streams at HBM speed.
// pnp_prox_prepare decides per batch whether the masks have this structure (device-side flag, no host round trip);
// pnp_prox_dual_prepared launches this kernel and the general cluster kernel, and the one whose case it is not exits at
// once.  A half-warp owns one row (register-resident radix-16 x radix-16 FFT, fft256_reg.cuh).
#pragma once
#include "common.cuh"
#include "fft_core.cuh"
#include "fft256_reg.cuh"

namespace pnp {

struct SepParams {
  const float* x;
  const float2* u_in;
  const float2* yt;         // [B][i][j]  Fc^-1 (s*D.y0)
  const uint16_t* mpack;    // [B or 1][16]: bit r of entry j0 = m(16 r + j0)
  int mask_per_image;       // 1: one mask per image, 0: one mask for the batch
  const int* flag;          // != 0: the masks are column-only (this kernel's case)
  const float* mu;
  int mu_stride;
  float2* z_out;
  float2* u_out;
  float* v_out;
  int rows_total;           // B * 256
  int prefetch_yt;          // 1: pull the row of Yt into L2 while the first transform runs
  const uint8_t* active;    // optional [B]: 0 = skip the image (its z, u, v stay untouched)
};

constexpr int kSepThreads = 256;                                  // 16 half-warps = 16 rows in flight per CTA
#ifndef PNP_SEP_EPI
#define PNP_SEP_EPI 8
#endif
constexpr int kSepEpi = PNP_SEP_EPI;                              // elements whose u, x loads are issued ahead of their stores
constexpr size_t kSepSmem = size_t(16) * kF2N * sizeof(float2) + 96 * sizeof(float2);

__global__ void __launch_bounds__(kSepThreads, 2) fftprox_rows256_kernel(const SepParams p) {
  if (*p.flag == 0) return;
  extern __shared__ float2 sep_sm[];
  float2* w256 = sep_sm + size_t(16) * kF2N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, j = lane & 15;
  if (threadIdx.x < 96) {
    const int t = threadIdx.x >> 4, jj = threadIdx.x & 15;
    const int m = (t < 4) ? t + 1 : (t == 4 ? 8 : 12);
    w256[threadIdx.x] = g_tw512[2 * jj * m];
  }
  __syncthreads();
  float2* row = sep_sm + (warp * 2 + half) * kF2N;
  const float inv = 1.0f / 16.0f;                                 // 1/sqrt(W), applied after each transform
  for (int r0 = blockIdx.x * 16 + warp * 2 + half; r0 < p.rows_total; r0 += gridDim.x * 16) {
    const int b = r0 >> 8, i = r0 & 255;
    if (p.active && p.active[b] == 0) continue;                   // uniform over the half-warp (one row)
    const size_t g0 = size_t(r0) * kF2N + j;
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float2 uu = p.u_in[g0 + 16 * r];                       // plain loads: u_out may alias u_in (include/pnp_b200.h)
      const float xx = __ldg(p.x + g0 + 16 * r);
      v[r] = make_float2(xx + uu.x, uu.y);
    }
    // the blend reads Yt under the mask right after the transform: lane j pulls 128-byte line j of the row into L2 now
    // (prefetching the half-warp's NEXT row of x and u the same way was measured: -5 % at B = 64, -20 % at B >= 256)
    if (p.prefetch_yt) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.yt + size_t(r0) * kF2N + 16 * j));
    const uint32_t mbits = __ldg(p.mpack + (p.mask_per_image ? b * 16 : 0) + j);
    const float mu = __ldg(p.mu + size_t(b) * p.mu_stride);
    const float inv1mu = 1.f / (1.f + mu);
    const bool neg = (i + j) & 1;                                 // D = (-1)^(i + col), col = j + 16 r has the parity of j
    if (neg) {
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = make_float2(-v[r].x, -v[r].y);
    }
    fft256_halfwarp(v, row, w256, j);                             // v[r] = G[16 r + j] * 16
    const float2* yp = p.yt + g0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float2 Z = make_float2(v[r].x * inv, v[r].y * inv);
      if ((mbits >> r) & 1u) {
        const float2 y = __ldg(yp + 16 * r);
        Z.x = (mu * Z.x + y.x) * inv1mu;
        Z.y = (mu * Z.y + y.y) * inv1mu;
      }
      v[r] = make_float2(Z.x, -Z.y);                              // conj: forward FFT == inverse
    }
    fft256_halfwarp(v, row, w256, j);
    const float sg = neg ? -inv : inv;
    // u_in is read with plain loads (u_out may alias it), which the compiler will not move across the stores below: the
    // loads of kSepEpi elements are issued explicitly before those elements' stores (one load -> store -> load chain per
    // element cost 25 % of the kernel: 32 -> 40 us at B = 64)
#pragma unroll
    for (int r4 = 0; r4 < 16; r4 += kSepEpi) {
      float2 uu[kSepEpi];
      float xx[kSepEpi];
#pragma unroll
      for (int q = 0; q < kSepEpi; ++q) {
        uu[q] = p.u_in[g0 + 16 * (r4 + q)];
        xx[q] = __ldg(p.x + g0 + 16 * (r4 + q));
      }
#pragma unroll
      for (int q = 0; q < kSepEpi; ++q) {
        const size_t g = g0 + 16 * (r4 + q);
        const float2 zz = make_float2(sg * v[r4 + q].x, -sg * v[r4 + q].y);
        const float2 un = make_float2(uu[q].x + xx[q] - zz.x, uu[q].y - zz.y);
        p.z_out[g] = zz;
        p.u_out[g] = un;
        if (p.v_out) p.v_out[g] = zz.x - un.x;
      }
    }
    __syncwarp();
  }
}

// One CTA per mask: is mask[i][j] == mask[0][j] for every row?  Clears *flag otherwise; always writes the packed row mask.
// Row mask storage per mask: 256 x 256: 16 packed uint16 (32 B, layout of fftprox_rows256_kernel); every other shape
// (including H x 256 with H != 256, which runs fftprox_rows_generic_kernel<256>): W plain bytes.
__host__ __device__ inline size_t sep_rowmask_stride(int H, int W) { return (H == 256 && W == 256) ? 32 : size_t(W); }

constexpr int kSepCheckSlices = 32;        // CTAs per mask (grid.y): each compares H / 32 rows with row 0
__global__ void __launch_bounds__(256) sep_check_kernel(const uint8_t* __restrict__ mask, long long bstride, int H, int W,
                                                        uint16_t* __restrict__ mpack, int* flag) {
  __shared__ __align__(16) uint8_t m0[512];
  __shared__ int bad;
  const uint8_t* mk = mask + size_t(blockIdx.x) * bstride;
  if (threadIdx.x == 0) bad = 0;
  for (int jx = threadIdx.x; jx < W; jx += blockDim.x) m0[jx] = mk[jx] ? 1 : 0;
  __syncthreads();
  // rows [r0, r1) of this slice, 16 bytes per thread and step (W is a multiple of 32, the mask 16-byte aligned)
  const int rows = (H + kSepCheckSlices - 1) / kSepCheckSlices;
  const int r0 = blockIdx.y * rows, r1 = min(H, r0 + rows);
  const int vec_per_row = W / 16;
  int mism = 0;
  for (int e = threadIdx.x; e < (r1 - r0) * vec_per_row; e += blockDim.x) {
    const int r = r0 + e / vec_per_row, c = e % vec_per_row;
    const uint4 a = *reinterpret_cast<const uint4*>(mk + size_t(r) * W + c * 16);
    const uint4 b = *reinterpret_cast<const uint4*>(m0 + c * 16);
    // bytes are compared as booleans (any non-zero value is "sampled", as .to(torch.bool) in reference env.py:64)
    const uint32_t w[4] = {a.x, a.y, a.z, a.w}, q[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t nz = w[k] | (w[k] >> 4); nz |= nz >> 2; nz |= nz >> 1; nz &= 0x01010101u;   // per byte: != 0
      mism |= (nz != q[k]);
    }
  }
  if (mism) bad = 1;
  __syncthreads();
  if (threadIdx.x == 0 && bad) atomicExch(flag, 0);
  if (blockIdx.y != 0) return;
  if (H == 256 && W == 256) {
    if (threadIdx.x < 16) {
      uint32_t bits = 0;
      for (int r = 0; r < 16; ++r) bits |= uint32_t(m0[16 * r + threadIdx.x]) << r;
      mpack[blockIdx.x * 16 + threadIdx.x] = uint16_t(bits);
    }
  } else {
    uint8_t* mrow = reinterpret_cast<uint8_t*>(mpack) + size_t(blockIdx.x) * W;
    for (int jx = threadIdx.x; jx < W; jx += blockDim.x) mrow[jx] = m0[jx];
  }
}

// Row-only prox for the other power-of-two sizes (32..512): same algebra as fftprox_rows256_kernel on the generic
// warp-per-row shared-memory FFT (fft_core.cuh).  A warp owns FftPlan<N>::G consecutive rows of one image.
struct SepGenParams {
  const float* x;
  const float2* u_in;
  const float2* yt;         // [B][H][W]
  const uint8_t* mrow;      // [B or 1][W]
  int mask_per_image;
  const int* flag;
  const float* mu;
  int mu_stride;
  float2* z_out;
  float2* u_out;
  float* v_out;
  int H, groups_total;      // B * H / G row groups
  const uint8_t* active;    // optional [B]: 0 = skip the image
};

template <int N>
__global__ void __launch_bounds__(256) fftprox_rows_generic_kernel(const SepGenParams p) {
  if (*p.flag == 0) return;
  constexpr int G = FftPlan<N>::G;
  constexpr int P = fft_pitch(N);
  __shared__ float2 tw[kTwTotal];
  extern __shared__ float2 gen_rows_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fft_load_twiddles(tw, g_tw512);
  __syncthreads();
  float2* mine = gen_rows_smem + warp * G * P;
  const float inv = 1.0f / sqrtf(float(N));
  for (int rg = blockIdx.x * 8 + warp; rg < p.groups_total; rg += gridDim.x * 8) {
    const int r0 = rg * G;
    const int b = r0 / p.H;
    if (p.active && p.active[b] == 0) continue;                    // uniform over the warp (a group never spans images)
    const float mu = __ldg(p.mu + size_t(b) * p.mu_stride);
    const float inv1mu = 1.f / (1.f + mu);
    const uint8_t* mrow = p.mrow + (p.mask_per_image ? size_t(b) * N : 0);
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      const int i = (r0 + g) % p.H;
      const size_t base = size_t(r0 + g) * N;
      for (int j = lane; j < N; j += 32) {
        const float2 uu = p.u_in[base + j];                        // plain loads: u_out may alias u_in
        float2 v = make_float2(__ldg(p.x + base + j) + uu.x, uu.y);
        if ((i + j) & 1) { v.x = -v.x; v.y = -v.y; }
        mine[g * P + fpad(j)] = v;
      }
    }
    __syncwarp();
    fft_warp_rows<N>(mine, P, tw, lane);
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      const size_t base = size_t(r0 + g) * N;
      for (int j = lane; j < N; j += 32) {
        float2 Z = mine[g * P + fpad(j)];
        Z.x *= inv; Z.y *= inv;
        if (mrow[j]) {
          const float2 y = __ldg(p.yt + base + j);
          Z.x = (mu * Z.x + y.x) * inv1mu;
          Z.y = (mu * Z.y + y.y) * inv1mu;
        }
        mine[g * P + fpad(j)] = make_float2(Z.x, -Z.y);
      }
    }
    __syncwarp();
    fft_warp_rows<N>(mine, P, tw, lane);
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      const int i = (r0 + g) % p.H;
      const size_t base = size_t(r0 + g) * N;
      // u_in is read with plain loads (u_out may alias it), which stay behind the preceding stores: the loads of KB
      // elements are issued ahead of their stores by hand (as in fftprox_rows256_kernel)
      constexpr int KB = (N / 32 < 4) ? N / 32 : 4;
      for (int j0 = lane; j0 < N; j0 += 32 * KB) {
        float2 uu[KB];
        float xx[KB];
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          uu[q] = p.u_in[base + j0 + 32 * q];
          xx[q] = __ldg(p.x + base + j0 + 32 * q);
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          const int j = j0 + 32 * q;
          const float2 t = mine[g * P + fpad(j)];
          const float sg = ((i + j) & 1) ? -inv : inv;
          const float2 zz = make_float2(sg * t.x, -sg * t.y);
          const float2 un = make_float2(uu[q].x + xx[q] - zz.x, uu[q].y - zz.y);
          p.z_out[base + j] = zz;
          p.u_out[base + j] = un;
          if (p.v_out) p.v_out[base + j] = zz.x - un.x;
        }
      }
    }
    __syncwarp();
  }
}

template <int N> static void launch_sep_generic(const SepGenParams& p, int num_sms, cudaStream_t st) {
  constexpr int G = FftPlan<N>::G;
  const size_t smem = size_t(8) * G * fft_pitch(N) * sizeof(float2);
  int grid = (p.groups_total + 7) / 8;
  const int cap = num_sms * 4;
  if (grid > cap) grid = cap;
  fftprox_rows_generic_kernel<N><<<grid, 256, smem, st>>>(p);
}

static int launch_sep(const SepParams& p, int num_sms, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(fftprox_rows256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSepSmem));
    if (e != cudaSuccess) return int(e);
    attr_done = true;
  }
  int grid = (p.rows_total + 15) / 16;
  const int cap = num_sms * 2;
  if (grid > cap) grid = cap;
  // (a programmatic-serialization launch was measured here: the successor's early-resident CTAs cost the row kernel
  // 30 % at B = 1024, profiles/r02_prox_cl_steps.txt)
  fftprox_rows256_kernel<<<grid, kSepThreads, kSepSmem, st>>>(p);
  return int(cudaGetLastError());
}

}  // namespace pnp

// Single-launch FFT-prox + dual update for square power-of-two images that fit a thread-block cluster's
// shared memory: 128x128 in one CTA, 256x256 in a 4-CTA cluster (64 rows each), DSMEM transposes.
//
// Per image (one cluster, persistent over the batch) every input byte is read from HBM once and every output
// byte written once (x and u are re-read for the epilogue, normally from L2):
//   P1  load rows       w = D.(x + u)                                   global -> smem   (coalesced)
//   P2  row FFTs        warp per row, in place                           smem
//   P3  transpose       all-to-all over the cluster through registers    smem -> DSMEM
//   P4  column FFTs                                                      smem
//   P5  blend           Z = (mu Z + sD y0)/(1+mu) under the mask; conj   smem (+ y0, mask from global, coalesced)
//   P6  column FFTs     (inverse via conjugation)                        smem
//   P7  transpose back                                                   smem -> DSMEM
//   P8  row FFTs                                                         smem
//   P9  epilogue        z = D.conj(.)/sqrt(HW); u' = u + x - z; v' = Re(z - u')   smem -> global (coalesced)
// Shared-memory rows are padded (fft_core.cuh) and the row pitch is odd in float2 units (2 mod 32 in words), so
// both the row-wise and the column-wise (transposed) walks are bank-conflict free.
// Included by fftprox.cu (same translation unit as the twiddle table).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "fft_core.cuh"

namespace pnp {
namespace cg = cooperative_groups;

struct FusedProxParams {
  const float* x;
  const float2* u_in;
  const float2* y0;
  const uint8_t* mask;
  long long mask_bstride;
  const float* mu;
  int mu_stride;
  float2* z_out;
  float2* u_out;
  float* v_out;
  int B;
  float sgn;          // (-1)^((H+W)/2)
};

constexpr int kFusedThreads = 512;

template <int N, int CL>
struct FusedCfg {
  static constexpr int R = N / CL;                    // rows (then columns) owned by one CTA
  static constexpr int P = fft_pitch(N) + 1;          // float2 per padded row; odd -> conflict-free transposed walks
  static constexpr int EPT = N * R / kFusedThreads;   // elements per thread in the all-to-all
  static constexpr size_t SMEM = size_t(R) * P * sizeof(float2) + kTwTotal * sizeof(float2);
  static_assert(N * R % kFusedThreads == 0, "tile must divide over the CTA");
};

template <int N, int CL>
__device__ __forceinline__ void fused_sync() {
  if constexpr (CL == 1) __syncthreads();
  else cg::this_cluster().sync();
}

// Cluster-wide transpose, pull style: after the call this CTA holds, for c in [0,R) and row in [0,N),
// tile[c][row] = (element (row % R, rank*R + c) of CTA row / R).  Remote reads walk contiguous runs of the
// peer's rows (coalesced DSMEM traffic); the strided side of the transpose is the local, conflict-free store.
// Everything is staged in registers so the transpose is in place.
template <int N, int CL>
__device__ __forceinline__ void transpose_exchange(float2* tile, unsigned rank) {
  using Cfg = FusedCfg<N, CL>;
  constexpr int R = Cfg::R, P = Cfg::P, EPT = Cfg::EPT;
  float2 v[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int e = i * kFusedThreads + threadIdx.x;
    const int row = e / R, c = e % R;
    const int d = row / R, a = row % R;
    const float2* src = tile;
    if constexpr (CL > 1) src = cg::this_cluster().map_shared_rank(tile, d);
    v[i] = src[a * P + fpad(int(rank) * R + c)];
  }
  fused_sync<N, CL>();          // every CTA has pulled what it needs; tiles may now be overwritten
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int e = i * kFusedThreads + threadIdx.x;
    const int row = e / R, c = e % R;
    tile[c * P + fpad(row)] = v[i];
  }
  __syncthreads();
}

template <int N, int CL>
__global__ void __launch_bounds__(kFusedThreads, 1) fftprox_fused_kernel(const FusedProxParams p) {
  using Cfg = FusedCfg<N, CL>;
  constexpr int R = Cfg::R, P = Cfg::P, G = FftPlan<N>::G;
  constexpr int NW = kFusedThreads / 32;
  extern __shared__ float2 fsm[];
  float2* tile = fsm;
  float2* tw = fsm + size_t(R) * P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned rank = 0;
  int cluster_id = blockIdx.x, n_clusters = gridDim.x;
  if constexpr (CL > 1) {
    rank = cg::this_cluster().block_rank();
    cluster_id = blockIdx.x / CL;
    n_clusters = gridDim.x / CL;
  }
  fft_load_twiddles(tw, g_tw512);
  const float inv = rsqrtf(float(N) * float(N));
  const int row0 = int(rank) * R;

  for (int b = cluster_id; b < p.B; b += n_clusters) {
    const size_t img = size_t(b) * N * N;
    // ---- P1: load my rows, w = D.(x+u); two pixels per lane (16-byte u loads, 8-byte x loads)
#pragma unroll 4
    for (int e = threadIdx.x; e < R * N / 2; e += kFusedThreads) {
      const int r = e / (N / 2), j = (e % (N / 2)) * 2;
      const size_t g = img + size_t(row0 + r) * N + j;
      const float4 uu = __ldg(reinterpret_cast<const float4*>(p.u_in + g));
      const float2 xx = __ldg(reinterpret_cast<const float2*>(p.x + g));
      const float sgn0 = ((row0 + r) & 1) ? -1.f : 1.f;       // j is even: D = +-1 for pixel j, -+1 for j+1
      float2* dst = tile + r * P + fpad(j);
      dst[0] = make_float2(sgn0 * (xx.x + uu.x), sgn0 * uu.y);
      dst[1] = make_float2(-sgn0 * (xx.y + uu.z), -sgn0 * uu.w);
    }
    __syncthreads();
    // ---- P2: row FFTs
    for (int r = warp * G; r < R; r += NW * G) fft_warp_rows<N>(tile + r * P, P, tw, lane);
    __syncthreads();
    // ---- P3: transpose -> I now own columns [row0, row0+R), stored as tile[c][k_i]
    if constexpr (CL > 1) cg::this_cluster().sync();   // peers' row FFTs are complete before anyone pulls
    transpose_exchange<N, CL>(tile, rank);
    // ---- P4: column FFTs
    for (int r = warp * G; r < R; r += NW * G) fft_warp_rows<N>(tile + r * P, P, tw, lane);
    __syncthreads();
    // ---- P5: blend in k-space, conjugate for the inverse
    {
      const float mu = p.mu[size_t(b) * p.mu_stride];
      const float inv1mu = 1.f / (1.f + mu);
      const uint8_t* mk = p.mask + size_t(b) * p.mask_bstride;
#pragma unroll 4
      for (int e = threadIdx.x; e < R * N; e += kFusedThreads) {
        const int ki = e / R, c = e % R;
        const int kj = row0 + c;
        const size_t g = size_t(ki) * N + kj;
        float2 Z = tile[c * P + fpad(ki)];
        Z.x *= inv; Z.y *= inv;
        if (mk[g]) {
          const float2 y = p.y0[img + g];
          const float sg = ((ki + kj) & 1) ? -p.sgn : p.sgn;
          Z.x = (mu * Z.x + sg * y.x) * inv1mu;
          Z.y = (mu * Z.y + sg * y.y) * inv1mu;
        }
        tile[c * P + fpad(ki)] = make_float2(Z.x, -Z.y);
      }
    }
    __syncthreads();
    // ---- P6: column FFTs (inverse)
    for (int r = warp * G; r < R; r += NW * G) fft_warp_rows<N>(tile + r * P, P, tw, lane);
    __syncthreads();
    // ---- P7: transpose back -> rows again
    if constexpr (CL > 1) cg::this_cluster().sync();
    transpose_exchange<N, CL>(tile, rank);
    // ---- P8: row FFTs (inverse)
    for (int r = warp * G; r < R; r += NW * G) fft_warp_rows<N>(tile + r * P, P, tw, lane);
    __syncthreads();
    // ---- P9: epilogue, two pixels per lane
#pragma unroll 4
    for (int e = threadIdx.x; e < R * N / 2; e += kFusedThreads) {
      const int r = e / (N / 2), j = (e % (N / 2)) * 2;
      const size_t g = img + size_t(row0 + r) * N + j;
      const float s0 = ((row0 + r) & 1) ? -inv : inv;
      const float2* src = tile + r * P + fpad(j);
      const float2 t0 = src[0], t1 = src[1];
      const float2 z0 = make_float2(s0 * t0.x, -s0 * t0.y);      // D . conj(.) / sqrt(HW)
      const float2 z1 = make_float2(-s0 * t1.x, s0 * t1.y);
      const float4 uu = __ldg(reinterpret_cast<const float4*>(p.u_in + g));
      const float2 xx = __ldg(reinterpret_cast<const float2*>(p.x + g));
      const float4 un = make_float4(uu.x + xx.x - z0.x, uu.y - z0.y, uu.z + xx.y - z1.x, uu.w - z1.y);
      *reinterpret_cast<float4*>(p.z_out + g) = make_float4(z0.x, z0.y, z1.x, z1.y);
      *reinterpret_cast<float4*>(p.u_out + g) = un;
      if (p.v_out) *reinterpret_cast<float2*>(p.v_out + g) = make_float2(z0.x - un.x, z1.x - un.z);
    }
    __syncthreads();   // tile is reused by the next image's P1
  }
  if constexpr (CL > 1) cg::this_cluster().sync();   // no CTA exits while a peer may still write into it
}

template <int N, int CL>
static int launch_fused(const FusedProxParams& p, int num_sms, cudaStream_t st) {
  using Cfg = FusedCfg<N, CL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(fftprox_fused_kernel<N, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(Cfg::SMEM));
    if (e != cudaSuccess) return int(e);
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(num_sms / CL * CL);
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent grid = the number of clusters that can be co-resident (GPC boundaries can strand SMs for CL=4)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fftprox_fused_kernel<N, CL>, &cfg) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      n = num_sms / CL;
    }
    max_clusters = n;
  }
  int clusters = max_clusters < p.B ? max_clusters : p.B;
  if (clusters < 1) clusters = 1;
  cfg.gridDim = dim3(clusters * CL);
  return int(cudaLaunchKernelEx(&cfg, fftprox_fused_kernel<N, CL>, p));
}

}  // namespace pnp

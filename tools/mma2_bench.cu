// cta_group::2 (CTA-pair) tcgen05.mma: semantics check + issue rate.  Two CTAs of a cluster each hold 128 rows of A and
// HALF of B (N/2 rows) in their own shared memory (128-byte rows, SWIZZLE_128B, K-major); the leader issues M=256 MMAs.
// Prints whether D matches the hypothesis  D[cta*128 + r][n] = sum_k A_cta[r][k] * Bfull[n][k],  Bfull = [B_cta0 ; B_cta1],
// then the cycles per MMA for back-to-back issue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dt4image_restoration_b200/csrc tools/mma2_bench.cu -o tools/mma2_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cooperative_groups.h>
#include "common.cuh"
using namespace pnp;
namespace cg = cooperative_groups;

__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__host__ __device__ inline float aval(int cta, int r, int k) { return float(((r * 3 + k * 5 + 7 * cta) % 7) - 3); }
__host__ __device__ inline float bval(int cta, int n, int k) { return float(((n * 2 + k + 3 * cta) % 5) - 2); }

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(float* out, long long* clk, int n_timing, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, sbar[4];
  __shared__ uint32_t tslot;
  cg::cluster_group cl = cg::this_cluster();
  const int cta = int(cl.block_rank());
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* a_s = smem;                       // 128 rows x 128 B
  uint8_t* b_s = smem + 16384;               // N/2 rows x 128 B
  // fill operands, 128-byte swizzle: 16-byte chunk index ^= (row & 7)
  for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
    const int r = e / 64, k = e % 64;
    const uint32_t off = uint32_t(r * 128 + (k / 8) * 16);
    const uint32_t phys = off ^ (((off >> 7) & 7) << 4);
    reinterpret_cast<__nv_bfloat16*>(a_s + phys)[k % 8] = __float2bfloat16_rn(aval(cta, r, k));
  }
  for (int e = threadIdx.x; e < (N / 2) * 64; e += blockDim.x) {
    const int n = e / 64, k = e % 64;
    const uint32_t off = uint32_t(n * 128 + (k / 8) * 16);
    const uint32_t phys = off ^ (((off >> 7) & 7) << 4);
    reinterpret_cast<__nv_bfloat16*>(b_s + phys)[k % 8] = __float2bfloat16_rn(bval(cta, n, k));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&sbar[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc2(&tslot, 512); tmem_relinquish2(); }
  fence_proxy_async_smem();
  tc_fence_before();
  cl.sync();
  tc_fence_after();
  const uint32_t tb = tslot;
  constexpr uint32_t idesc = umma_idesc_bf16(256, N);
  constexpr uint32_t hi = (uint32_t(1024) >> 4) | (1u << 14) | (2u << 29);
  const uint32_t a_lo = (smem_u32(a_s) >> 4) | (1u << 16);
  const uint32_t b_lo = (smem_u32(b_s) >> 4) | (1u << 16);
  if (cta == 0 && warp == 0) {
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma2_bf16_ss(tb, a_lo + k * 2, hi, b_lo + k * 2, hi, idesc, k ? 1u : 0u);
      tc_commit2(&bar, 3);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // every CTA dumps its 128 lanes x N columns
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tb + (uint32_t(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t(cta) * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  cl.sync();
  tc_fence_after();
  // timing: n_timing MMAs from the leader.  mode 0: back to back, one issuer.  mode 1: one issuer, per group of 4 MMAs a
  // (passing) barrier wait before and a multicast commit after.  mode 2: the same structure from TWO issuing warps
  // (warps 0 and 1, own accumulators), n_timing / 2 MMAs each.
  if (cta == 0 && (warp == 0 || (mode == 2 && warp == 1))) {
    const int me = warp;
    const int mine = mode == 2 ? n_timing / 2 : n_timing;
    const long long t0 = clock64();
    for (int g = 0; g < mine / 4; ++g) {
      if (mode) { mbar_wait(&sbar[0], 1); tc_fence_after(); }
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 4; ++i) umma2_bf16_ss(tb + me * N, a_lo + i * 2, hi, b_lo + i * 2, hi, idesc, 1u);
        if (mode) tc_commit2(&sbar[1 + me], 3);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit2(me ? &sbar[3] : &bar, 3);
    __syncwarp();
    mbar_wait(me ? &sbar[3] : &bar, me ? 0 : 1);
    if (lane == 0 && me == 0) clk[blockIdx.x / 2] = clock64() - t0;
  } else if (warp == 0) {
    mbar_wait(&bar, 1);
  }
  tc_fence_before();
  cl.sync();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tb, 512); }
}

template <int N> static void run(int mode) {
  const int pairs = 74;
  float* d; long long* c;
  cudaMalloc(&d, size_t(256) * N * sizeof(float)); cudaMalloc(&c, pairs * sizeof(long long));
  cudaMemset(d, 0xff, size_t(256) * N * sizeof(float));
  const int smem = 1024 + 16384 + (N / 2) * 128;
  cudaFuncSetAttribute(k2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_timing = 4096;
  k2<N><<<2 * pairs, 128, smem>>>(d, c, n_timing, mode);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> h(size_t(256) * N); std::vector<long long> hc(pairs);
  cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
  cudaMemcpy(hc.data(), c, pairs * sizeof(long long), cudaMemcpyDeviceToHost);
  int bad = 0; double maxerr = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += double(aval(m / 128, m % 128, k)) * double(bval(n / (N / 2), n % (N / 2), k));
      const double err = fabs(ref - h[size_t(m) * N + n]);
      if (err > 1e-3) { if (bad < 4) printf("  mismatch D[%d][%d] = %g, expected %g\n", m, n, h[size_t(m) * N + n], ref); ++bad; }
      maxerr = err > maxerr ? err : maxerr;
    }
  long long mx = 0; for (auto v : hc) mx = v > mx ? v : mx;
  static const char* mn[] = {"back to back", "wait + multicast commit per 4 MMAs", "two issuing warps, wait + commit per 4"};
  printf("cta_group::2 M=256 N=%3d [%s]: %s (%d mismatches)  %6.1f clk/MMA  (math rate %5.1f)  %s\n", N, mn[mode], bad ? "MISMATCH" : "D correct", bad,
         double(mx) / n_timing, N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(c);
}

int main() { for (int m = 0; m < 3; ++m) { run<64>(m); run<128>(m); run<256>(m); } return 0; }

// tcgen05.mma issue rate under concurrent shared-memory / L1 traffic (what the conv kernel's MMA warp sees):
//   BG 0: MMAs alone      BG 1: + bulk-copy (TMA engine) fills of shared memory from L2
//   BG 2: + epilogue-like strided 16-byte global stores   BG 3: + the same bytes stored fully coalesced
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dt4image_restoration_b200/csrc tools/mma_bench2.cu -o tools/mma_bench2
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace pnp;

template <int N, int BG, int FL = 7>
__global__ void __launch_bounds__(384, 1) k(int n_mma, const uint8_t* gsrc, uint4* gdst, long long* out, int fill_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, fbar[4];
  __shared__ uint32_t tslot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&fbar[i], 1); done = 0; fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  constexpr int ROWB = 128;
  if (BG == 7 && (warp == 0 || warp == 3)) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    constexpr uint32_t a_hi = (uint32_t(18 * ROWB) >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t b_hi = (uint32_t(8 * ROWB) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (1u << 16);
    const uint32_t b_lo0 = (smem_u32(smem + 48 * 1024) >> 4) | (1u << 16);
    const int me = warp == 0 ? 0 : 1;
    const long long t0 = clock64();
    for (int g = 0; g < n_mma / 16; ++g) {
      if (FL & 1) mbar_wait(&fbar[0], 1);
      tc_fence_after();
      const int tap = g % 9;
      const uint32_t a = a_lo0 + uint32_t(((tap / 3) * 18 + tap % 3) * ROWB) / 16;
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          umma_bf16_ss2(tb + (i & 1) * N, a + (i >> 1) * 2 + (i & 1) * 64, a_hi, b_lo0 + (i >> 1) * 2, b_hi, idesc, 1u);
        if (FL & 4) tc_commit(&fbar[1 + me]);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(me ? &fbar[3] : &bar);
    __syncwarp();
    mbar_wait(me ? &fbar[3] : &bar, 0);
    const long long t1 = clock64();
    __syncwarp();
    if (lane == 0 && me == 1) out[blockIdx.x * 2 + 1] = t1 - t0;
    if (lane == 0 && me == 0) { out[blockIdx.x * 2] = t1 - t0; }
  } else if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    constexpr uint32_t a_hi = (uint32_t(18 * ROWB) >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t b_hi = (uint32_t(8 * ROWB) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (1u << 16);
    const uint32_t b_lo0 = (smem_u32(smem + 48 * 1024) >> 4) | (1u << 16);
    const long long t0 = clock64();
    if (BG == 6) {
      constexpr int G = (FL & 32) ? 16 : 8;
      volatile int* cnt = &done;
      int c_next = 0;
      const bool single = (FL & 16) != 0;
      if (!single || lane == 0) {
        for (int g = 0; g < n_mma / G; ++g) {
          if (FL & 1) mbar_wait(&fbar[0], 1);
          if (FL & 2) { while (c_next != 0) c_next = *cnt; c_next = *cnt; }
          const int tap = g % 9;
          const uint32_t a = a_lo0 + uint32_t(((tap / 3) * 18 + tap % 3) * ROWB) / 16;
          if (single || elect_one()) {
#pragma unroll
            for (int i = 0; i < G; ++i)
              umma_bf16_ss2(tb + (i & 1) * N, a + ((i >> 1) & 3) * 2 + (i & 1) * 64, a_hi, b_lo0 + ((i >> 1) & 3) * 2, b_hi, idesc, 1u);
            if (FL & 4) tc_commit(&fbar[1 + (g & 1)]);
            if ((FL & 8) && (g & 1)) tc_commit(&fbar[1 + ((g >> 1) & 1)]);
          }
          if (!single) __syncwarp();
        }
      }
      __syncwarp();
    } else if (BG == 5) {
      // the conv kernel's loop structure: per tap a barrier poll + fence, 8 MMAs, a commit
      uint32_t okn = (FL & 64) ? 0u : 1u;
      long long t_issue = 0, t_wait = 0;
      for (int g = 0; g < n_mma / 8; ++g) {
        const long long ta = clock64();
        if (FL & 8) { if (!okn) mbar_wait(&fbar[0], 1); okn = mbar_test(&fbar[0], 1); }
        if (FL & 1) mbar_wait(&fbar[0], 1);
        const long long tb2 = clock64();
        t_wait += tb2 - ta;          // phase-1 wait on a fresh barrier returns immediately
        if (FL & 2) tc_fence_after();
        const int tap = g % 9;
        const uint32_t a = a_lo0 + uint32_t(((tap / 3) * 18 + tap % 3) * ROWB) / 16;
        if (FL & 64) {
          // readiness published by another warp as a plain shared-memory counter: LDS issued before the MMAs, used after
          volatile int* cnt = &done;   // stays 0 during the run
          if (okn != 0) { while (*cnt != 0) {} }
          const int c_next = *cnt;
          if (elect_one()) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              umma_bf16_ss2(tb + (i & 1) * N, a + (i >> 1) * 2 + (i & 1) * 64, a_hi, b_lo0 + (i >> 1) * 2, b_hi, idesc, 1u);
            if ((FL & 4) && (g & 1)) tc_commit(&fbar[1 + ((g >> 1) & 1)]);
          }
          okn = uint32_t(c_next);
          __syncwarp();
        } else if (FL & 16) {
          uint32_t r0 = 0, r1 = 0, r2 = 0;
          if (!okn) mbar_wait(&fbar[0], 1);
          if (elect_one())
            umma_tap_block<4>(tb, a, a_hi, b_lo0, b_hi, idesc, 1u, smem_u32(&fbar[0]), 1, smem_u32(&fbar[0]), 1, smem_u32(&fbar[0]), 1,
                              smem_u32(&fbar[1 + (g & 1)]), (FL & 32) ? smem_u32(&fbar[3]) : 0u, 64, N, r0, r1, r2);
          okn = __any_sync(0xffffffffu, r0 & r1 & r2);
        } else if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            umma_bf16_ss2(tb + (i & 1) * N, a + (i >> 1) * 2 + (i & 1) * 64, a_hi, b_lo0 + (i >> 1) * 2, b_hi, idesc, 1u);
          if (FL & 4) tc_commit(&fbar[1 + (g & 1)]);
        }
        __syncwarp();
        t_issue += clock64() - tb2;
      }
      if (lane == 0) { out[blockIdx.x * 2 + 1] = t_issue / (n_mma / 8) * 1000 + t_wait / (n_mma / 8); }
    } else
    for (int g = 0; g < n_mma / 36; ++g) {
      if (elect_one()) {
#pragma unroll 4
        for (int i = 0; i < 36; ++i) {
          const int tap = i % 9;
          const uint32_t a = a_lo0 + uint32_t(((tap / 3) * 18 + tap % 3) * ROWB) / 16 + (i & 1) * 2;
          umma_bf16_ss2(tb + (i & 1) * N, a, a_hi, b_lo0 + (i & 1) * 2, b_hi, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; done = 1; }
  } else if (warp == 2 && BG == 1) {
    if (lane == 0) {
      long long bytes = 0;
      uint32_t ph[4] = {0, 0, 0, 0};
      int s = 0;
      // 4 fills in flight, each `fill_bytes` into its own slot (above the MMA operands)
      for (int i = 0; i < 4; ++i) {
        mbar_arrive_expect_tx(&fbar[i], fill_bytes);
        bulk_load_1d(smem + 96 * 1024 + i * 16384, gsrc + (size_t(blockIdx.x) * 4 + i) * 16384, fill_bytes, &fbar[i]);
      }
      while (!done) {
        mbar_wait(&fbar[s], ph[s]);
        ph[s] ^= 1;
        bytes += fill_bytes;
        mbar_arrive_expect_tx(&fbar[s], fill_bytes);
        bulk_load_1d(smem + 96 * 1024 + s * 16384, gsrc + (size_t(blockIdx.x) * 4 + s) * 16384, fill_bytes, &fbar[s]);
        s = (s + 1) & 3;
      }
      for (int i = 0; i < 4; ++i) { mbar_wait(&fbar[s], ph[s]); ph[s] ^= 1; s = (s + 1) & 3; }
      out[blockIdx.x * 2 + 1] = bytes;
    }
  } else if (warp >= 4 && BG == 4) {
    // epilogue-like TMEM reads of accumulator columns the MMAs are not writing (cols 2N..2N+31 or 256+)
    long long n = 0;
    const int q = warp & 3;
    const uint32_t taddr = tb + (uint32_t(q * 32) << 16) + (2 * N <= 480 ? 2 * N : 0);
    uint32_t acc = 0;
    while (!done) {
      uint32_t r[32];
      tmem_ld_32x32(taddr, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r[i];
      ++n;
    }
    if (acc == 0x12345678u) gdst[0] = make_uint4(acc, 0, 0, 0);
    if (lane == 0 && warp == 4) out[blockIdx.x * 2 + 1] = n * 8 * 4096;
  } else if (warp >= 4 && (BG == 2 || BG == 3)) {
    // 8 warps, each thread "owns a pixel" of 64 bytes (BG 2) or the warp writes 2 KB contiguously (BG 3)
    long long bytes = 0;
    const int w = warp - 4;
    uint4* base = gdst + (size_t(blockIdx.x) * 8 + w) * 4096;     // 64 KB window per warp, reused (stays in L2)
    int it = 0;
    while (!done) {
      uint4* p = base + (it & 31) * 128;
      const uint4 v = make_uint4(it, lane, w, 0);
      if (BG == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) p[lane * 4 + j] = v;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j * 32 + lane] = v;
      }
      bytes += 2048;
      ++it;
    }
    if (lane == 0 && w == 0) out[blockIdx.x * 2 + 1] = bytes * 8;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, int BG, int FL = 7>
static void run(const uint8_t* gsrc, uint4* gdst, int fill_bytes = 16384) {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  cudaMemset(d, 0, 148 * 2 * sizeof(long long));
  const int smem = 1024 + 160 * 1024;
  cudaFuncSetAttribute(k<N, BG, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 36 * 128;
  for (int rep = 0; rep < 2; ++rep) k<N, BG, FL><<<148, 384, smem>>>(n_mma, gsrc, gdst, d, fill_bytes);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double clk = 0, by = 0;
  for (int i = 0; i < 148; ++i) { clk += h[2 * i] / 148.0; by += h[2 * i + 1] / 148.0; }
  static const char* names[] = {"alone", "+bulk fills", "+strided stores", "+coalesced stores", "+tcgen05.ld x8 warps", "conv loop structure", "clean loop", "two issuing warps"};
  printf("N=%3d FL=%d %-18s fill=%5d: %6.1f clk/MMA (ideal %5.1f), background %6.1f B/clk/SM   %s\n", N, FL, names[BG], fill_bytes,
         clk / n_mma, N / 2.0, BG == 5 ? double(h[1]) : (BG == 7 ? by / n_mma : by / clk), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  uint8_t* gsrc; uint4* gdst;
  cudaMalloc(&gsrc, size_t(148) * 4 * 16384);
  cudaMemset(gsrc, 0, size_t(148) * 4 * 16384);
  cudaMalloc(&gdst, size_t(148) * 8 * 4096 * 16);
  run<128, 7, 5>(gsrc, gdst); run<64, 7, 5>(gsrc, gdst); run<32, 7, 5>(gsrc, gdst); run<128, 7, 0>(gsrc, gdst); run<256, 7, 5>(gsrc, gdst);
  return 0;
  run<32, 0>(gsrc, gdst);  run<32, 1>(gsrc, gdst);  run<32, 1>(gsrc, gdst, 4096); run<32, 2>(gsrc, gdst);  run<32, 3>(gsrc, gdst);
  run<64, 0>(gsrc, gdst);  run<64, 1>(gsrc, gdst);  run<64, 2>(gsrc, gdst);  run<64, 3>(gsrc, gdst);
  run<128, 0>(gsrc, gdst); run<128, 1>(gsrc, gdst); run<128, 1>(gsrc, gdst, 4096); run<128, 2>(gsrc, gdst); run<128, 3>(gsrc, gdst);
  run<256, 0>(gsrc, gdst); run<256, 1>(gsrc, gdst); run<256, 2>(gsrc, gdst); run<256, 3>(gsrc, gdst);
  return 0;
}

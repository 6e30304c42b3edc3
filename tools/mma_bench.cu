// Issue-rate microbenchmark for tcgen05.mma kind::f16 (bf16 -> fp32), M=128, K=16, SS mode, cta_group::1.
// Measures cycles per MMA as a function of N and of the operand row pitch / swizzle (64 B vs 128 B rows), with the
// conv kernel's descriptor geometry (A: 8-row groups 18 rows apart inside an 18x18 halo tile, B: dense rows).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dt4image_restoration_b200/csrc tools/mma_bench.cu -o tools/mma_bench
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace pnp;

struct Res { long long clk; };

template <int N, int ROWB, int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int groups, int per_group, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  // zero operands (values irrelevant for timing, but keep them finite)
  for (int i = threadIdx.x; i < (48 * 1024 + 40 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    constexpr uint32_t kLayout = (ROWB == 128) ? 2u : 4u;
    // MODE 0: conv geometry (A groups 18 rows apart); MODE 1: dense A (groups 8 rows apart)
    constexpr uint32_t a_sbo = (MODE == 0 ? 18 : 8) * ROWB;
    constexpr uint32_t a_hi = (a_sbo >> 4) | (1u << 14) | (kLayout << 29);
    constexpr uint32_t b_hi = (uint32_t(8 * ROWB) >> 4) | (1u << 14) | (kLayout << 29);
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (1u << 16);
    const uint32_t b_lo0 = (smem_u32(smem + 48 * 1024) >> 4) | (1u << 16);
    long long t0 = 0, t1 = 0;
    uint32_t ph = 0;
    for (int rep = 0; rep < 2; ++rep) {       // rep 0 = warm-up
      t0 = clock64();
      for (int g = 0; g < groups; ++g) {
        if (elect_one()) {
#pragma unroll 4
          for (int i = 0; i < per_group; ++i) {
            const int tap = i % 9;
            const uint32_t a = a_lo0 + uint32_t(((tap / 3) * 18 + tap % 3) * ROWB) / 16 + (i & 1) * 2;
            const uint32_t b = b_lo0 + (i & 1) * 2;
            umma_bf16_ss2(tb + (i & 1) * N, a, a_hi, b, b_hi, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, ph);
      ph ^= 1;
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, int ROWB, int MODE>
static void run(const char* tag) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 1024 + 48 * 1024 + 40 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel<N, ROWB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int groups = 64, per = 36;
  mma_rate_kernel<N, ROWB, MODE><<<148, 128, smem>>>(groups, per, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 60;
  for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  const double n = double(groups) * per;
  printf("%-10s N=%3d rowB=%3d : %7.1f clk/MMA (min %7.1f)  ideal %5.1f  -> %4.1f%% of math rate   %s\n", tag, N, ROWB,
         mx / n, mn / n, N / 2.0, 100.0 * (N / 2.0) / (mx / n), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<32, 64, 0>("conv-geom");   run<64, 64, 0>("conv-geom");   run<96, 64, 0>("conv-geom");
  run<128, 64, 0>("conv-geom");  run<192, 64, 0>("conv-geom");  run<256, 64, 0>("conv-geom");
  run<32, 128, 0>("conv-geom");  run<64, 128, 0>("conv-geom");  run<96, 128, 0>("conv-geom");
  run<128, 128, 0>("conv-geom"); run<192, 128, 0>("conv-geom"); run<256, 128, 0>("conv-geom");
  run<32, 128, 1>("dense");      run<64, 128, 1>("dense");      run<128, 128, 1>("dense");   run<256, 128, 1>("dense");
  run<32, 64, 1>("dense");       run<128, 64, 1>("dense");
  return 0;
}

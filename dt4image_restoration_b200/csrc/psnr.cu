// PSNR reward (reference evaluation/env.py:120-125):  clamp(x,0,1); mse = mean((x-gt)^2) per image;
// psnr = 10*log10(1/mse).  A cluster of S CTAs per image (S = 1, 2, 4 or 8, chosen so that small batches still cover the
// chip: one CTA per image left a single SM busy at B = 1), each CTA sums a contiguous slice with float4 streaming loads and
// a warp-shuffle tree, the slices meet in the leader's shared memory (distributed shared memory, fixed summation order:
// the result does not depend on timing).  HBM-bound: 8 B/pixel/image read, 4 B/image written.
#include "common.cuh"
#include "pnp_internal.h"

namespace pnp {

__device__ __forceinline__ float sq_clamped(float a, float g) {
  const float d = fminf(fmaxf(a, 0.f), 1.f) - g;
  return d * d;
}

// Fused reward all-gather over peer memory: the only exchange of the multi-GPU path.  `base[p]` is the
// address, as mapped into THIS process, of rank p's copy of one symmetric buffer (NVLink / NVSwitch peer mapping):
//   floats [parity][rank][slot] : the gathered rewards, double-buffered by call parity
//   word   flag_word            : arrival counter, +1 from every rank per call
// The CTA that finishes an image stores its reward straight into every rank's buffer; the last CTA of the launch makes
// those stores visible system-wide, bumps every rank's arrival counter and then waits until its own counter shows that
// all ranks have delivered - when the kernel completes the gather is complete, no separate collective launch.
struct PeerGather {
  unsigned long long base[8];
  int rank, world;
  int slot;                     // floats per rank
  int parity;                   // call number & 1
  int flag_word;                // 4-byte word index of the arrival counter
  unsigned int count_target;    // value of *local_count once every CTA of THIS launch has finished
  unsigned int flag_target;     // world * call number
  unsigned int* local_count;    // device counter of finished CTAs (monotonic over calls)
  int* err;                     // set to 1 if the wait timed out (a rank that never arrives must not hang the GPU)
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool GATHER>
__global__ void __launch_bounds__(512) psnr_kernel_t(const float* __restrict__ x, const float* __restrict__ gt,
                                                     long long gt_bstride, float* __restrict__ out, int HW,
                                                     const PeerGather g) {
  uint32_t crank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int b = blockIdx.x / int(csize);
  const float* xb = x + size_t(b) * HW;
  const float* gb = gt + size_t(b) * gt_bstride;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const int n4 = ((reinterpret_cast<uintptr_t>(xb) | reinterpret_cast<uintptr_t>(gb)) & 15) == 0 ? HW / 4 : 0;
  const float4* x4 = reinterpret_cast<const float4*>(xb);
  const float4* g4 = reinterpret_cast<const float4*>(gb);
  // slice of this CTA: float4 indices [lo4, hi4); the scalar tail (unaligned or HW % 4) belongs to the leader
  const int per = (n4 + int(csize) - 1) / int(csize);
  const int lo4 = min(n4, int(crank) * per), hi4 = min(n4, lo4 + per);
  for (int i = lo4 + threadIdx.x; i < hi4; i += blockDim.x) {
    const float4 a = __ldg(x4 + i), g = __ldg(g4 + i);
    acc0 += sq_clamped(a.x, g.x);
    acc1 += sq_clamped(a.y, g.y);
    acc2 += sq_clamped(a.z, g.z);
    acc3 += sq_clamped(a.w, g.w);
  }
  if (crank == 0)
    for (int i = n4 * 4 + threadIdx.x; i < HW; i += blockDim.x) acc0 += sq_clamped(xb[i], gb[i]);
  float s = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float part[16];
  __shared__ float slices[8];                    // leader: the partial sums of the cluster's CTAs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) part[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && csize > 1) {                // hand the slice sum to the leader (rank 0) through DSMEM
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&slices[crank])), "r"(0));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(s) : "memory");
    }
  }
  if (csize > 1) {                               // uniform over the cluster
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    if (crank != 0) return;
  }
  if (warp == 0) {
    if (lane == 0 && csize > 1) {
      s = 0.f;
      for (uint32_t r = 0; r < csize; ++r) s += slices[r];
    }
    if (lane == 0) {
      const float mse = s / float(HW);
      const float r = 10.f * log10f(1.f / mse);
      if (out) out[b] = r;
      if constexpr (GATHER) {
        const size_t dst = (size_t(g.parity) * g.world + g.rank) * g.slot + b;
        for (int p = 0; p < g.world; ++p) reinterpret_cast<volatile float*>(g.base[p])[dst] = r;
        __threadfence_system();
        if (atomicAdd(g.local_count, 1u) + 1u == g.count_target) {          // last CTA of this launch
          __threadfence_system();
          for (int p = 0; p < g.world; ++p) atomicAdd_system(reinterpret_cast<unsigned int*>(g.base[p]) + g.flag_word, 1u);
          const unsigned int* mine = reinterpret_cast<const unsigned int*>(g.base[g.rank]) + g.flag_word;
          const unsigned long long t0 = global_timer_ns();
          while (ld_acquire_sys(mine) < g.flag_target) {
            if (global_timer_ns() - t0 > 10000000000ull) { *g.err = 1; break; }   // 10 s
            __nanosleep(200);
          }
        }
      }
    }
  }
}

// CTAs per image: enough slices to put ~2 CTAs on every SM, at most 8 (portable cluster size), at least 32 KB of input each
template <bool GATHER>
static int launch_psnr(const float* x, const float* gt, long long gt_bstride, float* out, int B, int HW, const PeerGather& g,
                       cudaStream_t st) {
  int S = 1;
  while (S < 8 && B * S * 2 <= 2 * num_sms() && HW / (S * 2) >= 8192) S *= 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(B * S);
  cfg.blockDim = dim3(512);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return int(cudaLaunchKernelEx(&cfg, psnr_kernel_t<GATHER>, x, gt, gt_bstride, out, HW, g));
}

int psnr_launch(const float* x, const float* gt, long long gt_bstride, float* out, int B, int HW, cudaStream_t st) {
  if (B <= 0 || HW <= 0) return -1;
  return launch_psnr<false>(x, gt, gt_bstride, out, B, HW, PeerGather{}, st);
}

int psnr_allgather_launch(const float* x, const float* gt, long long gt_bstride, float* out_local,
                          const unsigned long long* peer_base, int rank, int world, int slot, int parity, int flag_word,
                          unsigned int* local_count, unsigned int count_target, unsigned int flag_target, int* err,
                          int B, int HW, cudaStream_t st) {
  if (B <= 0 || HW <= 0 || world < 1 || world > 8 || rank < 0 || rank >= world || B > slot) return -1;
  PeerGather g{};
  for (int p = 0; p < world; ++p) g.base[p] = peer_base[p];
  g.rank = rank; g.world = world; g.slot = slot; g.parity = parity & 1; g.flag_word = flag_word;
  g.count_target = count_target; g.flag_target = flag_target; g.local_count = local_count; g.err = err;
  return launch_psnr<true>(x, gt, gt_bstride, out_local, B, HW, g, st);
}

}  // namespace pnp

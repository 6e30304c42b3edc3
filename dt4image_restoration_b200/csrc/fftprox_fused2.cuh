// FFT-prox + dual update for 256x256 images, second-generation single-launch kernel (4-CTA cluster).
//
// Same dataflow as fftprox_fused.cuh (rows -> cluster transpose -> columns -> blend -> columns -> transpose ->
// rows -> epilogue) but built around a register-resident radix-16 x radix-16 256-point FFT:
//   * a half-warp owns one row (lane j, register r <-> element j + 16 r); pass A is a 16-point DFT in
//     registers, ONE swizzled shared-memory round trip re-distributes the data, pass B (twiddles + 16-point DFT)
//     leaves element 16 r + j in (lane j, register r) - the same pattern pass A consumes - so global loads feed
//     pass A directly, the k-space blend and the following inverse transform run register to register, and the
//     last pass stores straight to global memory.  All global accesses are 128-byte coalesced.
//   * shared memory holds only the 64 x 256 tile per CTA (128 KB, XOR-swizzled so that row walks, column walks
//     and the 16-byte vector stores of pass A are all bank-conflict free): 8 tile passes per image instead of 16.
//   * y0 and the mask are consumed TRANSPOSED ([k_j][k_i]); they are constants of a trajectory and are prepared
//     once (prox_prepare_transposed) so that the blend reads them coalesced.
// Included by fftprox.cu.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "fft_core.cuh"

namespace pnp {
namespace cg = cooperative_groups;

struct Fused2Params {
  const float* x;
  const float2* u_in;
  const float2* y0T;        // [B][kj][ki]  (transposed), already multiplied by s*D
  const uint8_t* maskT;     // [B or 1][kj][ki]
  long long mask_bstride;
  const float* mu;
  int mu_stride;
  float2* z_out;
  float2* u_out;
  float* v_out;
  int B;
  const int* skip_flag;     // optional: != 0 means the column-only-mask kernel (fftprox_sep.cuh) handles this batch
  int prefetch_y0;          // 1: pull the row of y0T into L2 while the column transform runs
  int relaxed_barrier;      // 1: the execution-only cluster barriers inside the transposes arrive relaxed
};

constexpr int kF2N = 256;
// CL = cluster size: 4 CTAs x 64 rows x 512 threads (one CTA per SM) or 8 CTAs x 32 rows x 256 threads (two CTAs of
// different images per SM, so one image's memory phases overlap the other's arithmetic).
template <int CL> struct F2Cfg {
  static constexpr int R = kF2N / CL;
  static constexpr int THREADS = R * 8;                 // a half-warp per row, two rows per half-warp
  static constexpr size_t SMEM = size_t(R) * kF2N * sizeof(float2) + 98 * sizeof(float2);   // tile + twiddle rows + sink
  // DIRECT variant: + one 256-point scratch row per half-warp (the tile stays live while peers gather from it)
  static constexpr size_t SMEM_DIRECT = SMEM + size_t(THREADS / 16) * kF2N * sizeof(float2);
};

// Optional per-phase cycle accounting (build with -DPNP_PROX_PHASE_TIMING, tools/prox_phases.py): thread 0 of every CTA
// reads %clock64 at the phase boundaries of each image and adds the differences to g_f2_phase[]; slot 8 counts images.
// The default build contains none of this.
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

// distributed shared memory: address of `saddr` (shared::cta window) in CTA `rank` of the cluster, 8-byte accesses
__device__ __forceinline__ uint32_t f2_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ float2 f2_ld_cluster(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void f2_st_cluster(uint32_t addr, float2 v) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// tile element (row rho, index i): conflict free for row walks (lanes along i) and column walks (lanes along rho)
__device__ __forceinline__ int t_idx(int rho, int i) { return rho * kF2N + (i ^ (rho & 15)); }

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

// Cluster transpose (pull): afterwards tile[c][row] holds what was element (row % 64, rank*64 + c) of CTA row / 64.
template <int CL>
__device__ __forceinline__ void f2_transpose(float2* tile, unsigned rank, bool relaxed) {
  const uint32_t sink = smem_u32(tile + size_t(F2Cfg<CL>::R) * kF2N + 96);
  constexpr int kF2R = F2Cfg<CL>::R, kF2Threads = F2Cfg<CL>::THREADS;
  constexpr int EPT = kF2R * kF2N / kF2Threads;   // 32
  float2 v[EPT];
  cg::cluster_group cl = cg::this_cluster();
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int e = i * kF2Threads + threadIdx.x;
    const int row = e / kF2R, c = e % kF2R;          // lanes along c: contiguous 8-byte reads of the peer's row
    const float2* src = cl.map_shared_rank(tile, row / kF2R);
    v[i] = src[t_idx(row % kF2R, int(rank) * kF2R + c)];
  }
  // Execution-only cluster barrier (nothing is published here; the only hazard is a peer overwriting its tile while one of
  // these remote loads is still in flight).  A released arrive costs a full memory barrier (14 % of the kernel's stall
  // samples, profiles/r01_ncu_full_v3_prox_fused2.txt), so the arrive is RELAXED and ordered after the loads by a data
  // dependency instead: the xor chain cannot issue before every loaded register has arrived, and a warp issues in order.
  if (relaxed) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) acc ^= __float_as_uint(v[i].x) ^ __float_as_uint(v[i].y);
    // consume acc: a conditional store into a sink word nobody reads (taken once in 2^32, harmless)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %0, 0x7fc12345;\n\t@p st.shared.u32 [%1], %0;\n\t}" ::"r"(acc), "r"(sink) : "memory");
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    cl.sync();
  }
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int e = i * kF2Threads + threadIdx.x;
    const int row = e / kF2R, c = e % kF2R;
    tile[t_idx(c, row)] = v[i];
  }
  __syncthreads();
}

// DIRECT = false: two in-place cluster transposes (pull into registers, cluster barrier, local store) around the column
// pass.  DIRECT = true: no transposes - the column pass gathers its 256 elements straight from the eight peers' tiles
// (16 remote 8-byte loads per thread, bank-conflict free thanks to the XOR swizzle), transforms / blends / transforms
// back in registers and scatters the results to the addresses they came from; two cluster barriers per image instead
// of four plus three CTA barriers, half the shared-memory traffic, no 32-element staging array.
template <int CL, bool DIRECT>
__global__ void __launch_bounds__(F2Cfg<CL>::THREADS, CL == 8 ? 2 : 1) fftprox_fused2_kernel(const Fused2Params p) {
  constexpr int kF2R = F2Cfg<CL>::R, kF2Threads = F2Cfg<CL>::THREADS, kF2CL = CL;
  if (p.skip_flag && *p.skip_flag != 0) return;      // uniform over the whole grid, before any cluster operation
  extern __shared__ float2 f2sm[];
  float2* tile = f2sm;
  float2* w256 = f2sm + size_t(kF2R) * kF2N;       // twiddle rows (see fft256_halfwarp)
  float2* scratch = w256 + 98;                     // DIRECT only: [THREADS / 16][256]  (w256[96..97] = dependency sink)
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  const int cluster_id = blockIdx.x / kF2CL, n_clusters = gridDim.x / kF2CL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, j = lane & 15;
  if (threadIdx.x < 96) {
    const int t = threadIdx.x >> 4, jj = threadIdx.x & 15;
    const int m = (t < 4) ? t + 1 : (t == 4 ? 8 : 12);
    w256[threadIdx.x] = g_tw512[2 * jj * m];         // exp(-2 pi i jj m / 256), jj m <= 180
  }
  __syncthreads();
  const float inv = 1.0f / 256.0f;                    // 1/sqrt(H W)
  const int row0 = int(rank) * kF2R;

  for (int b = cluster_id; b < p.B; b += n_clusters) {
    const size_t img = size_t(b) * kF2N * kF2N;
    F2_PHASE_BEGIN();
    // ================= rows forward: global -> registers -> tile =================
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int rho = warp * 4 + it * 2 + half;         // local row owned by this half-warp
      const size_t g0 = img + size_t(row0 + rho) * kF2N + j;
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float2 uu = __ldg(p.u_in + g0 + 16 * r);
        const float xx = __ldg(p.x + g0 + 16 * r);
        v[r] = make_float2(xx + uu.x, uu.y);
      }
      if ((row0 + rho + j) & 1) {                       // D = (-1)^(row + col); col = j + 16 r has the parity of j
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = make_float2(-v[r].x, -v[r].y);
      }
      float2* row = tile + rho * kF2N;
      fft256_halfwarp(v, row, w256, j);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) row[16 * r + (j ^ (rho & 15))] = v[r];
    }
    F2_PHASE(0);                                        // rows forward (loads of x, u + FFT + tile store)
    cl.sync();                                          // every CTA's rows are complete
    F2_PHASE(1);                                        // wait for the slowest CTA of the cluster
    if constexpr (!DIRECT) f2_transpose<CL>(tile, rank, p.relaxed_barrier != 0);
    F2_PHASE(2);                                        // transpose (remote loads, cluster barrier, local stores)
    if (b + n_clusters < p.B) {                         // warm L2 with the next image's rows of x and u
      const size_t nimg = size_t(b + n_clusters) * kF2N * kF2N + size_t(row0) * kF2N;
      const char* pu = reinterpret_cast<const char*>(p.u_in + nimg);
      const char* px = reinterpret_cast<const char*>(p.x + nimg);
      for (int o = threadIdx.x * 128; o < kF2R * kF2N * 8; o += kF2Threads * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pu + o));
      for (int o = threadIdx.x * 128; o < kF2R * kF2N * 4; o += kF2Threads * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o));
    }
    // ================= columns: forward, blend, inverse (register to register) =================
    {
      const float mu = __ldg(p.mu + size_t(b) * p.mu_stride);
      const float inv1mu = 1.f / (1.f + mu);
#pragma unroll 1
      for (int it = 0; it < 2; ++it) {
        const int c = warp * 4 + it * 2 + half;         // local column
        const int kj = row0 + c;
        float2* row = DIRECT ? scratch + (warp * 2 + half) * kF2N : tile + c * kF2N;
        float2 v[16];
        // DIRECT: element (image row j + 16 r, column kj) lives in CTA r / 2, local row j + 16 (r & 1)
        uint32_t ra[DIRECT ? 16 : 1];
        if constexpr (DIRECT) {
          const uint32_t a0 = smem_u32(tile) + uint32_t(t_idx(j, kj)) * 8u;      // (rho & 15) == j for both local rows
#pragma unroll
          for (int r = 0; r < 16; ++r) ra[r] = f2_mapa(a0 + uint32_t(r & 1) * (16u * kF2N * 8u), uint32_t(r >> 1));
#pragma unroll
          for (int r = 0; r < 16; ++r) v[r] = f2_ld_cluster(ra[r]);
        } else {
#pragma unroll
          for (int r = 0; r < 16; ++r) v[r] = row[16 * r + (j ^ (c & 15))];
          __syncwarp();
        }
        // mask bits for (kj, ki = 16 r + j) are fetched before the FFT (one register); y0T after it
        const float2* yp = p.y0T + img + size_t(kj) * kF2N + j;
        if (p.prefetch_y0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.y0T + img + size_t(kj) * kF2N + 16 * j));
        // row kj of maskT starts with 16 packed lane masks (mask_pack_rows_kernel): bit r of entry j = maskT[kj][16 r + j]
        const uint32_t mbits =
            __ldg(reinterpret_cast<const uint16_t*>(p.maskT + size_t(b) * p.mask_bstride + size_t(kj) * kF2N) + j);
        fft256_halfwarp(v, row, w256, j);
#pragma unroll
        for (int r = 0; r < 16; ++r) {                  // element ki = 16 r + j
          float2 Z = make_float2(v[r].x * inv, v[r].y * inv);
          if ((mbits >> r) & 1u) {
            const float2 y = __ldg(yp + 16 * r);
            Z.x = (mu * Z.x + y.x) * inv1mu;
            Z.y = (mu * Z.y + y.y) * inv1mu;
          }
          if ((r & 7) == 7) asm volatile("" ::: "memory");
          v[r] = make_float2(Z.x, -Z.y);                // conj: forward FFT == inverse
        }
        fft256_halfwarp(v, row, w256, j);
        __syncwarp();
        if constexpr (DIRECT) {
#pragma unroll
          for (int r = 0; r < 16; ++r) f2_st_cluster(ra[r], v[r]);   // (lane j, register r) holds element 16 r + j before and after
        } else {
#pragma unroll
          for (int r = 0; r < 16; ++r) row[16 * r + (j ^ (c & 15))] = v[r];
        }
      }
    }
    F2_PHASE(3);                                        // L2 prefetch + columns: FFT, blend (mask, y0T), inverse FFT
    cl.sync();
    F2_PHASE(4);
    if constexpr (!DIRECT) f2_transpose<CL>(tile, rank, p.relaxed_barrier != 0);
    F2_PHASE(5);
    // ================= rows inverse: tile -> registers -> global =================
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int rho = warp * 4 + it * 2 + half;
      float2* row = tile + rho * kF2N;
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = row[16 * r + (j ^ (rho & 15))];
      __syncwarp();
      fft256_halfwarp(v, row, w256, j);
      const size_t g0 = img + size_t(row0 + rho) * kF2N + j;
      const float sg = ((row0 + rho + j) & 1) ? -inv : inv;     // D / sqrt(HW); col = 16 r + j has the parity of j
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const size_t g = g0 + 16 * r;
        const float2 zz = make_float2(sg * v[r].x, -sg * v[r].y);
        const float2 uu = __ldg(p.u_in + g);
        const float xx = __ldg(p.x + g);
        const float2 un = make_float2(uu.x + xx - zz.x, uu.y - zz.y);
        p.z_out[g] = zz;
        p.u_out[g] = un;
        if (p.v_out) p.v_out[g] = zz.x - un.x;
        if ((r & 3) == 3) asm volatile("" ::: "memory");   // bound the loads hoisted ahead (register pressure)
      }
    }
    F2_PHASE(6);                                        // rows inverse + epilogue (re-read x, u; store z, u, v)
#ifdef PNP_PROX_PHASE_TIMING
    if (threadIdx.x == 0) atomicAdd(&g_f2_phase[8], 1ull);
#endif
    // No CTA barrier here: a half-warp owns the same tile rows in the last phase of this image and in the first phase of
    // the next one, and every peer has finished pulling from this tile before the last cluster barrier above.
  }
  cl.sync();
}

// y0T[b][kj][ki] = s * (-1)^(ki+kj) * y0[b][ki][kj];  maskT[b][kj][ki] = mask[b][ki][kj]   (32x32 smem tiles)
__global__ void __launch_bounds__(256) prox_prepare_kernel(const float2* __restrict__ y0, const uint8_t* __restrict__ mask,
                                                           float2* __restrict__ y0T, uint8_t* __restrict__ maskT, int N,
                                                           int nb_mask, float sgn, const int* skip_flag) {
  if (skip_flag && *skip_flag != 0) return;          // column-only masks: the row-only kernel does not need the transposes
  __shared__ float2 ty[32][33];
  __shared__ uint8_t tm[32][33];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
  const size_t img = size_t(b) * N * N;
  for (int r = ty0; r < 32; r += 8) {
    const size_t g = size_t(i0 + r) * N + j0 + tx;
    float2 y = y0[img + g];
    const float s = ((i0 + r + j0 + tx) & 1) ? -sgn : sgn;
    ty[r][tx] = make_float2(s * y.x, s * y.y);
    if (b < nb_mask) tm[r][tx] = mask[size_t(b) * N * N + g];
  }
  __syncthreads();
  for (int r = ty0; r < 32; r += 8) {
    const size_t g = size_t(j0 + r) * N + i0 + tx;
    y0T[img + g] = ty[tx][r];
    if (b < nb_mask) maskT[size_t(b) * N * N + g] = tm[tx][r];
  }
}

// In-place packing of the transposed mask for the column pass: row kj (256 bytes, one per k-space column) gets its first
// 32 bytes replaced by 16 uint16 lane masks, bit r of entry j = maskT[kj][16 r + j] - one 2-byte load per thread and
// column instead of sixteen byte loads (7 % of the kernel's stall samples were waits for them).  A half-warp per row.
__global__ void __launch_bounds__(256) mask_pack_rows_kernel(uint8_t* __restrict__ maskT, int rows, const int* skip_flag) {
  if (skip_flag && *skip_flag != 0) return;
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 4, j = threadIdx.x & 15;
  uint32_t bits = 0;
  if (row < rows) {
    const uint8_t* mp = maskT + size_t(row) * kF2N + j;
#pragma unroll
    for (int r = 0; r < 16; ++r) bits |= (mp[16 * r] ? 1u : 0u) << r;
  }
  __syncwarp();                                      // every byte of the row has been read before its head is overwritten
  if (row < rows) reinterpret_cast<uint16_t*>(maskT + size_t(row) * kF2N)[j] = uint16_t(bits);
}

template <int CL, bool DIRECT>
static int launch_fused2_t(const Fused2Params& p, int num_sms, cudaStream_t st) {
  constexpr size_t kSmem = DIRECT ? F2Cfg<CL>::SMEM_DIRECT : F2Cfg<CL>::SMEM;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(fftprox_fused2_kernel<CL, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(kSmem));
    if (e != cudaSuccess) return int(e);
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(num_sms / CL * CL);
  cfg.blockDim = dim3(F2Cfg<CL>::THREADS);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fftprox_fused2_kernel<CL, DIRECT>, &cfg) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      n = num_sms / CL;
    }
    max_clusters = n;
  }
  int clusters = max_clusters < p.B ? max_clusters : p.B;
  if (clusters < 1) clusters = 1;
  cfg.gridDim = dim3(clusters * CL);
  return int(cudaLaunchKernelEx(&cfg, fftprox_fused2_kernel<CL, DIRECT>, p));
}

static int launch_fused2(const Fused2Params& p, int num_sms, cudaStream_t st) {
  static const int cl = [] { const char* e = getenv("PNP_PROX_CLUSTER"); return e ? atoi(e) : 8; }();
  static const int direct = [] { const char* e = getenv("PNP_PROX_DIRECT"); return e ? atoi(e) : 0; }();
  if (cl == 4) return launch_fused2_t<4, false>(p, num_sms, st);
  return direct ? launch_fused2_t<8, true>(p, num_sms, st) : launch_fused2_t<8, false>(p, num_sms, st);
}

}  // namespace pnp

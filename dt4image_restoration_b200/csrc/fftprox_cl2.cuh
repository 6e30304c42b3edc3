// FFT-prox + dual update for arbitrary sampling masks, 256x256, batch variant of the cluster kernel (fftprox_cl.cuh):
// 8-CTA clusters x 32 rows, 256 threads, ONE exchange buffer per CTA, two CTAs per SM.
//
// Why a second shape.  fftprox_cl_kernel<16> keeps three 32 KB buffers + x per CTA (112 KB, two CTAs per SM) and needs
// 16-CTA clusters; the chip then holds 14 clusters (tools/cluster_occ.cu: GPC packing) = 14 images on 112 of the 148 SMs,
// and a batch of 64 takes five rounds.  Here a CTA keeps a single 64 KB buffer M that is the column-domain buffer after
// exchange 1 and the row-domain buffer after exchange 2, plus 32 KB of transform scratch: 97 KB, two CTAs per SM with
// 8-CTA clusters, of which 33 fit (132 SMs) - a batch of 64 takes two rounds.  The price: x and u are read with plain
// coalesced loads (no room to stage them; the lines are pulled into L2 one image ahead), and M is time-shared under two
// cluster barriers, both split into arrive / wait with a transform in between:
//   X  "every CTA has read its last column element of M"   -> exchange 2 may overwrite M
//   Y  "every CTA has read its last row element of M"      -> exchange 1 of the next image may overwrite M
// A thread owns two rows in the row phases (half-warp h: rows h and h + 16) and two columns in the column phase.
// Algebra, prepared constants (y0R, packed mask) and exchanges (st.async + mbarrier tx counts) as in fftprox_cl.cuh.
#pragma once
#include "fftprox_cl.cuh"

namespace pnp {

// relaxed cluster-barrier arrive that the compiler cannot hoist above the computation of `a` and `b`
__device__ __forceinline__ void cl_cluster_arrive_after(float a, float b) {
  asm volatile("{\n\t.reg .f32 t;\n\tadd.f32 t, %0, %1;\n\tbarrier.cluster.arrive.relaxed;\n\t}" ::"f"(a), "f"(b) : "memory");
}

constexpr int kCl2CL = 8, kCl2R = 32, kCl2Threads = 256;
constexpr int kCl2Buf = kCl2R * kClN;                  // float2 elements of M
constexpr int kCl2Blk = kCl2R * kCl2R;                 // float2 elements per exchange block
constexpr size_t kCl2Smem = size_t(kCl2Buf) * 8 + size_t(16) * kClN * 8 + 96 * 8 + 64;

// 256-point DFT of a row held by a half-warp, contiguous 256-float2 scratch row: in v[r] = x[j + 16 r], out v[r] = X[16 r + j]
template <bool INV>
__device__ __forceinline__ void fft256_row_contig(float2 (&v)[16], float2* row, const float2* wtab, int j) {
  dft16t<INV>(v);
  {
    float4* dst = reinterpret_cast<float4*>(row + 16 * j);
#pragma unroll
    for (int m = 0; m < 8; ++m) dst[m ^ (j & 7)] = make_float4(v[2 * m].x, v[2 * m].y, v[2 * m + 1].x, v[2 * m + 1].y);
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = row[16 * r + (j ^ ((r & 7) << 1))];
  __syncwarp();
  twiddle16<INV>(v, wtab, j);
  dft16t<INV>(v);
}

__global__ void __launch_bounds__(kCl2Threads, 2) fftprox_cl2_kernel(const ClParams p) {
  constexpr int R = kCl2R, BLK = kCl2Blk;
  extern __shared__ __align__(128) uint8_t cl2_smem[];
  float2* M = reinterpret_cast<float2*>(cl2_smem);       // [8 senders][32][32]: columns after exchange 1, rows after exchange 2
  float2* S = M + kCl2Buf;                               // 16 scratch rows (one per half-warp) for the row transforms
  float2* wf = S + 16 * kClN;                            // twiddle rows (forward; the inverse passes conjugate them)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wf + 96);
  uint64_t* bfull = bars;                                // exchange 1 received
  uint64_t* afull = bars + 1;                            // exchange 2 received
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cl_cluster_rank();
  const int cluster_id = blockIdx.x / kCl2CL, n_clusters = gridDim.x / kCl2CL;
  const int row0 = int(rank) * R;
  if (tid < 96) {
    const int t = tid >> 4, jj = tid & 15;
    const int m = (t < 4) ? t + 1 : (t == 4 ? 8 : 12);
    wf[tid] = g_tw512[2 * jj * m];
  }
  if (tid == 0) {
    mbar_init(bfull, 1);
    mbar_init(afull, 1);
    fence_mbar_init();
  }
  grid_dep_wait();                                       // programmatic stream serialization: see fftprox_cl.cuh
  grid_dep_launch();
  if (p.skip_flag && *p.skip_flag != 0) return;
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);                          // w = x + u of the image in flight: 64 columns per thread
    tmem_relinquish();
  }
  if (tid == 0) mbar_arrive_expect_tx(bfull, uint32_t(kCl2Buf) * 8);   // before this CTA's first arrive: no peer sends earlier
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_w = *tmem_slot + (uint32_t(32 * (warp & 3)) << 16) + uint32_t(64 * (warp >> 2));
  cl_cluster_arrive();                                   // Y of "image -1": barriers initialised, M free

  const int hw = tid >> 4, j = tid & 15;                 // row phases: half-warp hw owns local rows hw and hw + 16, lane j
  const int c0 = tid & 15, jc = tid >> 4;                // column phase: columns c0 and c0 + 16, residue jc
  const float inv2 = 1.0f / 65536.0f;
  const uint32_t bfull_a = smem_u32(bfull), afull_a = smem_u32(afull);
  float2* srow = S + hw * kClN;

  int it = 0;
  for (int b = cluster_id; b < p.B; b += n_clusters, ++it) {
    const uint32_t par = it & 1;
    const size_t img = size_t(b) * kClN * kClN;
    const float mu = __ldg(p.mu + size_t(b) * p.mu_stride);
    if (tid == 0) mbar_arrive_expect_tx(afull, uint32_t(kCl2Buf) * 8);   // before this thread's rows leave

    // ================= rows forward: global -> registers -> peers' M =================
    F2_PHASE_BEGIN();
    bool waited = false;
#pragma unroll 1
    for (int h2 = 0; h2 < 2; ++h2) {
      const int rho = hw + 16 * h2;
      const size_t g0 = img + size_t(row0 + rho) * kClN + j;
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float2 uu = p.u_in[g0 + 16 * r];           // plain load: u_out may alias u_in (rows are read before written)
        v[r] = make_float2(__ldg(p.x + g0 + 16 * r) + uu.x, uu.y);
      }
      {
        uint32_t wr[32];
#pragma unroll
        for (int r = 0; r < 16; ++r) { wr[2 * r] = __float_as_uint(v[r].x); wr[2 * r + 1] = __float_as_uint(v[r].y); }
        tmem_st_32x32(tmem_w + 32 * h2, wr);
      }
      fft256_row_contig<false>(v, srow, wf, j);            // v[r] = H[row][16 r + j]
      if (!waited) { cl_cluster_wait(); waited = true; }  // Y: every peer has read its rows of the previous image out of M
      const uint32_t dst0 = smem_u32(M + rank * BLK + rho * R + j);
#pragma unroll
      for (int r = 0; r < 16; ++r)
        cl_st_async(cl_mapa(dst0 + uint32_t((16 * r) % R) * 8u, (16 * r) / R), v[r], cl_mapa(bfull_a, (16 * r) / R));
    }
    F2_PHASE(2);                                         // rows forward + sends
    // next image: pull this CTA's rows of x and u into L2 (768 lines of 128 bytes, three per thread)
    if (b + n_clusters < p.B) {
      const size_t g = img + size_t(n_clusters) * kClN * kClN + size_t(row0) * kClN;
      const char* pu = reinterpret_cast<const char*>(p.u_in + g);
      const char* px = reinterpret_cast<const char*>(p.x + g);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pu + size_t(tid) * 128));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pu + size_t(tid + 256) * 128));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(px + size_t(tid) * 128));
    }
    uint32_t mbits[2];
    mbits[0] = __ldg(p.mpack + size_t(b) * p.mpack_bstride + jc * kClN + row0 + c0);
    mbits[1] = __ldg(p.mpack + size_t(b) * p.mpack_bstride + jc * kClN + row0 + c0 + 16);
    mbar_wait(bfull, par);
    if (tid == 0) mbar_arrive_expect_tx(bfull, uint32_t(kCl2Buf) * 8);   // next image's exchange 1
    F2_PHASE(3);                                         // wait for the peers' rows

    // ================= columns: forward, blend, inverse - in place in M =================
    const float bb = 1.f / (1.f + mu), aa = mu * bb;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {                        // stage A: first radix-16 pass of the column transform
      float2* Bc = M + c0 + 16 * s;
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = Bc[cl_col_idx<R>(jc + 16 * r, 0)];
      dft16t<false>(v);
      twiddle16<false>(v, wf, jc);
#pragma unroll
      for (int q = 0; q < 16; ++q) Bc[cl_col_idx<R>(jc + 16 * q, 0)] = v[q];
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {                        // stage B: second pass, blend, first pass of the inverse
      float2* Bc = M + c0 + 16 * s;
      const uint32_t mb = mbits[s];
      float2 y[16];
      {
        const float2* yp = p.y0R + img + size_t(jc) * kClN + row0 + c0 + 16 * s;
#pragma unroll
        for (int q = 0; q < 16; ++q) y[q] = ((mb >> q) & 1u) ? __ldg(yp + 16 * q * kClN) : make_float2(0.f, 0.f);
      }
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = Bc[cl_col_idx<R>(r + 16 * jc, 0)];
      dft16t<false>(v);                                  // v[q] = H[kappa_i = jc + 16 q][kappa_j]
#pragma unroll
      for (int q = 0; q < 16; ++q)
        if ((mb >> q) & 1u) v[q] = make_float2(aa * v[q].x + bb * y[q].x, aa * v[q].y + bb * y[q].y);
      dft16t<true>(v);
      twiddle16<true>(v, wf, jc);
#pragma unroll
      for (int q = 0; q < 16; ++q) Bc[cl_col_idx<R>(q + 16 * jc, 0)] = v[q];
    }
    __syncthreads();
    {                                                    // stage C: last pass; the results wait in registers for barrier X
      float2 va[16], vb[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) va[r] = M[c0 + cl_col_idx<R>(jc + 16 * r, 0)];
      dft16t<true>(va);
#pragma unroll
      for (int r = 0; r < 16; ++r) vb[r] = M[c0 + 16 + cl_col_idx<R>(jc + 16 * r, 0)];
      // element 0 of a DFT depends on all sixteen loaded values, and a warp issues in order: making the relaxed arrive
      // depend on it orders this thread's last reads of M before the barrier without a memory fence
      dft16t<true>(vb);
      cl_cluster_arrive_after(va[0].x, vb[0].x);         // X
      F2_PHASE(4);                                       // columns
      cl_cluster_wait();                                 // X: every peer is done reading M
      F2_PHASE(5);                                       // cluster wait X
      // exchange 2: element (row i = jc + 16 q, col) -> CTA i / R, slot [rank][i % R][c]
      const uint32_t dst0 = smem_u32(M + rank * BLK + jc * R + c0);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const uint32_t a = cl_mapa(dst0 + uint32_t(((16 * q) % R) * R) * 8u, (16 * q) / R), bar = cl_mapa(afull_a, (16 * q) / R);
        cl_st_async(a, va[q], bar);
        cl_st_async(a + 16u * 8u, vb[q], bar);
      }
    }
    mbar_wait(afull, par);
    F2_PHASE(6);                                         // exchange 2 (send + wait)

    // ================= rows inverse: M -> registers -> epilogue -> global =================
#pragma unroll 1
    for (int h2 = 0; h2 < 2; ++h2) {
      const int rho = hw + 16 * h2;
      const float2* rowp = M + rho * R;
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = rowp[((j + 16 * r) / R) * (R * R) + ((j + 16 * r) % R)];
      fft256_row_contig<true>(v, srow, wf, j);
      if (h2 == 1) cl_cluster_arrive_after(v[0].x, v[1].x);   // Y (the transform consumed this thread's last read of M)
      uint32_t wr[32];
      tmem_st_wait();
      tmem_ld_32x32(tmem_w + 32 * h2, wr);
      tmem_ld_wait();
      const size_t g0 = img + size_t(row0 + rho) * kClN + j;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float2 zz = make_float2(v[r].x * inv2, v[r].y * inv2);
        const float2 un = make_float2(__uint_as_float(wr[2 * r]) - zz.x, __uint_as_float(wr[2 * r + 1]) - zz.y);
        p.z_out[g0 + 16 * r] = zz;
        p.u_out[g0 + 16 * r] = un;
        if (p.v_out) p.v_out[g0 + 16 * r] = zz.x - un.x;
      }
    }
    F2_PHASE(7);                                         // rows inverse + epilogue
#ifdef PNP_PROX_PHASE_TIMING
    if (tid == 0) atomicAdd(&g_f2_phase[8], 1ull);
#endif
  }
  tc_fence_before();
  cl_cluster_wait();                                     // the pending Y: no CTA leaves while a peer may still write to it
  __syncthreads();
  if (warp == 1) tmem_dealloc(*tmem_slot, 128);
}

__global__ void __launch_bounds__(kCl2Threads) cl2_occupancy_probe(int* p) {
  extern __shared__ int probe2_sm[];
  if (p) p[0] = probe2_sm[0];
}

static int launch_cl2(const ClParams& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(fftprox_cl2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kCl2Smem));
    if (e != cudaSuccess) return int(e);
    e = cudaFuncSetAttribute(cl2_occupancy_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kCl2Smem));
    if (e != cudaSuccess) return int(e);
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCl2CL * 64);
  cfg.blockDim = dim3(kCl2Threads);
  cfg.dynamicSmemBytes = kCl2Smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl2CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, cl2_occupancy_probe, &cfg) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      n = 1;
    }
    max_clusters = n;
    if (const char* e = getenv("PNP_PROX_MAXCL2")) max_clusters = atoi(e);
  }
  int clusters = max_clusters < p.B ? max_clusters : p.B;
  if (clusters < 1) clusters = 1;
  const int rounds = (p.B + clusters - 1) / clusters;
  clusters = (p.B + rounds - 1) / rounds;
  cfg.gridDim = dim3(clusters * kCl2CL);
  cfg.numAttrs = 2;
  return int(cudaLaunchKernelEx(&cfg, fftprox_cl2_kernel, p));
}

}  // namespace pnp

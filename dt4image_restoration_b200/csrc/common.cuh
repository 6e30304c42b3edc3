// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05 / TMEM wrappers (inline PTX).
// Everything here is written for Blackwell B200 only (compile with -gencode arch=compute_100a,code=sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

namespace pnp {

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still running.  grid_dep_launch() lets OUR successor start early
// (it then overlaps its prologue - barrier init, TMEM allocation, weight loads - with our tail);
// grid_dep_wait() blocks until the predecessor has completed and its global writes are visible, and must precede the
// first access to anything the predecessor wrote and any global write of our own.  Both are no-ops without PDL.
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-suspending test of a phase.  The result is only needed later, so the ~100-cycle latency of the barrier unit
// overlaps whatever is issued in between (used to look one pipeline step ahead in the MMA issuer).
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug becomes a trap (reported as a CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("pnp_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load, coordinates fastest-first (c, x, y, n); out-of-range elements are zero-filled.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both sides 16 B aligned).
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: the mbarrier gets one arrival when all MMAs issued so far by this thread have completed.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Same, descriptors given as (lo, hi) 32-bit halves so loop-invariant high words stay immediates.
__device__ __forceinline__ void umma_bf16_ss2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One filter tap of the implicit GEMM as a single instruction block: 2*NK MMAs (NK k-steps of 16 channels x the two
// M=128 pixel blocks, accumulators d0 and d0+BN), bracketed by up to three non-suspending mbarrier phase tests issued
// BEFORE the MMAs (their ~100-150 cycle latency overlaps the MMA issue, which blocks while the tensor pipe drains) and
// up to two tcgen05.commit after them (barrier address 0 = none).  r0..r2 return the three test results.
// Why one block: measured on B200 (tools/mma_bench2.cu) a try_wait between MMA groups leaves the tensor pipe idle for
// its whole latency, because the pipe queues only ~1 MMA ahead of the issuing thread.
template <int NK>
__device__ __forceinline__ void umma_tap_block(uint32_t d0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate_first, uint32_t tbar0, uint32_t tpar0,
                                               uint32_t tbar1, uint32_t tpar1, uint32_t tbar2, uint32_t tpar2,
                                               uint32_t cbar0, uint32_t cbar1, uint32_t mb_step, uint32_t bn,
                                               uint32_t& r0, uint32_t& r1, uint32_t& r2) {
  static_assert(NK == 2 || NK == 4, "KC must be 32 or 64");
#define PNP_TAP_OPERANDS                                                                                              \
  : "=r"(r0), "=r"(r1), "=r"(r2)                                                                                      \
  : "r"(d0), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate_first), "r"(tbar0), "r"(tpar0),   \
    "r"(tbar1), "r"(tpar1), "r"(tbar2), "r"(tpar2), "r"(cbar0), "r"(cbar1), "r"(mb_step), "r"(bn)                     \
  : "memory"
  if constexpr (NK == 2) {
    asm volatile(
        "{\n\t.reg .pred p, pt, q0, q1, q2, c0, c1;\n\t.reg .b64 da, db;\n\t.reg .b32 ra, rb, rd1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q0, [%10], %11;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q1, [%12], %13;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q2, [%14], %15;\n\t"
        "setp.ne.b32 p, %9, 0;\n\t"
        "setp.eq.b32 pt, %9, %9;\n\t"
        "setp.ne.b32 c0, %16, 0;\n\t"
        "setp.ne.b32 c1, %17, 0;\n\t"
        "add.u32 rd1, %3, %19;\n\t"
        "add.u32 rb, %6, 0;\n\t"
        "mov.b64 db, {rb, %7};\n\t"
        "add.u32 ra, %4, 0;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%3], da, db, %8, p;\n\t"
        "add.u32 ra, ra, %18;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [rd1], da, db, %8, p;\n\t"
        "add.u32 rb, %6, 2;\n\t"
        "mov.b64 db, {rb, %7};\n\t"
        "add.u32 ra, %4, 2;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%3], da, db, %8, pt;\n\t"
        "add.u32 ra, ra, %18;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [rd1], da, db, %8, pt;\n\t"
        "@c0 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%16];\n\t"
        "@c1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%17];\n\t"
        "selp.b32 %0, 1, 0, q0;\n\t"
        "selp.b32 %1, 1, 0, q1;\n\t"
        "selp.b32 %2, 1, 0, q2;\n\t}\n"
        PNP_TAP_OPERANDS);
  } else {
    asm volatile(
        "{\n\t.reg .pred p, pt, q0, q1, q2, c0, c1;\n\t.reg .b64 da, db;\n\t.reg .b32 ra, rb, rd1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q0, [%10], %11;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q1, [%12], %13;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q2, [%14], %15;\n\t"
        "setp.ne.b32 p, %9, 0;\n\t"
        "setp.eq.b32 pt, %9, %9;\n\t"
        "setp.ne.b32 c0, %16, 0;\n\t"
        "setp.ne.b32 c1, %17, 0;\n\t"
        "add.u32 rd1, %3, %19;\n\t"
        "add.u32 rb, %6, 0;\n\t"
        "mov.b64 db, {rb, %7};\n\t"
        "add.u32 ra, %4, 0;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%3], da, db, %8, p;\n\t"
        "add.u32 ra, ra, %18;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [rd1], da, db, %8, p;\n\t"
        "add.u32 rb, %6, 2;\n\t"
        "mov.b64 db, {rb, %7};\n\t"
        "add.u32 ra, %4, 2;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%3], da, db, %8, pt;\n\t"
        "add.u32 ra, ra, %18;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [rd1], da, db, %8, pt;\n\t"
        "add.u32 rb, %6, 4;\n\t"
        "mov.b64 db, {rb, %7};\n\t"
        "add.u32 ra, %4, 4;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%3], da, db, %8, pt;\n\t"
        "add.u32 ra, ra, %18;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [rd1], da, db, %8, pt;\n\t"
        "add.u32 rb, %6, 6;\n\t"
        "mov.b64 db, {rb, %7};\n\t"
        "add.u32 ra, %4, 6;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%3], da, db, %8, pt;\n\t"
        "add.u32 ra, ra, %18;\n\t"
        "mov.b64 da, {ra, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [rd1], da, db, %8, pt;\n\t"
        "@c0 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%16];\n\t"
        "@c1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%17];\n\t"
        "selp.b32 %0, 1, 0, q0;\n\t"
        "selp.b32 %1, 1, 0, q1;\n\t"
        "selp.b32 %2, 1, 0, q2;\n\t}\n"
        PNP_TAP_OPERANDS);
  }
#undef PNP_TAP_OPERANDS
}

// Instruction descriptor for kind::f16 with bf16 inputs, fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4)                       // D format  = F32
         | (1u << 7)                     // A format  = BF16
         | (1u << 10)                    // B format  = BF16
         | (uint32_t(N >> 3) << 17)      // N / 8
         | (uint32_t(M >> 4) << 24);     // M / 16
}

// Shared-memory matrix descriptor, K-major operand whose rows are `row_bytes` (64 or 128) wide and
// stored with the matching TMA swizzle (SWIZZLE_64B / SWIZZLE_128B).  Groups of 8 rows are `sbo_bytes`
// apart.  The swizzle is a function of the absolute shared-memory address (CUTLASS' Swizzle<B,4,3> o
// smem_ptr), so `addr` may point at any row of a swizzled tile.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t row_bytes,
                                                   uint32_t base_offset) {
  const uint64_t layout = (row_bytes == 128) ? 2ull : 4ull;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return uint64_t((addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16)  // LBO: unused for swizzled K-major (1)
         | (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46)  // descriptor version (Blackwell)
         | (uint64_t(base_offset & 7u) << 49) | (layout << 61);
}

// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 registers per thread -> 32 lanes x 32 consecutive fp32 columns of TMEM (the inverse of tmem_ld_32x32).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// vectorised global access helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace pnp

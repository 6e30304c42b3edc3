// 3x3 stride-1 zero-pad-1 convolution + bias + LeakyReLU as an implicit GEMM on the 5th-gen tensor cores.
//
// Replaces the reference's nn.Conv2d + LeakyReLU(0.2) pairs (reference evaluation/noise.py:75-98) for
// every conv whose Cin is a multiple of 32 (26 of the 28 convs; the 2->32 first conv and the 1x1 output
// conv are handled elsewhere / in this kernel's FINAL epilogue).
//
// Mapping (B200-first, not a GEMM library call):
//   * activations are NHWC bf16; one CTA owns a 16x16 output-pixel tile of one image and BN output channels;
//   * for every KC-channel slice of the input, ONE TMA box load brings the 18x18 halo tile into shared
//     memory (rows = pixels, KC*2 bytes each, hardware 64B/128B swizzle, out-of-image pixels zero-filled =
//     the conv's zero padding).  All 9 filter taps are then issued as tcgen05.mma instructions whose
//     A-operand descriptors simply point at a shifted window of that same halo tile (start address moved by
//     (kh*18+kw) rows, 8-row groups 18 rows apart), so each input byte is read from L2 once per 9 taps;
//   * the 16x16 tile is two M=128 MMAs (16 rows x 8 columns each), accumulators live in TMEM
//     (2 x BN fp32 columns, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1);
//   * weights are pre-packed on the device at plan creation into ready-to-use swizzled [BN x KC] K-major
//     blobs, one per (channel slice, tap, n-tile), fetched with 1-D bulk copies into a 4-8 deep ring;
//   * warp roles: warp0 = TMA producer, warp1 = MMA issuer (one elected lane), warp2 = TMEM allocator,
//     warps 4..11 = epilogue (tcgen05.ld -> bias -> LeakyReLU -> bf16 -> global), synchronised only through
//     mbarriers;
//   * the concat of the `up` blocks (reference noise.py:59) is never materialised: the K loop walks two
//     tensor maps (skip tensor, then upsampled tensor);
//   * FINAL epilogue (last conv of up4): the 1x1 output conv (noise.py:64-71), the global residual
//     `noisy[:, :1] + residual` (noise.py:132-133) and the clamp (noise.py:164) are applied in fp32 on the
//     un-rounded accumulators and written as fp32.
#pragma once
#include "common.cuh"

namespace pnp {

constexpr int kTile = 16;                 // output tile edge (pixels)
constexpr int kHalo = kTile + 2;          // halo tile edge
constexpr int kConvThreads = 384;         // 12 warps
constexpr int kEpiWarp0 = 4;              // first epilogue warp
constexpr int kNumEpiWarps = 8;

struct ConvParams {
  int B, H, W;              // images, spatial size (input == output size)
  int tiles_x, tiles_y;     // ceil(W/16), ceil(H/16)
  int n_tiles;              // Cout / BN
  int nchunks0, nchunks1;   // KC-channel slices taken from tensor map 0 / 1
  int Cout;
  int sa, sb;               // ring depths: halo tiles / weight tiles (sb unused when wres)
  int wres;                 // 1: all weights of this CTA's n-tile stay resident in smem (loaded once)
  int img0;                 // first image of this launch (micro-batching over the batch dimension)
  const uint8_t* wpk;       // packed weights, blob index ((chunk*9 + tap)*n_tiles + nt), BN*KC*2 bytes each
  const float* bias;        // [Cout]
  __nv_bfloat16* out;       // NHWC [B,H,W,Cout]                      (EPI_BF16)
  const float* wout;        // [32] 1x1 output conv weights          (EPI_FINAL)
  const float* bout;        // 1x1 output conv bias (device scalar)
  const float* noisy;       // [B,H,W] fp32 denoiser input (channel 0)
  float* x_out;             // [B,H,W] fp32 clamp(noisy + residual, 0, 1)
  float* preclamp;          // optional [B,H,W] fp32 noisy + residual
  float slope;              // LeakyReLU negative slope (0.2)
};

enum { EPI_BF16 = 0, EPI_FINAL = 1 };

template <int KC, int BN>
struct ConvCfg {
  static constexpr int ROWB = KC * 2;                                    // bytes per pixel row in smem
  static constexpr int A_BYTES = kHalo * kHalo * ROWB;                   // TMA box bytes
  static constexpr int A_STAGE = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_BYTES = BN * ROWB;
  static constexpr int B_STAGE = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int NACC = 2;                                         // TMEM accumulator stages
  static constexpr int TMEM_COLS = 2 * BN * NACC;                        // 2 M-blocks x BN x stages
  static constexpr int MAX_RING = 16;                                    // upper bound for sa, sb
  static constexpr int BAR_BYTES = (4 * MAX_RING + 2 * NACC + 2) * 8 + 16;
  // dynamic smem = 1024 (alignment slack) + sa*A_STAGE + (wres ? nchunks*9*B_BYTES : sb*B_STAGE) + BAR_BYTES
  static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM cols");
};

template <int KC, int BN, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_umma_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmA0,
                    const __grid_constant__ CUtensorMap tmA1) {
  using Cfg = ConvCfg<KC, BN>;
  constexpr int ROWB = Cfg::ROWB;
  constexpr int NACC = Cfg::NACC;
  const int SA = p.sa, SB = p.sb;
  const int nchunks = p.nchunks0 + p.nchunks1;
  const int b_region = p.wres ? nchunks * 9 * Cfg::B_BYTES : SB * Cfg::B_STAGE;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + SA * Cfg::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + ((b_region + 1023) & ~1023));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + Cfg::MAX_RING;
  uint64_t* b_full = a_empty + Cfg::MAX_RING;
  uint64_t* b_empty = b_full + Cfg::MAX_RING;
  uint64_t* acc_full = b_empty + Cfg::MAX_RING;
  uint64_t* acc_empty = acc_full + NACC;
  uint64_t* w_full = acc_empty + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.tiles_y * p.tiles_x * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kNumEpiWarps); }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int sa = 0, pa = 0, sb = 0, pb = 0;
      if (p.wres && int(blockIdx.x) < total_tiles) {
        // n_tiles == 1 here: the layer's packed weights are one contiguous run of nchunks*9 blobs
        const uint32_t wbytes = uint32_t(nchunks) * 9 * Cfg::B_BYTES;
        mbar_arrive_expect_tx(w_full, wbytes);
        for (uint32_t off = 0; off < wbytes; off += 9 * Cfg::B_BYTES)
          bulk_load_1d(b_smem + off, p.wpk + off, 9 * Cfg::B_BYTES, w_full);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t = tile;
        const int nt = t % p.n_tiles; t /= p.n_tiles;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int img = p.img0 + t / p.tiles_y;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_arrive_expect_tx(&a_full[sa], Cfg::A_BYTES);
          const bool seg0 = c < p.nchunks0;
          tma_load_4d(a_smem + sa * Cfg::A_STAGE, seg0 ? &tmA0 : &tmA1, &a_full[sa],
                      (seg0 ? c : c - p.nchunks0) * KC, tx * kTile - 1, ty * kTile - 1, img);
          if (++sa == SA) { sa = 0; pa ^= 1; }
          if (p.wres) continue;
          const uint8_t* wsrc = p.wpk + (size_t(c) * 9 * p.n_tiles + nt) * Cfg::B_BYTES;
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_arrive_expect_tx(&b_full[sb], Cfg::B_BYTES);
            bulk_load_1d(b_smem + sb * Cfg::B_STAGE, wsrc + size_t(tap) * p.n_tiles * Cfg::B_BYTES, Cfg::B_BYTES,
                         &b_full[sb]);
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // The whole warp walks the loop (warp-uniform control flow keeps descriptors in uniform registers);
    // only the tcgen05 instructions themselves are issued by one elected lane.
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    // high words of the smem descriptors are loop invariant: SBO | version | layout
    constexpr uint32_t kLayout = (ROWB == 128) ? 2u : 4u;
    constexpr uint32_t a_hi = (uint32_t(kHalo * ROWB) >> 4) | (1u << 14) | (kLayout << 29);
    constexpr uint32_t b_hi = (uint32_t(8 * ROWB) >> 4) | (1u << 14) | (kLayout << 29);
    int sa = 0, pa = 0, sb = 0, pb = 0;
    int it = 0;
    if (p.wres && int(blockIdx.x) < total_tiles) mbar_wait(w_full, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it % NACC;
      const uint32_t aph = (it / NACC) & 1;
      mbar_wait(&acc_empty[as], aph ^ 1);
      const uint32_t d0 = tmem_base + as * (2 * BN);
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        // descriptor low word: (addr >> 4) | LBO(=1) << 16 ; smem addresses are < 256 KB so no masking is needed
        const uint32_t a_lo0 = (smem_u32(a_smem + sa * Cfg::A_STAGE) >> 4) | (1u << 16);
        if (p.wres) {
          const uint32_t b_lo0 = (smem_u32(b_smem + c * 9 * Cfg::B_BYTES) >> 4) | (1u << 16);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_tap = a_lo0 + uint32_t(((tap / 3) * kHalo + (tap % 3)) * ROWB) / 16;
              const uint32_t b_tap = b_lo0 + uint32_t(tap * Cfg::B_BYTES) / 16;
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {   // alternate the two independent accumulators
                  umma_bf16_ss2(d0 + mb * BN, a_tap + uint32_t(mb * 8 * ROWB) / 16 + k * 2, a_hi, b_tap + k * 2, b_hi,
                                idesc, (c | tap | k) != 0 ? 1u : 0u);
                }
              }
            }
          }
          __syncwarp();
        } else {
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t b_tap = (smem_u32(b_smem + sb * Cfg::B_STAGE) >> 4) | (1u << 16);
            const uint32_t a_tap = a_lo0 + uint32_t(((tap / 3) * kHalo + (tap % 3)) * ROWB) / 16;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                  umma_bf16_ss2(d0 + mb * BN, a_tap + uint32_t(mb * 8 * ROWB) / 16 + k * 2, a_hi, b_tap + k * 2, b_hi,
                                idesc, (c | tap | k) != 0 ? 1u : 0u);
                }
              }
              tc_commit(&b_empty[sb]);   // weight slot free once these MMAs have read it
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
        if (elect_one()) tc_commit(&a_empty[sa]);     // halo slot free
        __syncwarp();
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
      if (elect_one()) tc_commit(&acc_full[as]);      // accumulators complete -> epilogue
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================== epilogue =========================================
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int mb = (warp - kEpiWarp0) >> 2;    // which M-block (left / right 8 columns of the tile)
    const int m = q * 32 + lane;               // row of the M=128 accumulator = pixel within the 16x8 block
    const int prow = m >> 3, pcol = (m & 7) + mb * 8;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int nt = t % p.n_tiles; t /= p.n_tiles;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int img = p.img0 + t / p.tiles_y;
      const int as = it % NACC;
      const uint32_t aph = (it / NACC) & 1;
      const int y = ty * kTile + prow, x = tx * kTile + pcol;
      const bool valid = (y < p.H) && (x < p.W);
      const size_t pix = (size_t(img) * p.H + y) * p.W + x;
      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * (2 * BN) + mb * BN;
      if constexpr (EPI == EPI_BF16) {
        __nv_bfloat16* optr = p.out + pix * p.Cout + nt * BN;
        const float* bptr = p.bias + nt * BN;
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + cc * 32, r);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bptr + cc * 32 + i));
            float v0 = __uint_as_float(r[i]) + b4.x, v1 = __uint_as_float(r[i + 1]) + b4.y;
            float v2 = __uint_as_float(r[i + 2]) + b4.z, v3 = __uint_as_float(r[i + 3]) + b4.w;
            v0 = v0 > 0.f ? v0 : v0 * p.slope; v1 = v1 > 0.f ? v1 : v1 * p.slope;
            v2 = v2 > 0.f ? v2 : v2 * p.slope; v3 = v3 > 0.f ? v3 : v3 * p.slope;
            o[i / 2] = pack_bf16x2(v0, v1);
            o[i / 2 + 1] = pack_bf16x2(v2, v3);
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(optr + cc * 32);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
            dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
          }
        }
      } else {
        static_assert(EPI == EPI_BF16 || BN == 32, "FINAL epilogue needs all 32 channels in one thread");
        uint32_t r[32];
        tmem_ld_32x32(taddr, r);
        tmem_ld_wait();
        float s = __ldg(p.bout);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = __uint_as_float(r[i]) + __ldg(p.bias + i);
          v = v > 0.f ? v : v * p.slope;
          s = fmaf(v, __ldg(p.wout + i), s);
        }
        if (valid) {
          const float o = __ldg(p.noisy + pix) + s;
          if (p.preclamp) p.preclamp[pix] = o;
          p.x_out[pix] = fminf(fmaxf(o, 0.f), 1.f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace pnp

// Warp-level shared-memory FFT core (power-of-two lengths 32..512), used by the FFT-prox kernels.
//
// One warp transforms G = max(1, 256/N) rows at a time, fully in place, with only __syncwarp() between
// the read and write half of each Stockham pass (every lane keeps its 8 (N<=256) or 16 (N=512) complex
// values in registers across the sync).  Rows live in shared memory with one padding float2 after every
// 8 elements (index i -> i + i/8) so that the stride-R writes of the first pass are conflict free.
//
// Only the forward transform is implemented; callers get the inverse through
//     IFFT(x) = conj(FFT(conj(x)))   (conjugations are folded into neighbouring pointwise steps).
//
// Twiddles: exp(-2*pi*i*k/512) computed in double precision on the host.  Each pass needs w^(k r), r < R:
// only w^k is looked up - from a per-pass COMPACT table (consecutive k -> consecutive words, so the lookup is
// bank-conflict free) - and the powers are formed by complex multiplication.  Complex arithmetic uses the
// sm_100 packed fp32x2 instructions (__fadd2_rn / __fmul2_rn / __ffma2_rn): one issue slot per complex add,
// three per complex multiply.
#pragma once
#include <cuda_runtime.h>

namespace pnp {

__device__ __forceinline__ int fpad(int i) { return i + (i >> 3); }
__host__ __device__ constexpr int fft_pitch(int n) { return n + (n >> 3); }   // float2 elements per padded row

// Shared-memory twiddle block: [0,512) full table, then compact tables for steps 2, 4, 8, 16.
constexpr int kTwFull = 512;
constexpr int kTwOff2 = 512;         // 64 entries: w512^(2k)
constexpr int kTwOff4 = 576;         // 32 entries: w512^(4k)
constexpr int kTwOff8 = 608;         //  8 entries: w512^(8k)
constexpr int kTwOff16 = 616;        //  8 entries: w512^(16k)
constexpr int kTwTotal = 624;

template <int STEP> __device__ __forceinline__ const float2* tw_compact(const float2* tw) {
  if constexpr (STEP == 1) return tw;
  else if constexpr (STEP == 2) return tw + kTwOff2;
  else if constexpr (STEP == 4) return tw + kTwOff4;
  else if constexpr (STEP == 8) return tw + kTwOff8;
  else return tw + kTwOff16;
}

// Fill the shared-memory twiddle block from the 512-entry global table (all threads of the CTA).
__device__ __forceinline__ void fft_load_twiddles(float2* tw, const float2* g_tw) {
  for (int k = threadIdx.x; k < kTwTotal; k += blockDim.x) {
    int src = k;
    if (k >= kTwOff16) src = (k - kTwOff16) * 16;
    else if (k >= kTwOff8) src = (k - kTwOff8) * 8;
    else if (k >= kTwOff4) src = (k - kTwOff4) * 4;
    else if (k >= kTwOff2) src = (k - kTwOff2) * 2;
    tw[k] = g_tw[src];
  }
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  // (a.x b.x - a.y b.y, a.x b.y + a.y b.x)
  const float2 t = __fmul2_rn(make_float2(a.x, a.x), b);
  return __ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), t);
}
// a + (-i) b  and  a - (-i) b   ((-i) b = (b.y, -b.x))
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) { return __ffma2_rn(make_float2(b.y, b.x), make_float2(1.f, -1.f), a); }
__device__ __forceinline__ float2 csub_mi(float2 a, float2 b) { return __ffma2_rn(make_float2(b.y, b.x), make_float2(-1.f, 1.f), a); }

// In-register forward DFTs, natural-order output.
__device__ __forceinline__ void dft2(float2& a, float2& b) {
  const float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}
__device__ __forceinline__ void dft4(float2 (&v)[4]) {
  const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
  const float2 b0 = cadd(v[1], v[3]), b1 = csub(v[1], v[3]);
  v[0] = cadd(a0, b0);
  v[2] = csub(a0, b0);
  v[1] = cadd_mi(a1, b1);     // a1 + (-i) b1
  v[3] = csub_mi(a1, b1);     // a1 - (-i) b1
}
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  // radix-2 split: evens / odds, then 4-point DFTs and twiddles w8^k
  float2 e[4] = {v[0], v[2], v[4], v[6]};
  float2 o[4] = {v[1], v[3], v[5], v[7]};
  dft4(e);
  dft4(o);
  const float h = 0.70710678118654752440f;
  // o1 * (1 - i)/sqrt2 = h (o1.x + o1.y, o1.y - o1.x);  o3 * (-1 - i)/sqrt2 = h (o3.y - o3.x, -(o3.x + o3.y))
  const float2 o1 = __fmul2_rn(__ffma2_rn(make_float2(o[1].y, o[1].x), make_float2(1.f, -1.f), o[1]), make_float2(h, h));
  const float2 o3 = __fmul2_rn(__ffma2_rn(make_float2(o[3].y, o[3].x), make_float2(1.f, -1.f),
                                          make_float2(-o[3].x, -o[3].y)), make_float2(h, h));
  v[0] = cadd(e[0], o[0]);     v[4] = csub(e[0], o[0]);
  v[1] = cadd(e[1], o1);       v[5] = csub(e[1], o1);
  v[2] = cadd_mi(e[2], o[2]);  v[6] = csub_mi(e[2], o[2]);
  v[3] = cadd(e[3], o3);       v[7] = csub(e[3], o3);
}
template <int R>
__device__ __forceinline__ void dftR(float2 (&v)[R]) {
  if constexpr (R == 2) dft2(v[0], v[1]);
  else if constexpr (R == 4) dft4(v);
  else dft8(v);
}

// One Stockham pass of radix R over a batch of G rows of length N owned by this warp.
//   Ns  = product of the radices of the passes already done (power of two)
//   tw  = shared-memory twiddle block (fft_load_twiddles)
// Index identities used (NB = N/R and Ns*R are multiples of 8 or the added term is < 8):
//   fpad(j + r*NB) = fpad(j) + r*fpad(NB)            read side
//   fpad(j0 + r*Ns) = fpad(j0) + r*(Ns + Ns/8)       write side, Ns >= 8
//   fpad(R*j + r)   = R*j + j*(R/8) + r              write side, Ns == 1, R == 8
template <int N, int R, int Ns, int G>
__device__ __forceinline__ void fft_warp_pass(float2* rows, int pitch, const float2* tw, int lane) {
  constexpr int NB = N / R;                 // butterflies per row
  constexpr int BF = (G * NB + 31) / 32;    // butterflies per lane
  static_assert((G * NB) % 32 == 0, "batch must fill the warp");
  static_assert(NB % 8 == 0 || NB < 8, "read-side index identity");
  float2 v[BF][R];
  float2* wbase[BF];
  int kk[BF];
#pragma unroll
  for (int i = 0; i < BF; ++i) {
    const int b = lane + 32 * i;
    const int row = b / NB, j = b % NB;
    const float2* src = rows + row * pitch + fpad(j);
    if constexpr (NB % 8 == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[i][r] = src[r * fft_pitch(NB)];
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) v[i][r] = rows[row * pitch + fpad(j + r * NB)];
    }
    const int k = j % Ns;
    kk[i] = k;
    const int j0 = (j / Ns) * (Ns * R) + k;
    wbase[i] = rows + row * pitch + ((Ns == 1) ? (R * j + j * (R / 8)) : fpad(j0));
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < BF; ++i) {
    if constexpr (Ns > 1) {
      constexpr int step = 512 / (Ns * R);
      const float2 w1 = tw_compact<step>(tw)[kk[i]];
      if constexpr (R == 2) {
        v[i][1] = cmul(v[i][1], w1);
      } else if constexpr (R == 4) {
        const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
        v[i][1] = cmul(v[i][1], w1); v[i][2] = cmul(v[i][2], w2); v[i][3] = cmul(v[i][3], w3);
      } else {
        const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
        const float2 w5 = cmul(w4, w1), w6 = cmul(w3, w3), w7 = cmul(w4, w3);
        v[i][1] = cmul(v[i][1], w1); v[i][2] = cmul(v[i][2], w2); v[i][3] = cmul(v[i][3], w3);
        v[i][4] = cmul(v[i][4], w4); v[i][5] = cmul(v[i][5], w5); v[i][6] = cmul(v[i][6], w6);
        v[i][7] = cmul(v[i][7], w7);
      }
    }
    dftR<R>(v[i]);
    if constexpr (Ns == 1) {
      static_assert(Ns != 1 || R == 8, "first pass is radix 8");
#pragma unroll
      for (int r = 0; r < R; ++r) wbase[i][r] = v[i][r];
    } else {
      static_assert(Ns == 1 || Ns % 8 == 0, "write-side index identity");
#pragma unroll
      for (int r = 0; r < R; ++r) wbase[i][r * (Ns + Ns / 8)] = v[i][r];
    }
  }
  __syncwarp();
}

template <int N> struct FftPlan;
template <> struct FftPlan<32>  { static constexpr int G = 8; };
template <> struct FftPlan<64>  { static constexpr int G = 4; };
template <> struct FftPlan<128> { static constexpr int G = 2; };
template <> struct FftPlan<256> { static constexpr int G = 1; };
template <> struct FftPlan<512> { static constexpr int G = 1; };

// Forward FFT (unnormalised) of FftPlan<N>::G consecutive padded rows starting at `rows`.
template <int N>
__device__ __forceinline__ void fft_warp_rows(float2* rows, int pitch, const float2* tw, int lane) {
  constexpr int G = FftPlan<N>::G;
  if constexpr (N == 32) {
    fft_warp_pass<32, 8, 1, G>(rows, pitch, tw, lane);
    fft_warp_pass<32, 4, 8, G>(rows, pitch, tw, lane);
  } else if constexpr (N == 64) {
    fft_warp_pass<64, 8, 1, G>(rows, pitch, tw, lane);
    fft_warp_pass<64, 8, 8, G>(rows, pitch, tw, lane);
  } else if constexpr (N == 128) {
    fft_warp_pass<128, 8, 1, G>(rows, pitch, tw, lane);
    fft_warp_pass<128, 4, 8, G>(rows, pitch, tw, lane);
    fft_warp_pass<128, 4, 32, G>(rows, pitch, tw, lane);
  } else if constexpr (N == 256) {
    fft_warp_pass<256, 8, 1, G>(rows, pitch, tw, lane);
    fft_warp_pass<256, 8, 8, G>(rows, pitch, tw, lane);
    fft_warp_pass<256, 4, 64, G>(rows, pitch, tw, lane);
  } else {
    fft_warp_pass<512, 8, 1, G>(rows, pitch, tw, lane);
    fft_warp_pass<512, 8, 8, G>(rows, pitch, tw, lane);
    fft_warp_pass<512, 8, 64, G>(rows, pitch, tw, lane);
  }
}

}  // namespace pnp

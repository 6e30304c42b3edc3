// Warp-level shared-memory FFT core (power-of-two lengths 32..512), used by the FFT-prox kernels.
//
// One warp transforms G = max(1, 256/N) rows at a time, fully in place, with only __syncwarp() between
// the read and write half of each Stockham pass (every lane keeps its 8 (N<=256) or 16 (N=512) complex
// values in registers across the sync).  Rows live in shared memory with one padding float2 after every
// 8 elements (index i -> i + i/8) so that the stride-R writes of the first pass are conflict free.
//
// Only the forward transform is implemented; callers get the inverse through
//     IFFT(x) = conj(FFT(conj(x)))   (conjugations are folded into neighbouring pointwise steps).
// Twiddles come from a 512-entry table exp(-2*pi*i*k/512) computed in double precision on the host.
#pragma once
#include <cuda_runtime.h>

namespace pnp {

__device__ __forceinline__ int fpad(int i) { return i + (i >> 3); }
__host__ __device__ constexpr int fft_pitch(int n) { return n + (n >> 3); }   // float2 elements per padded row

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// In-register forward DFTs, natural-order output.
__device__ __forceinline__ void dft2(float2& a, float2& b) {
  const float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}
__device__ __forceinline__ void dft4(float2 (&v)[4]) {
  float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
  float2 b0 = cadd(v[1], v[3]), b1 = mul_mi(csub(v[1], v[3]));
  v[0] = cadd(a0, b0);
  v[2] = csub(a0, b0);
  v[1] = cadd(a1, b1);
  v[3] = csub(a1, b1);
}
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  // radix-2 split: evens / odds, then 4-point DFTs and twiddles w8^k
  float2 e[4] = {v[0], v[2], v[4], v[6]};
  float2 o[4] = {v[1], v[3], v[5], v[7]};
  dft4(e);
  dft4(o);
  const float h = 0.70710678118654752440f;
  const float2 o1 = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));     // * (1 - i)/sqrt2
  const float2 o2 = mul_mi(o[2]);                                                     // * (-i)
  const float2 o3 = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));    // * (-1 - i)/sqrt2
  v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
  v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
  v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
  v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
}
template <int R>
__device__ __forceinline__ void dftR(float2 (&v)[R]) {
  if constexpr (R == 2) dft2(v[0], v[1]);
  else if constexpr (R == 4) dft4(v);
  else dft8(v);
}

// One Stockham pass of radix R over a batch of G rows of length N owned by this warp.
//   Ns  = product of the radices of the passes already done
//   tw  = shared-memory copy of the 512-entry table
template <int N, int R, int Ns, int G>
__device__ __forceinline__ void fft_warp_pass(float2* rows, int pitch, const float2* tw, int lane) {
  constexpr int NB = N / R;                 // butterflies per row
  constexpr int BF = (G * NB + 31) / 32;    // butterflies per lane
  static_assert((G * NB) % 32 == 0 || G * NB < 32, "batch must fill the warp");
  float2 v[BF][R];
#pragma unroll
  for (int i = 0; i < BF; ++i) {
    const int b = lane + 32 * i;
    const int row = b / NB, j = b % NB;
    if (G * NB >= 32 || b < G * NB) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[i][r] = rows[row * pitch + fpad(j + r * NB)];
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < BF; ++i) {
    const int b = lane + 32 * i;
    const int row = b / NB, j = b % NB;
    if (G * NB >= 32 || b < G * NB) {
      if constexpr (Ns > 1) {
        const int k = j % Ns;
        constexpr int step = 512 / (Ns * R);
#pragma unroll
        for (int r = 1; r < R; ++r) v[i][r] = cmul(v[i][r], tw[(k * r * step) & 511]);
      }
      dftR<R>(v[i]);
      const int j0 = (j / Ns) * (Ns * R) + (j % Ns);
#pragma unroll
      for (int r = 0; r < R; ++r) rows[row * pitch + fpad(j0 + r * Ns)] = v[i][r];
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

/* A C host that LAUNCHES kernels through the C-ABI (include/pnp_b200.h) and checks them against plain-C arithmetic: the
 * reference's centred transforms (evaluation/utils/transformations.py:6-19) as O(N^4) double-precision sums, its masked
 * k-space solve and dual update (evaluation/env.py:87-93) and its PSNR (env.py:120-125).  No Python, no torch: cudaMalloc /
 * cudaMemcpy from the CUDA runtime, the library for everything else.  Needs a B200.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/host_prox.c -o host_prox \
 *       -L dt4image_restoration_b200/csrc -lpnp_b200 -Wl,-rpath,$PWD/dt4image_restoration_b200/csrc \
 *       -L /usr/local/cuda/lib64 -lcudart -lm
 *
 * Cases: 32 x 32 (radix kernels, three-launch general path) and 18 x 24 (dense-DFT any-size path), B = 2, per-image masks.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "pnp_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %d at line %d\n", (int)e_, __LINE__); return 2; } } while (0)
#define CHECK_PNP(x) do { int r_ = (x); if (r_ != 0) { printf("pnp error %d at line %d: %s\n", r_, __LINE__, pnp_last_error()); return 3; } } while (0)

static const double kPi = 3.14159265358979323846;

static unsigned int rng_state = 12345u;
static double frand(void) {                       /* xorshift, [0, 1) */
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5;
  return (double)(rng_state & 0xffffffu) / 16777216.0;
}

/* centred orthonormal 2-D DFT of one H x W image, out[k1][k2] = 1/sqrt(HW) sum in[n1][n2] w^((n1-h1)(k1-h1)) ... (see
 * csrc/fftprox_any.cuh for why this equals fftshift(fft2(ifftshift(.)))); sign = -1 forward, +1 inverse */
static void dft2c(const double* in_re, const double* in_im, double* out_re, double* out_im, int H, int W, int sign) {
  const int h1 = H / 2, h2 = W / 2;
  double* tr = (double*)malloc(sizeof(double) * H * W);
  double* ti = (double*)malloc(sizeof(double) * H * W);
  int i, j, k;
  for (i = 0; i < H; ++i)
    for (k = 0; k < W; ++k) {                     /* rows */
      double sr = 0, si = 0;
      for (j = 0; j < W; ++j) {
        const double a = sign * 2.0 * kPi * (double)((j - h2) * (k - h2)) / W;
        const double c = cos(a), s = sin(a);
        sr += in_re[i * W + j] * c - in_im[i * W + j] * s;
        si += in_re[i * W + j] * s + in_im[i * W + j] * c;
      }
      tr[i * W + k] = sr / sqrt((double)W); ti[i * W + k] = si / sqrt((double)W);
    }
  for (k = 0; k < H; ++k)
    for (j = 0; j < W; ++j) {                     /* columns */
      double sr = 0, si = 0;
      for (i = 0; i < H; ++i) {
        const double a = sign * 2.0 * kPi * (double)((i - h1) * (k - h1)) / H;
        const double c = cos(a), s = sin(a);
        sr += tr[i * W + j] * c - ti[i * W + j] * s;
        si += tr[i * W + j] * s + ti[i * W + j] * c;
      }
      out_re[k * W + j] = sr / sqrt((double)H); out_im[k * W + j] = si / sqrt((double)H);
    }
  free(tr); free(ti);
}

static int run_case(int B, int H, int W) {
  const int n = H * W, N = B * n;
  float* x = (float*)malloc(sizeof(float) * N);
  float* gt = (float*)malloc(sizeof(float) * N);
  float* u = (float*)malloc(sizeof(float) * 2 * N);
  float* y0 = (float*)malloc(sizeof(float) * 2 * N);
  unsigned char* mask = (unsigned char*)malloc(N);
  float mu_h[2] = {0.37f, 0.81f};
  float *z_h = (float*)malloc(sizeof(float) * 2 * N), *un_h = (float*)malloc(sizeof(float) * 2 * N);
  float *v_h = (float*)malloc(sizeof(float) * N), *f_h = (float*)malloc(sizeof(float) * 2 * N), psnr_h[2];
  double *wr = (double*)malloc(sizeof(double) * n), *wi = (double*)malloc(sizeof(double) * n);
  double *Zr = (double*)malloc(sizeof(double) * n), *Zi = (double*)malloc(sizeof(double) * n);
  double *zr = (double*)malloc(sizeof(double) * n), *zi = (double*)malloc(sizeof(double) * n);
  float *d_x, *d_gt, *d_mu, *d_v, *d_psnr;
  void *d_u, *d_y0, *d_z, *d_un, *d_work, *d_f;
  unsigned char* d_mask;
  double err_z = 0, err_u = 0, err_v = 0, err_f = 0, err_p = 0;
  int i, b;
  if (B > 2) return 1;
  for (i = 0; i < N; ++i) {
    x[i] = (float)frand(); gt[i] = (float)frand();
    u[2 * i] = (float)(0.2 * frand() - 0.1); u[2 * i + 1] = (float)(0.2 * frand() - 0.1);
    y0[2 * i] = (float)(2 * frand() - 1); y0[2 * i + 1] = (float)(2 * frand() - 1);
    mask[i] = frand() < 0.3 ? 1 : 0;
  }
  CHECK_CUDA(cudaMalloc((void**)&d_x, sizeof(float) * N)); CHECK_CUDA(cudaMalloc((void**)&d_gt, sizeof(float) * N));
  CHECK_CUDA(cudaMalloc(&d_u, sizeof(float) * 2 * N)); CHECK_CUDA(cudaMalloc(&d_y0, sizeof(float) * 2 * N));
  CHECK_CUDA(cudaMalloc(&d_z, sizeof(float) * 2 * N)); CHECK_CUDA(cudaMalloc(&d_un, sizeof(float) * 2 * N));
  CHECK_CUDA(cudaMalloc(&d_f, sizeof(float) * 2 * N));
  CHECK_CUDA(cudaMalloc((void**)&d_v, sizeof(float) * N)); CHECK_CUDA(cudaMalloc((void**)&d_mask, N));
  CHECK_CUDA(cudaMalloc((void**)&d_mu, sizeof(float) * 2)); CHECK_CUDA(cudaMalloc((void**)&d_psnr, sizeof(float) * 2));
  CHECK_CUDA(cudaMalloc(&d_work, pnp_prox_workspace_bytes(B, H, W)));
  CHECK_CUDA(cudaMemcpy(d_x, x, sizeof(float) * N, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_gt, gt, sizeof(float) * N, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_u, u, sizeof(float) * 2 * N, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_y0, y0, sizeof(float) * 2 * N, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_mask, mask, N, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_mu, mu_h, sizeof(float) * 2, cudaMemcpyHostToDevice));
  /* one prox + dual step per image mu, the centred FFT of u, the reward of x - all on the default stream */
  CHECK_PNP(pnp_prox_dual(d_x, d_u, d_y0, d_mask, (long long)n, d_mu, 1, d_z, d_un, d_v, d_work, B, H, W, NULL));
  CHECK_PNP(pnp_fft2c(d_u, d_f, B, H, W, 0, NULL));
  CHECK_PNP(pnp_psnr(d_x, d_gt, (long long)n, d_psnr, B, n, NULL));
  CHECK_CUDA(cudaDeviceSynchronize());
  CHECK_CUDA(cudaMemcpy(z_h, d_z, sizeof(float) * 2 * N, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(un_h, d_un, sizeof(float) * 2 * N, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(v_h, d_v, sizeof(float) * N, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(f_h, d_f, sizeof(float) * 2 * N, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(psnr_h, d_psnr, sizeof(float) * B, cudaMemcpyDeviceToHost));
  for (b = 0; b < B; ++b) {
    const double mu = mu_h[b];
    double mse = 0;
    for (i = 0; i < n; ++i) { wr[i] = (double)x[b * n + i] + u[2 * (b * n + i)]; wi[i] = u[2 * (b * n + i) + 1]; }
    dft2c(wr, wi, Zr, Zi, H, W, -1);                                     /* env.py:87 */
    for (i = 0; i < n; ++i)
      if (mask[b * n + i]) {                                             /* env.py:88-90 */
        Zr[i] = (mu * Zr[i] + y0[2 * (b * n + i)]) / (1 + mu);
        Zi[i] = (mu * Zi[i] + y0[2 * (b * n + i) + 1]) / (1 + mu);
      }
    dft2c(Zr, Zi, zr, zi, H, W, +1);                                     /* env.py:91 */
    for (i = 0; i < n; ++i) {
      const double ur = wr[i] - zr[i], ui = wi[i] - zi[i];               /* env.py:93: u + x - z */
      const double xc = x[b * n + i] < 0 ? 0 : (x[b * n + i] > 1 ? 1 : x[b * n + i]);
      err_z = fmax(err_z, fmax(fabs(z_h[2 * (b * n + i)] - zr[i]), fabs(z_h[2 * (b * n + i) + 1] - zi[i])));
      err_u = fmax(err_u, fmax(fabs(un_h[2 * (b * n + i)] - ur), fabs(un_h[2 * (b * n + i) + 1] - ui)));
      err_v = fmax(err_v, fabs(v_h[b * n + i] - (zr[i] - ur)));
      mse += (xc - gt[b * n + i]) * (xc - gt[b * n + i]);
    }
    err_p = fmax(err_p, fabs(psnr_h[b] - 10.0 * log10(1.0 / (mse / n))));  /* env.py:120-125 */
    for (i = 0; i < n; ++i) { wr[i] = u[2 * (b * n + i)]; wi[i] = u[2 * (b * n + i) + 1]; }
    dft2c(wr, wi, Zr, Zi, H, W, -1);
    for (i = 0; i < n; ++i)
      err_f = fmax(err_f, fmax(fabs(f_h[2 * (b * n + i)] - Zr[i]), fabs(f_h[2 * (b * n + i) + 1] - Zi[i])));
  }
  printf("%dx%d B=%d: max|z| %.2e max|u| %.2e max|v| %.2e max|fft| %.2e max|psnr| %.2e dB\n", H, W, B, err_z, err_u, err_v,
         err_f, err_p);
  cudaFree(d_x); cudaFree(d_gt); cudaFree(d_u); cudaFree(d_y0); cudaFree(d_z); cudaFree(d_un); cudaFree(d_f); cudaFree(d_v);
  cudaFree(d_mask); cudaFree(d_mu); cudaFree(d_psnr); cudaFree(d_work);
  free(x); free(gt); free(u); free(y0); free(mask); free(z_h); free(un_h); free(v_h); free(f_h);
  free(wr); free(wi); free(Zr); free(Zi); free(zr); free(zi);
  return (err_z < 2e-5 && err_u < 2e-5 && err_v < 4e-5 && err_f < 2e-5 && err_p < 1e-3) ? 0 : 1;
}

int main(void) {
  int rc;
  CHECK_PNP(pnp_init());
  rc = run_case(2, 32, 32);
  if (rc == 0) rc = run_case(2, 18, 24);
  printf(rc == 0 ? "host_prox: ok\n" : "host_prox: MISMATCH\n");
  return rc;
}

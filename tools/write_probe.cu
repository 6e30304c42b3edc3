// Pure-write HBM bandwidth probe: how fast can 1 GiB be written by (a) 16-byte stores, (b) streaming 16-byte stores,
// (c) bulk-async stores from shared memory (cp.async.bulk.global.shared::cta), next to a device-to-device copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_probe tools/write_probe.cu && tools/write_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_st(uint4* dst, size_t n) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) dst[i] = v;
}
__global__ void k_st_cs(uint4* dst, size_t n) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) __stcs(dst + i, v);
}
// each CTA owns a 16 KB staging tile in shared memory and streams it out with bulk stores
__global__ void k_bulk(uint8_t* dst, size_t bytes) {
  extern __shared__ __align__(128) uint8_t tile[];
  constexpr uint32_t T = 16384;
  for (uint32_t i = threadIdx.x * 16; i < T; i += blockDim.x * 16) *reinterpret_cast<uint4*>(tile + i) = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t s = uint32_t(__cvta_generic_to_shared(tile));
    int inflight = 0;
    for (size_t off = size_t(blockIdx.x) * T; off + T <= bytes; off += size_t(gridDim.x) * T) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(s), "r"(T) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const size_t bytes = size_t(1) << 30;
  uint8_t *a, *b;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn, double factor) {
    for (int i = 0; i < 3; ++i) fn();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) fn();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.1f GB/s  (%s)\n", name, factor * bytes * 10 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  for (int bpsm : {4, 8, 16}) {
    const int grid = 148 * bpsm;
    char nm[96];
    snprintf(nm, sizeof nm, "st.global.v4, %d CTAs x 256", grid);
    timeit(nm, [&] { k_st<<<grid, 256>>>(reinterpret_cast<uint4*>(a), bytes / 16); }, 1.0);
    snprintf(nm, sizeof nm, "st.global.cs.v4, %d CTAs x 256", grid);
    timeit(nm, [&] { k_st_cs<<<grid, 256>>>(reinterpret_cast<uint4*>(a), bytes / 16); }, 1.0);
  }
  for (int bpsm : {2, 4, 8}) {
    const int grid = 148 * bpsm;
    char nm[96];
    snprintf(nm, sizeof nm, "cp.async.bulk smem->global 16 KB, %d CTAs", grid);
    timeit(nm, [&] { k_bulk<<<grid, 128, 16384>>>(a, bytes); }, 1.0);
  }
  timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(a, 1, bytes); }, 1.0);
  timeit("cudaMemcpyAsync d2d (read + write)", [&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 2.0);
  return 0;
}

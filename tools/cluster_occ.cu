// How many thread-block clusters of a given shape can be resident on this GPU at once?  (cudaOccupancyMaxActiveClusters)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/cluster_occ tools/cluster_occ.cu && ./tools/cluster_occ
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int cls[] = {1, 2, 4, 8, 16};
  const int smems[] = {230 * 1024, 115 * 1024, 113 * 1024, 75 * 1024, 56 * 1024, 37 * 1024};
  for (int pol = 0; pol < 3; ++pol)
  for (int cl : cls) for (int sm : smems) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cl * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = sm;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
    at[1].val.clusterSchedulingPolicyPreference = pol == 0 ? cudaClusterSchedulingPolicyDefault : (pol == 1 ? cudaClusterSchedulingPolicySpread : cudaClusterSchedulingPolicyLoadBalancing);
    cfg.attrs = at; cfg.numAttrs = 2;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("policy %d cluster %2d smem %6d B/CTA (256 thr): max active clusters %3d = %4d CTAs (%s)\n", pol, cl, sm, n, n * cl, cudaGetErrorString(e));
    (void)cudaGetLastError();
  }
  return 0;
}

// Probe: how many thread-block clusters of each size are co-resident for a 1-CTA/SM persistent kernel.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int smem : {100 * 1024, 200 * 1024, 220 * 1024}) {
    for (int cs : {1, 2, 4, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 32);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute a[1];
      a[0].id = cudaLaunchAttributeClusterDimension;
      a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      cfg.attrs = a; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %3d -> %3d CTAs (%s)\n", smem / 1024, cs, n, n * cs, cudaGetErrorString(e));
    }
  }
  return 0;
}

// How many thread-block clusters of size 2/4/6/8 (1 CTA per SM: 200 KB dynamic smem, 320 threads) can be co-resident.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  for (int cs : {1, 2, 3, 4, 6, 8, 12, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute a; a.id = cudaLaunchAttributeClusterDimension; a.val.clusterDim.x = cs; a.val.clusterDim.y = 1; a.val.clusterDim.z = 1;
    cfg.attrs = &a; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %d (%d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaGetLastError();
  }
  return 0;
}

// Probe (GPU box): how many clusters of 1 / 2 / 4 / 8 CTAs with this library's GEMM footprint (one CTA per SM: ~220 KB of
// dynamic shared memory, 576 threads) can be co-resident -- i.e. how many of the 148 SMs a 4-CTA-cluster kernel can use.
//   nvcc -arch=sm_100a -o /tmp/cluster_occupancy tools/cluster_occupancy.cu && /tmp/cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(576, 1) probe_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = smem[0];
}

int main() {
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs: %d\n", sms);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / cs * cs);
    cfg.blockDim = dim3(576);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d SMs usable (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}

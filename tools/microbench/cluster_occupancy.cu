// How many clusters of 1/2/4/8 CTAs (256 threads, 127 registers, 104 KB of dynamic shared memory: the pair-tile kernel's
// footprint) can be resident on this GPU at once?   nvcc -arch=sm_100a -o cluster_occupancy cluster_occupancy.cu && ./cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) dummy(float* p) {
  extern __shared__ float s[];
  float acc[96];
#pragma unroll
  for (int i = 0; i < 96; i++) acc[i] = p[i] * s[i];
  float t = 0;
#pragma unroll
  for (int i = 0; i < 96; i++) t += acc[i] * acc[(i * 7) % 96];
  p[threadIdx.x] = t;
}
int main() {
  const int smem = 104 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, dummy);
  printf("dummy kernel: %d registers\n", fa.numRegs);
  for (int cs : {1, 2, 4, 8}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1184);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %d: max active clusters %d (%d CTAs)  [%s]\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}

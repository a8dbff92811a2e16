#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 3) probe(const int* flag, float* out) {
  extern __shared__ float s[];
  if (*flag == 0) return;
  s[threadIdx.x] = 1.0f;
  __syncthreads();
  out[blockIdx.x] = s[(threadIdx.x + 1) & 511];
}
int main() {
  int* flag; float* out;
  cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4); cudaMalloc(&out, 1 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int smems[] = {1024, 76000};
  const int grids[] = {8192, 444, 65536};
  for (int smem : smems) {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : grids) {
      for (int i = 0; i < 5; ++i) probe<<<grid, 512, smem>>>(flag, out);
      cudaEventRecord(e0);
      for (int i = 0; i < 100; ++i) probe<<<grid, 512, smem>>>(flag, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("smem %6d grid %6d : %.2f us per empty launch\n", smem, grid, ms * 10.0f);
    }
  }
  return 0;
}

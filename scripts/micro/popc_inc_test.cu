// Does a no-return shared atomicAdd(+1) with per-lane addresses (ptxas: ATOMS.POPC.INC.32) count correctly?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void hist(const uint32_t* keys, int m, uint32_t* out, int mode) {
  __shared__ uint32_t bcnt[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) bcnt[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    uint32_t b = keys[i] >> 23;
    if (mode == 0) atomicAdd(&bcnt[b], 1u);
    else if (mode == 1) { uint32_t o = atomicAdd(&bcnt[b], 1u); if (o == 0xdeadbeefu) out[300] = 1; }
    else asm volatile("red.shared.add.u32 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bcnt[b])) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = bcnt[i];
}
int main() {
  const int m = 60;
  uint32_t h[m];
  uint32_t seed = 12345;
  for (int i = 0; i < m; i++) { seed = seed * 1664525u + 1013904223u; h[i] = seed >> 3; }   // 29-bit keys
  uint32_t *dk, *dout; cudaMalloc(&dk, sizeof(h)); cudaMalloc(&dout, 1024 * 4);
  cudaMemcpy(dk, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 3; mode++) {
    hist<<<1, 256>>>(dk, m, dout, mode);
    uint32_t o[256]; cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost);
    uint32_t exp[256] = {0}; for (int i = 0; i < m; i++) exp[h[i] >> 23]++;
    int bad = 0, tot = 0; for (int b = 0; b < 256; b++) { bad += (o[b] != exp[b]); tot += o[b]; }
    printf("mode %d: bad bins %d, total %d (expected %d) %s\n", mode, bad, tot, m, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

// Microbenchmark: shared-memory atomics vs plain LDS/STS throughput on one SM-filling grid.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_bench atoms_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int TS = 4096;
template <int MODE>
__global__ void k(uint32_t* out, int iters) {
  __shared__ unsigned long long tab[TS];
  uint32_t* tab32 = reinterpret_cast<uint32_t*>(tab);
  for (int i = threadIdx.x; i < TS; i += blockDim.x) tab[i] = ~0ull;
  __syncthreads();
  uint32_t key = threadIdx.x * 2654435761u + blockIdx.x;
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    key = key * 1664525u + 1013904223u;
    uint32_t slot = (key >> 8) & (TS - 1);
    if (MODE == 0) acc += atomicCAS(&tab32[slot], 0xffffffffu, key);
    if (MODE == 1) acc += (uint32_t)atomicCAS(&tab[slot], ~0ull, (unsigned long long)key);
    if (MODE == 2) acc += atomicAdd(&tab32[slot], 1u);
    if (MODE == 3) { atomicAdd(&tab32[slot], 1u); }                 // no return (RED)
    if (MODE == 4) { tab32[slot] = key; acc += tab32[(slot + 7) & (TS - 1)]; }   // STS + LDS
    if (MODE == 5) acc += __match_any_sync(0xffffffffu, key & 1023u);
    if (MODE == 6) acc += atomicMin(&tab32[slot], key);
    if (MODE == 7) acc += atomicExch(&tab32[slot], key);
    if (MODE == 8) acc += atomicOr(&tab32[slot], 1u << (key & 31));
    if (MODE == 9) { atomicOr(&tab32[slot], 1u << (key & 31)); }
    if (MODE == 10) { uint32_t b = 1u << (2 * (key & 15)); uint32_t o = atomicOr(&tab32[slot], b); if (o & b) acc += atomicOr(&tab32[slot], b << 1); if ((it & 63) == 63) tab32[slot] = 0; }
    if (MODE == 11) { acc += tab32[slot]; }                          // LDS only
    if (MODE == 12) { if (key & 0x30000) acc += atomicAdd(&tab32[slot], 1u); }   // 75 % of the lanes active
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, uint32_t* d, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int blocks = 148 * 4, threads = 256;
  k<MODE><<<blocks, threads>>>(d, iters);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(d, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double warp_ops = (double)blocks * threads / 32 * iters;
  // cycles per warp-op per SM assuming 1.9 GHz
  double cyc = ms * 1e-3 * 1.9e9 / (warp_ops / 148);
  printf("%-28s %8.3f ms  %6.2f SM-cycles per warp-op\n", name, ms, cyc);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  int iters = 4096;
  run<0>("atomicCAS u32 (return)", d, iters);
  run<1>("atomicCAS u64 (return)", d, iters);
  run<2>("atomicAdd u32 (return)", d, iters);
  run<3>("atomicAdd u32 (no return)", d, iters);
  run<4>("STS + LDS", d, iters);
  run<5>("match_any", d, iters);
  run<6>("atomicMin u32", d, iters);
  run<7>("atomicExch u32", d, iters);
  run<8>("atomicOr u32 (return)", d, iters);
  run<9>("atomicOr u32 (no return)", d, iters);
  run<10>("atomicOr 2-level nibble", d, iters);
  run<11>("LDS random", d, iters);
  run<12>("atomicAdd 75% lanes", d, iters);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

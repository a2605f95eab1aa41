// Are no-return shared-memory atomics ordered by __syncthreads() on sm_100a?
// Each thread ORs its bit pattern into pseudo-random slots (no return value used), barrier,
// then every thread checks slots written by OTHER warps.  Prints the number of violations.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TS = 2048;
template <int MODE>
__global__ void k(unsigned long long* bad, int iters) {
  __shared__ uint32_t tab[TS];
  __shared__ uint32_t info[TS];
  const int tid = threadIdx.x;
  unsigned long long nbad = 0;
  for (int it = 0; it < iters; it++) {
    for (int i = tid; i < TS; i += blockDim.x) { tab[i] = 0xffffffffu; info[i] = 0; }
    __syncthreads();
    // phase 1: 4 keys per thread, CAS-insert with linear probing, owner ORs its id (no return)
    uint32_t myslot[4];
    for (int e = 0; e < 4; e++) {
      uint32_t key = (uint32_t)(tid * 4 + e) * 2654435761u + it * 40503u + blockIdx.x;
      key &= 0x7fffffffu;
      uint32_t c = (key * 0x9E3779B1u) >> 21;
      for (;;) {
        uint32_t old = atomicCAS(&tab[c], 0xffffffffu, key);
        if (old == 0xffffffffu) { atomicOr(&info[c], (uint32_t)(tid * 4 + e) + 1u); myslot[e] = c; break; }
        if (old == key) { atomicOr(&info[c], 1u << 30); myslot[e] = 0xffffffffu; break; }
        c = (c + 1) & (TS - 1);
      }
    }
    if (MODE == 1) __threadfence_block();
    if (MODE == 2) { uint32_t v = atomicAdd(&info[tid], 0u); if (v == 0xdeadbeefu) bad[1] = 1; }
    __syncthreads();
    // phase 2: look up the keys of thread (tid + 37) % 256 and check the owner's id is visible
    const int other = (tid + 37) & 255;
    for (int e = 0; e < 4; e++) {
      uint32_t key = (uint32_t)(other * 4 + e) * 2654435761u + it * 40503u + blockIdx.x;
      key &= 0x7fffffffu;
      uint32_t c = (key * 0x9E3779B1u) >> 21;
      for (;;) {
        uint32_t t = tab[c];
        if (t == 0xffffffffu) { nbad++; break; }            // key not found: insert not visible
        if (t == key) { if ((info[c] & 0xfffu) != (uint32_t)(other * 4 + e) + 1u) nbad++; break; }
        c = (c + 1) & (TS - 1);
      }
    }
    __syncthreads();
  }
  if (nbad) atomicAdd(bad, nbad);
}
template <int MODE> void run(const char* name) {
  unsigned long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  k<MODE><<<148 * 8, 256>>>(d, 200);
  unsigned long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-34s violations: %llu  (%s)\n", name, h, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>("plain __syncthreads");
  run<1>("__threadfence_block + barrier");
  run<2>("returning atomic + barrier");
  run<0>("plain __syncthreads (again)");
  return 0;
}

// int_pipe_bench.cu -- issue cost of the integer instructions kernels A1 / A2 lean on, in SM-cycles per warp instruction
// with every scheduler saturated (32 resident warps per SM, eight independent chains per thread).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/int_pipe_bench scripts/micro/int_pipe_bench.cu
// Development tool only (profiles/r02_int_pipe_bench.txt); nothing in the product path uses it.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 8, kIters = 512;

template <int OP>
__global__ void __launch_bounds__(256) bench(unsigned* out, unsigned seed, unsigned mul, long long* cycles) {
  unsigned x[kChains];
  unsigned long long w[kChains];
#pragma unroll
  for (int k = 0; k < kChains; k++) { x[k] = seed + threadIdx.x * 7u + k; w[k] = x[k]; }
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < kIters; i++) {
#pragma unroll
    for (int k = 0; k < kChains; k++) {
      if (OP == 0) x[k] = x[k] * mul + seed;                                         // IMAD (low 32 bits)
      if (OP == 1) x[k] = __umulhi(x[k], mul) + seed;                                // IMAD.HI
      if (OP == 2) w[k] = (unsigned long long)(unsigned)w[k] * mul + w[k];           // IMAD.WIDE, 64-bit addend
      if (OP == 3) x[k] = (x[k] & mul) ^ seed;                                       // LOP3
      if (OP == 4) x[k] = __byte_perm(x[k], seed, mul);                              // PRMT
      if (OP == 5) x[k] = __vabsdiffu4(x[k], mul);                                   // VABSDIFF4
      if (OP == 6) x[k] = __dp4a(x[k], mul, seed);                                   // IDP.4A
      if (OP == 7) x[k] = __usad(x[k], mul, seed);                                   // VABSDIFF (+ add)
      if (OP == 8) x[k] = __funnelshift_r(x[k], seed, mul);                          // SHF
      if (OP == 9) x[k] = x[k] + mul + seed;                                         // IADD3
    }
  }
  const long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < kChains; k++) s ^= x[k] ^ (unsigned)w[k] ^ (unsigned)(w[k] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, unsigned* out, long long* cyc, int sms) {
  const int ctas = sms * 4;                                     // 4 CTAs x 8 warps = 32 warps per SM
  bench<OP><<<ctas, 256>>>(out, 12345u, 0x9e3779b1u, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<ctas, 256>>>(out, 12345u, 0x9e3779b1u, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[4096];
  cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < ctas; i++) mean += (double)h[i];
  mean /= ctas;
  const double warp_inst_per_sm = 32.0 * kChains * kIters;     // per SM
  printf("%-10s %7.3f SM-cycles per warp instruction (%.0f cycles per CTA, %.3f ms)\n", name, mean / warp_inst_per_sm, mean, ms);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* out; long long* cyc;
  cudaMalloc(&out, sizeof(unsigned) * sms * 4 * 256);
  cudaMalloc(&cyc, sizeof(long long) * sms * 4);
  run<0>("IMAD", out, cyc, sms); run<1>("IMAD.HI", out, cyc, sms); run<2>("IMAD.WIDE", out, cyc, sms); run<3>("LOP3", out, cyc, sms);
  run<4>("PRMT", out, cyc, sms); run<5>("VABSDIFF4", out, cyc, sms); run<6>("IDP.4A", out, cyc, sms); run<7>("VABSDIFF", out, cyc, sms);
  run<8>("SHF", out, cyc, sms); run<9>("IADD3", out, cyc, sms);
  return 0;
}

// pyramid.cu -- image pyramid construction for the multi-level configuration (BASELINE.json
// configs[3]).  The reference has no pyramid; SURVEY.md 8d defines it: level l+1 is the 2x2 mean
// with floor, (a + b + c + d) / 4, of the RAW level-l image, and the single-level path (box,
// Sobel, hashing, matching) then runs independently per level.
#include <algorithm>

#include "gpc_device.cuh"

namespace gpc {

// one thread = 4 output pixels of one row (reads 2 x 8 source bytes, writes one word)
__global__ void __launch_bounds__(256)
downsample2x_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int sw, int dw, int dh, int n_img, size_t sP, size_t dP) {
  const int quads = dw / 4;
  const long long total = (long long)n_img * dh * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % quads);
    const int y = (int)((i / quads) % dh);
    const int img = (int)(i / ((long long)quads * dh));
    const uint8_t* s0 = src + (size_t)img * sP + (size_t)(2 * y) * sw + 8 * q;
    const uint2 a = *reinterpret_cast<const uint2*>(s0);
    const uint2 b = *reinterpret_cast<const uint2*>(s0 + sw);
    auto mean4 = [](uint32_t top, uint32_t bot, int k) {       // pixels 2k, 2k+1 of the two source rows
      const uint32_t t = (top >> (16 * k)) & 0xffffu, u = (bot >> (16 * k)) & 0xffffu;
      return ((t & 0xffu) + (t >> 8) + (u & 0xffu) + (u >> 8)) >> 2;
    };
    const uint32_t o = mean4(a.x, b.x, 0) | (mean4(a.x, b.x, 1) << 8) | (mean4(a.y, b.y, 0) << 16) | (mean4(a.y, b.y, 1) << 24);
    *reinterpret_cast<uint32_t*>(dst + (size_t)img * dP + (size_t)y * dw + 4 * q) = o;
  }
}

// src: n_img images of sw x sh (tightly packed, sw % 8 == 0); dst: n_img images of (sw/2) x (sh/2)
cudaError_t launch_downsample2x(const uint8_t* src, uint8_t* dst, int sw, int sh, int n_img, cudaStream_t stream) {
  const int dw = sw / 2, dh = sh / 2;
  if (dw <= 0 || dh <= 0 || n_img <= 0) return cudaSuccess;
  const long long total = (long long)n_img * dh * (dw / 4);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  downsample2x_kernel<<<blocks, 256, 0, stream>>>(src, dst, sw, dw, dh, n_img, (size_t)sw * sh, (size_t)dw * dh);
  return cudaGetLastError();
}

}  // namespace gpc

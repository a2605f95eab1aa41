// hash_tiles.cu -- kernel A2: fern hashing, one 128 x 32 pixel tile per CTA, operands staged by TMA.
//
// Replaces the reference's ndb::gpcFilter / gpcFilterTau (filter.hpp:547-606, :619-683) and the
// gather of Forest::evalFastMaskOnSubsetSSE (inference.hpp:266-292).  Nothing here is a translation
// of the SSE code.  The kernel is bound by the SM's ALU pipe (LOP3/SHF/PRMT: one warp instruction
// per two cycles per scheduler), not by HBM, so everything is arranged to minimise ALU-pipe work:
//   * the tile of the biased smoothed image (kernel A1), 13-pixel halo included, is fetched by ONE
//     TMA tensor copy (zero fill outside the image, no staging instructions); the CTA then keeps it
//     four times, copy k shifted left by k bytes (TMA itself needs 16-byte aligned source
//     coordinates, so copies 1..3 are funnel-shifted out of copy 0 in shared memory), which makes
//     the 4-pixel operand of every test ONE aligned LDS with a uniform offset;
//   * in the biased domain the reference's "signed-saturating b - tau, then unsigned compare"
//     (filter.hpp:647-652) is clamp(x - tau, 0, 255) -- two DPX VIADDMNMX on 16-bit lanes --
//     followed by a SIGNED byte compare, which costs the same carry trick as the unsigned one;
//   * result bits are accumulated with one LEA.HI each: acc += (r & 0x80808080) >> (7 - p) puts test p of pixel j
//     at bit 8j+p without carries (no wide multiply in the loop: IMAD.WIDE / IMAD.HI hold the dispatch port);
//   * the tests of one state byte form a single basic block (no per-test guards: the forest is
//     padded with never-true tests), so the compiler interleaves their dependency chains.
// Each pixel's state is written once to the hash image (bit 31 = candidate).
#ifndef __CUDACC_RTC__
#include <cuda.h>
#else
struct alignas(64) CUtensorMap_st { unsigned long long opaque[16]; };
typedef CUtensorMap_st CUtensorMap;
#endif

#include "gpc_device.cuh"

namespace gpc {

constexpr uint32_t kMsb = 0x80808080u;
constexpr uint32_t kLow7 = 0x7f7f7f7fu;

// One fern test on 4 horizontally adjacent pixels; operands are BIASED bytes (pixel ^ 0x80).
//   zero forest (filter.hpp:575):      a > b  unsigned           ==  xa > xb        signed
//   tau forest  (filter.hpp:647-652):  a > (uint8)sat_int8((int8)b - tau)  unsigned
//                                      ==  xa > clamp(xb - tau, 0, 255)    signed
// Signed byte compare: msb = (~a7 & c7) | (~(a7 ^ c7) & carry7), carry from the low 7 bits.
// The returned word is masked with `msk`: the msbs of the quad's CANDIDATE pixels only, so that a pixel that is no
// candidate accumulates nothing (its state word stays 0 without any per-pixel select afterwards).  A tau forest sends
// every test through the clamp (tau == 0 leaves x unchanged).
#ifdef GPC_JIT_HEADER
#include GPC_JIT_HEADER      // forest baked into the code: jit_imm_a(t), jit_imm_b(t), jit_mtau2(t), kJitTests
#endif

// kMode: 0 = a > b (biased bytes, signed compare), 1 = tau forest of the SSE build, 2 / 3 = the same two for the
// reference's SSE=OFF build (unbiased bytes; gpcFilterNaive / gpcFilterTauNaive, filter.hpp:245-293).
constexpr int kModeZero = 0, kModeTau = 1, kModeNaiveZero = 2, kModeNaiveTau = 3;

template <int kMode>
__device__ __forceinline__ uint32_t eval_test(const uint8_t* base, const ForestDev& forest, const int t, const uint32_t msk) {
#ifdef GPC_JIT_HEADER
  if (t >= kJitTests) return 0u;                                             // compile time after unrolling
  const uint32_t a = *reinterpret_cast<const uint32_t*>(base + jit_imm_a(t));
  uint32_t c = *reinterpret_cast<const uint32_t*>(base + jit_imm_b(t));
  if (kMode == kModeTau && jit_mtau2(t) != 0u) {
    const uint32_t mt = jit_mtau2(t);
#else
  const uint32_t a = *reinterpret_cast<const uint32_t*>(base + forest.imm_a[t]);
  uint32_t c = *reinterpret_cast<const uint32_t*>(base + forest.imm_b[t]);
  if (kMode == kModeNaiveTau) {
    // a > b - tau in plain int arithmetic, two pixels per word in 16-bit lanes: lane = b + (32768 - tau) - a stays
    // inside 16 bits, and its bit 15 is CLEAR exactly when a + tau > b
    const uint32_t k2 = forest.mtau2[t];
    const uint32_t lo = __byte_perm(c, 0u, 0x4140) + k2 - __byte_perm(a, 0u, 0x4140);
    const uint32_t hi = __byte_perm(c, 0u, 0x4342) + k2 - __byte_perm(a, 0u, 0x4342);
    return ~__byte_perm(lo, hi, 0x7531) & msk;
  }
  if (kMode == kModeNaiveZero) {                                             // unsigned a > c on unbiased bytes
    const uint32_t s = (a & kLow7) + (~c & kLow7);
    return ((a & ~c) | (~(a ^ c) & s)) & msk;
  }
  if (kMode == kModeTau) {
    const uint32_t mt = forest.mtau2[t];
#endif
    const uint32_t lo = __viaddmin_s16x2_relu(__byte_perm(c, 0u, 0x4140), mt, 0x00ff00ffu);
    const uint32_t hi = __viaddmin_s16x2_relu(__byte_perm(c, 0u, 0x4342), mt, 0x00ff00ffu);
    c = __byte_perm(lo, hi, 0x6420);                                         // clamp(x - tau, 0, 255)
  }
  const uint32_t s = (a & kLow7) + (~c & kLow7);                             // bit 7: low7(a) > low7(c)
  return ((~a & c) | (~(a ^ c) & s)) & msk;
}

// All tests of state byte G (filter.hpp:574-584: tests 0..8 -> byte 0 with test 8 OR-ed into bit 0
// under m8, 9..16 -> byte 1, 17..24 -> byte 2, 25..31 -> byte 3).  A test's result word carries its four flags in
// bit 7 of the bytes; bit p of the state byte is reached by a right shift by 7 - p, and "acc + (r >> k)" is ONE
// ALU-pipe instruction (LEA.HI: funnel shift + add).  Measured against the alternatives on the FMA pipe -- IMAD.WIDE
// (left shifts into a 64-bit accumulator) and IMAD.HI (multiply-high by 2^(25+p)) -- it wins although the ALU pipe is
// the busiest one: both wide multiplies hold the dispatch port for several cycles (0.68 / 0.63 / 0.59 ms, zero forest).
template <int kMode, int G>
__device__ __forceinline__ uint32_t eval_group(const uint8_t* base, const ForestDev& forest, uint32_t m8, uint32_t msk, uint32_t acc) {
  constexpr int t0 = (G == 0) ? 1 : 8 * G + 1;
  constexpr int t1 = (G == 0) ? 8 : (G == 3) ? kMaxTests : 8 * G + 9;       // exclusive
  if (G == 0) {
    const uint32_t r0 = eval_test<kMode>(base, forest, 0, msk), r8 = eval_test<kMode>(base, forest, 8, msk);
    acc += (r0 | (r8 & m8)) >> 7;
  }
#pragma unroll
  for (int t = t0; t < t1; t++) {
    const uint32_t r = eval_test<kMode>(base, forest, t, msk);
    const int p = ((t < 8) ? t : t - 1) & 7;                                // compile time after unrolling
    acc += r >> (7 - p);
  }
  return acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#ifndef GPC_MINB_A
#define GPC_MINB_A 1
#endif
template <int kMode>
__device__ __forceinline__ void hash_tiles_body(const CUtensorMap& tmap, const HashArgs& args, const ForestDev& forest) {
  extern __shared__ __align__(128) uint8_t smem[];                // 4 copies of [kSmRows][kPitch] biased bytes
  __shared__ __align__(8) unsigned long long mbar;

  const int W = args.W, H = args.H;
  const int img = args.img0 + blockIdx.z;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int tid = threadIdx.x;
  const size_t img_off = (size_t)img * W * H;

  // ---- candidate masks of this thread's quads: issued first, consumed after the tile has landed --------
  const int qx = tid % kQuadsX;
  const int gx = x0 + 4 * qx;
  constexpr int kRowStep = kThreadsA / kQuadsX;
  constexpr int kIters = kTileH / kRowStep;
  static_assert(kIters <= 8, "candidate nibbles of a thread are packed into one word");
  uint32_t craw[kIters];
  {
    // one 64-bit address per thread, the rows of its iterations are 32-bit offsets from it (no 64-bit multiply per row)
    const uint16_t* const cand0 = args.cand + ((size_t)img * H + (y0 + tid / kQuadsX)) * (size_t)(W / 16) + (gx >> 4);
    const uint32_t row_stride = (uint32_t)kRowStep * (uint32_t)(W / 16);
#pragma unroll
    for (int i = 0; i < kIters; i++) {
      const int gy = y0 + tid / kQuadsX + i * kRowStep;
      craw[i] = 0;
      if (gy < H && gx < W) craw[i] = __ldg(cand0 + (uint32_t)i * row_stride);
    }
  }

  // ---- TMA: copy 0 = image columns x0 - 16 .., rows y0 - 13 ..; zero fill outside the image ---------
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(kSmRows * kPitch) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem)), "l"(&tmap), "r"(x0 - 16), "r"(y0 - kRadius), "r"(img), "r"(smem_u32(&mbar))
        : "memory");
  }
  __syncthreads();                                                 // the initialised barrier is visible to all waiters

  {                                                                // wait for the tile
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
  }

  uint32_t cms = 0;                                                // nibble i = candidate bits of iteration i
#pragma unroll
  for (int i = 0; i < kIters; i++) cms |= ((craw[i] >> ((qx & 3) * 4)) & 15u) << (4 * i);

  // ---- copies 1..3 = copy 0 shifted left by 1..3 bytes (two words per step) ----------------------------
  // The rows of a copy lie back to back, so the tile is shifted as ONE linear array: the last word of a row then takes
  // its top bytes from the next row instead of zeros, which no test can see -- an operand's four bytes end at column
  // 16 + 127 + 13 + 3 = kPitch - 1 of its own row at the latest.
  {
    static_assert(kPitchW % 2 == 0 && kCopyBytes % 8 == 0, "64-bit accesses");
    static_assert(16 + (kTileW - 1) + kRadius + 3 < kPitch, "operands never reach past their row");
    const uint32_t* x32 = reinterpret_cast<const uint32_t*>(smem);
    constexpr int kPairs = kSmRows * (kPitchW / 2);
    for (int i = tid; i < kPairs; i += kThreadsA) {
      const uint2 w = *reinterpret_cast<const uint2*>(x32 + 2 * i);
      const uint32_t nx = (i + 1 < kPairs) ? x32[2 * i + 2] : 0u;
      uint8_t* dst = smem + 8 * i;
      *reinterpret_cast<uint2*>(dst + 1 * kCopyBytes) = make_uint2(__funnelshift_r(w.x, w.y, 8), __funnelshift_r(w.y, nx, 8));
      *reinterpret_cast<uint2*>(dst + 2 * kCopyBytes) = make_uint2(__funnelshift_r(w.x, w.y, 16), __funnelshift_r(w.y, nx, 16));
      *reinterpret_cast<uint2*>(dst + 3 * kCopyBytes) = make_uint2(__funnelshift_r(w.x, w.y, 24), __funnelshift_r(w.y, nx, 24));
    }
  }
  __syncthreads();

  // ---- fern tests, 4 pixels per step ---------------------------------------------------------------------
  uint32_t* __restrict__ hash_out = args.hash + img_off + (size_t)(y0 + tid / kQuadsX) * W + gx;   // this thread's quad, iteration 0
  const uint32_t m8 = (gx & 4) ? kMsb : 0x80808000u;              // test #8: byte lanes x%8==0 dropped (filter.hpp:582); naive mode: slot 8 is a dummy
#ifdef GPC_JIT_HEADER
  constexpr int T = kJitTests;                                    // all groups in ONE basic block: their chains interleave
#else
  const int T = forest.n_tests;
#endif
  const int n_groups = (T <= 9) ? 1 : (T <= 17) ? 2 : (T <= 25) ? 3 : 4;
#pragma unroll 1
  for (int i = 0; i < kIters; i++) {
    const int ry = tid / kQuadsX + i * kRowStep;
    const int gy = y0 + ry;
    const uint32_t cm = (cms >> (4 * i)) & 15u;
    // msb of byte j set iff pixel j is a candidate (the 16 partial products of the multiplication hit 16 distinct bits)
    const uint32_t msk = (cm * 0x10204080u) & kMsb;
    uint32_t st[4];
    if (cm != 0u && gy >= kRadius && gy < args.hash_y_end) {       // hashed rows (filter.hpp:601-604)
      const uint8_t* base = smem + (ry + kRadius) * kPitch + 16 + 4 * qx;
      // group 3 starts from the candidate flag: bit 7 of its byte = bit 31 of the state word
      uint32_t w[4] = {0u, 0u, 0u, msk};
      w[0] = eval_group<kMode, 0>(base, forest, m8, msk, w[0]);
      if (n_groups > 1) w[1] = eval_group<kMode, 1>(base, forest, m8, msk, w[1]);      // uniform branches
      if (n_groups > 2) w[2] = eval_group<kMode, 2>(base, forest, m8, msk, w[2]);
      if (n_groups > 3) w[3] = eval_group<kMode, 3>(base, forest, m8, msk, w[3]);
      // byte j of w[g] = state byte g of pixel j; 4x4 byte transpose -> one state per pixel
      uint32_t lo01 = __byte_perm(w[0], w[1], 0x5140), hi01 = __byte_perm(w[0], w[1], 0x7362);
      uint32_t lo23 = __byte_perm(w[2], w[3], 0x5140), hi23 = __byte_perm(w[2], w[3], 0x7362);
      st[0] = __byte_perm(lo01, lo23, 0x5410);
      st[1] = __byte_perm(lo01, lo23, 0x7632);
      st[2] = __byte_perm(hi01, hi23, 0x5410);
      st[3] = __byte_perm(hi01, hi23, 0x7632);
    } else {                                                       // candidates of rows that are not hashed: state 0
#pragma unroll
      for (int j = 0; j < 4; j++) st[j] = ((cm >> j) & 1u) ? kCandFlag : 0u;
    }
    const uint4 o = make_uint4(st[0], st[1], st[2], st[3]);
    if (gx < W && gy < H) *reinterpret_cast<uint4*>(hash_out) = o;
    hash_out += (uint32_t)kRowStep * (uint32_t)W;
  }
}

template <int kMode>
__global__ void __launch_bounds__(kThreadsA, GPC_MINB_A)
hash_tiles_kernel(const __grid_constant__ CUtensorMap tmap, const HashArgs args, const ForestDev forest) {
  hash_tiles_body<kMode>(tmap, args, forest);
}

#ifdef GPC_JIT_HEADER
// Entry point of the forest-specialised build (NVRTC, jit.cu): unmangled name, forest baked in.
extern "C" __global__ void __launch_bounds__(kThreadsA, GPC_MINB_A)
gpc_hash_tiles_jit(const __grid_constant__ CUtensorMap tmap, const HashArgs args, const ForestDev forest) {
  hash_tiles_body<kJitTau ? kModeTau : kModeZero>(tmap, args, forest);
}
#endif

#ifndef __CUDACC_RTC__
size_t hash_smem_bytes() { return (size_t)4 * kCopyBytes; }

cudaError_t configure_hash_tiles() {   // per device: opt in to > 48 KB dynamic shared memory
  cudaError_t e = cudaFuncSetAttribute(hash_tiles_kernel<kModeZero>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hash_smem_bytes());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(hash_tiles_kernel<kModeTau>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hash_smem_bytes());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(hash_tiles_kernel<kModeNaiveZero>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hash_smem_bytes());
  if (e == cudaSuccess) e = cudaFuncSetAttribute(hash_tiles_kernel<kModeNaiveTau>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hash_smem_bytes());
  return e;
}

// Tensor map over u8 images [n_img][H][W] with a box of box_w x box_h x 1 (zero fill outside the tensor).
int make_u8_tensor_map(void* out_map, const uint8_t* base, int W, int H, int n_img, int box_w, int box_h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
  const cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * (cuuint64_t)H};      // bytes, dims 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// Tensor map over the biased smoothed images; box = one operand tile of kernel A2.
int make_smooth_tensor_map(void* out_map, const uint8_t* base, int W, int H, int n_img) {
  return make_u8_tensor_map(out_map, base, W, H, n_img, kPitch, kSmRows);
}

cudaError_t launch_hash_tiles(const void* tensor_map, const HashArgs& args, const ForestDev& forest, int n_img, cudaStream_t stream) {
  dim3 grid((args.W + kTileW - 1) / kTileW, (args.H + kTileH - 1) / kTileH, n_img);
  const CUtensorMap& tmap = *reinterpret_cast<const CUtensorMap*>(tensor_map);
  if (forest.naive) {
    if (forest.type != 0) hash_tiles_kernel<kModeNaiveTau><<<grid, kThreadsA, hash_smem_bytes(), stream>>>(tmap, args, forest);
    else hash_tiles_kernel<kModeNaiveZero><<<grid, kThreadsA, hash_smem_bytes(), stream>>>(tmap, args, forest);
  } else if (forest.type != 0) hash_tiles_kernel<kModeTau><<<grid, kThreadsA, hash_smem_bytes(), stream>>>(tmap, args, forest);
  else hash_tiles_kernel<kModeZero><<<grid, kThreadsA, hash_smem_bytes(), stream>>>(tmap, args, forest);
  return cudaGetLastError();
}
#endif   // !__CUDACC_RTC__

}  // namespace gpc

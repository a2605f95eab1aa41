// preprocess_hash.cu -- kernel A: box blur + Sobel candidates + fern hashing, one tile per CTA.
//
// Replaces, fused and for a whole batch of images, the reference's
//   ndb::box + Buffer::clearBoundary   (filter.hpp:293-392, buffer.hpp:630-654)
//   ndb::sobel                         (filter.hpp:404-519, incl. the lane duplication at :504-507)
//   ndb::arr2ind + border lambda       (filter.hpp:60-87, inference.hpp:318-330)
//   ndb::gpcFilter / gpcFilterTau      (filter.hpp:547-606, :619-683)
// Nothing here is a translation of the SSE code.  A CTA stages a (32+28) x (256+32) raw tile in
// shared memory, derives the smoothed tile with dp4a row sums, evaluates the Sobel predicate per
// 16-pixel segment, and then evaluates all fern tests 4 pixels at a time with byte-SIMD integer
// arithmetic.  The kernel is bound by the SM's ALU pipe (LOP3/SHF/PRMT, one warp instruction per
// two cycles per scheduler), not by HBM, so the test loop is written to minimise ALU-pipe work:
//   * the smoothed tile is stored BIASED (s ^ 0x80) and four times, copy k shifted left by k
//     bytes: every 4-pixel operand is one aligned LDS with a uniform offset (no funnel shifts);
//   * in the biased domain the reference's "signed-saturating b - tau, then unsigned compare"
//     (filter.hpp:647-652) is clamp(x - tau, 0, 255) -- two DPX VIADDMNMX on 16-bit lanes --
//     followed by a SIGNED byte compare, which costs the same carry trick as the unsigned one;
//   * result bits are accumulated with IMAD.WIDE on the otherwise idle FMA pipe:
//     acc64 += (r & 0x80808080) * 2^p puts test p of pixel j at bit 8j+7+p without carries.
// Each pixel's state is written once to the hash image (bit 31 = candidate).
#include "gpc_device.cuh"

namespace gpc {

__device__ __forceinline__ uint32_t third(uint32_t s) { return __umulhi(s, 21846u << 16); }   // (s*21846)>>16
__device__ __forceinline__ uint32_t ninth(uint32_t s) { return __umulhi(s, 7282u << 16); }    // (s*7282)>>16

constexpr uint32_t kMsb = 0x80808080u;
constexpr uint32_t kLow7 = 0x7f7f7f7fu;

// Horizontal floor-thirds of 4 consecutive pixels: h[k] = (p[x+k-1] + p[x+k] + p[x+k+1]) / 3.
__device__ __forceinline__ void hthirds(uint32_t wm1, uint32_t w, uint32_t wp1, uint32_t h[4]) {
  h[0] = third(__dp4a(__funnelshift_r(wm1, w, 24), 0x00010101u, 0u));
  h[1] = third(__dp4a(w, 0x00010101u, 0u));
  h[2] = third(__dp4a(w, 0x01010100u, 0u));
  h[3] = third(__dp4a(__funnelshift_r(w, wp1, 16), 0x00010101u, 0u));
}

// One fern test on 4 horizontally adjacent pixels; operands are BIASED bytes (pixel ^ 0x80).
// Returns a word whose byte msbs are the test results (other bits are garbage).
//   zero forest (filter.hpp:575):      a > b  unsigned           ==  xa > xb        signed
//   tau forest  (filter.hpp:647-652):  a > (uint8)sat_int8((int8)b - tau)  unsigned
//                                      ==  xa > clamp(xb - tau, 0, 255)    signed
// Signed byte compare: msb = (~a7 & c7) | (~(a7 ^ c7) & carry7), carry from the low 7 bits.
// The returned word is masked to the msbs.  No branches: a tau forest sends every test through
// the clamp (tau == 0 leaves x unchanged), so that the tests of one state byte form a single
// basic block the compiler can interleave.
template <bool kTau>
__device__ __forceinline__ uint32_t eval_test(const uint8_t* base, const ForestDev& forest, const int t) {
  const uint32_t a = *reinterpret_cast<const uint32_t*>(base + forest.imm_a[t]);
  uint32_t c = *reinterpret_cast<const uint32_t*>(base + forest.imm_b[t]);
  if (kTau) {
    const uint32_t mt = forest.mtau2[t];
    const uint32_t lo = __viaddmin_s16x2_relu(__byte_perm(c, 0u, 0x4140), mt, 0x00ff00ffu);
    const uint32_t hi = __viaddmin_s16x2_relu(__byte_perm(c, 0u, 0x4342), mt, 0x00ff00ffu);
    c = __byte_perm(lo, hi, 0x6420);                                         // clamp(x - tau, 0, 255)
  }
  const uint32_t s = (a & kLow7) + (~c & kLow7);                             // bit 7: low7(a) > low7(c)
  return ((~a & c) | (~(a ^ c) & s)) & kMsb;
}

// All tests of state byte G (filter.hpp:574-584: tests 0..8 -> byte 0 with test 8 OR-ed into bit 0
// under m8, 9..16 -> byte 1, 17..24 -> byte 2, 25..31 -> byte 3).  The forest is padded with
// never-true dummy tests up to the end of its last group (bake_forest), so there is no per-test
// guard.  acc += msb word * 2^p lands pixel j's bit at 8j + 7 + p without carries.
template <bool kTau, int G>
__device__ __forceinline__ unsigned long long eval_group(const uint8_t* base, const ForestDev& forest, uint32_t m8) {
  constexpr int t0 = (G == 0) ? 0 : 8 * G + 1;
  constexpr int t1 = (G == 3) ? kMaxTests : 8 * G + 9;           // exclusive
  unsigned long long acc = 0ull;
  if (G == 0) {
    const uint32_t r0 = eval_test<kTau>(base, forest, 0), r8 = eval_test<kTau>(base, forest, 8);
    acc = (unsigned long long)(r0 | (r8 & m8));
  }
#pragma unroll
  for (int t = (G == 0) ? 1 : t0; t < ((G == 0) ? 8 : t1); t++)
    acc += (unsigned long long)eval_test<kTau>(base, forest, t) * (unsigned long long)forest.pmul[t];
  return acc;
}

// kMode 0: product path.  1: also writes smooth / grad (stage-parity seam gpc_preprocess).
// 2: evalFastMaskOnSubsetSSE seam (gpc_hash_smooth): args.raw is an already smoothed image and
//    args.flags a u8 image whose non-zero bytes mark the pixels to hash; phases 1-2 are skipped.
//
// Shared memory: copies 0..3 of the biased smoothed tile, [kSmRows][kPitch] bytes each; the raw
// tile [kRawRows][kPitch] is staged over copies 1..3 and is dead by the time they are built;
// then the candidate masks [kTileH][kTileW/16] u16.
template <int kMode, bool kTau>
__global__ void __launch_bounds__(kThreadsA)
preprocess_hash_kernel(const PreprocessArgs args, const ForestDev forest) {
  constexpr bool kDebugOut = (kMode == 1);
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* x32 = reinterpret_cast<uint32_t*>(smem);                            // copy 0: [kSmRows][kPitchW]
  uint32_t* raw32 = reinterpret_cast<uint32_t*>(smem + kCopyBytes);             // [kRawRows][kPitchW], aliases copies 1..3
  uint16_t* cand = reinterpret_cast<uint16_t*>(smem + 4 * kCopyBytes);          // [kTileH][kTileW/16]
  static_assert(kRawRows * kPitch <= 3 * kCopyBytes, "raw staging must fit over copies 1..3");

  const int W = args.W, H = args.H;
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int tid = threadIdx.x;
  const size_t img_off = (size_t)img * W * H;
  const uint8_t* __restrict__ raw = args.raw + img_off;

  // ---- phase 0: stage the raw tile (zero outside the image) --------------------------------
  if (kMode == 2) {
    constexpr int kChunks = kPitch / 16;
    uint4* dst = reinterpret_cast<uint4*>(x32);
    for (int c = tid; c < kSmRows * kChunks; c += kThreadsA) {
      int r = c / kChunks, k = c - r * kChunks;
      int gy = y0 - kRadius + r, gx = x0 - 16 + 16 * k;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(reinterpret_cast<const uint4*>(raw + (size_t)gy * W + gx));
      v.x ^= kMsb; v.y ^= kMsb; v.z ^= kMsb; v.w ^= kMsb;
      dst[c] = v;
    }
    const uint8_t* __restrict__ flags = args.flags + img_off;
    for (int sr = tid; sr < kTileH * (kTileW / 16); sr += kThreadsA) {
      const int ry = sr / (kTileW / 16), sg = sr - ry * (kTileW / 16);
      const int gy = y0 + ry, gxs = x0 + 16 * sg;
      uint32_t m = 0;
      if (gy >= kRadius && gy < H - kRadius && gxs < W) {
        const uint4 f = __ldg(reinterpret_cast<const uint4*>(flags + (size_t)gy * W + gxs));
        const uint32_t wv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int b = 0; b < 4; b++)
            if ((wv[k] >> (8 * b)) & 0xffu) m |= 1u << (4 * k + b);
        const int lowcut = kRadius - gxs;
        if (lowcut >= 16) m = 0; else if (lowcut > 0) m &= ~((1u << lowcut) - 1u);
        const int keep = W - kRadius - gxs;
        if (keep <= 0) m = 0; else if (keep < 16) m &= (1u << keep) - 1u;
      }
      cand[sr] = (uint16_t)m;
    }
  } else {
    constexpr int kChunks = kPitch / 16;
    uint4* dst = reinterpret_cast<uint4*>(raw32);
    for (int c = tid; c < kRawRows * kChunks; c += kThreadsA) {
      int r = c / kChunks, k = c - r * kChunks;
      int gy = y0 - (kRadius + 1) + r, gx = x0 - 16 + 16 * k;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(reinterpret_cast<const uint4*>(raw + (size_t)gy * W + gx));
      dst[c] = v;
    }
  }
  if (kMode != 2) __syncthreads();

  // ---- phase 1: smoothed tile (biased) -> copy 0 ---------------------------------------------
  // thread = (quad column, one of 3 row segments); walks down its rows with a 3-row window of
  // horizontal thirds held in registers.
  if (kMode != 2) {
    constexpr int kSegs = kThreadsA / kPitchW;
    constexpr int kSegRows = (kSmRows + kSegs - 1) / kSegs;
    const int q = tid % kPitchW, seg = tid / kPitchW;
    if (seg < kSegs) {
      const int j0 = seg * kSegRows, j1 = min(kSmRows, j0 + kSegRows);
      const int last_written = (H & 1) ? H - 3 : H - 4;    // box writes rows 1..last (filter.hpp:307,388)
      const int gxq = x0 - 16 + 4 * q;
      uint32_t colmask = 0xffffffffu;                      // clearBoundary: columns 0,1 and W-1
      if (gxq == 0) colmask = 0xffff0000u;
      if (gxq == W - 4) colmask &= 0x00ffffffu;
      if (gxq < 0 || gxq >= W) colmask = 0u;
      uint32_t ha[4], hb[4], hc[4];
      auto load_h = [&](int r, uint32_t h[4]) {
        const uint32_t* row = raw32 + r * kPitchW;
        uint32_t wm1 = q > 0 ? row[q - 1] : 0u, w = row[q], wp1 = q + 1 < kPitchW ? row[q + 1] : 0u;
        hthirds(wm1, w, wp1, h);
      };
      load_h(j0, ha);       // raw-tile row j = image row y0-14+j; smooth row j needs raw rows j..j+2
      load_h(j0 + 1, hb);
#pragma unroll 3
      for (int j = j0; j < j1; j++) {
        load_h(j + 2, hc);
        uint32_t v = third(ha[0] + hb[0] + hc[0]) | (third(ha[1] + hb[1] + hc[1]) << 8) |
                     (third(ha[2] + hb[2] + hc[2]) << 16) | (third(ha[3] + hb[3] + hc[3]) << 24);
        const int gy = y0 - kRadius + j;
        if (gy < 1 || gy > last_written) v = 0u; else v &= colmask;
        x32[j * kPitchW + q] = v ^ kMsb;
        if (kDebugOut && args.smooth_out && j >= kRadius && j < kRadius + kTileH && q >= 4 && q < 4 + kTileW / 4 &&
            gy < H && gxq < W)
          *reinterpret_cast<uint32_t*>(args.smooth_out + img_off + (size_t)gy * W + gxq) = v;
#pragma unroll
        for (int k = 0; k < 4; k++) { ha[k] = hb[k]; hb[k] = hc[k]; }
      }
    }
  }

  // ---- phase 2: Sobel predicate per 16-pixel segment -> candidate bit masks -----------------
  if (kMode != 2) {
    const uint8_t* raw8 = reinterpret_cast<const uint8_t*>(raw32);
    for (int sr = tid; sr < kTileH * (kTileW / 16); sr += kThreadsA) {
      const int ry = sr / (kTileW / 16), sg = sr - ry * (kTileW / 16);
      const int gy = y0 + ry, gxs = x0 + 16 * sg;
      uint32_t m = 0;
      if (gy >= 1 && gy < H - 3 && gxs < W) {            // rows the reference writes (filter.hpp:517)
        const uint8_t* r0 = raw8 + (ry + kRadius) * kPitch + 16 + 16 * sg;   // image row gy-1, col gxs
        const uint8_t* r1 = r0 + kPitch;
        const uint8_t* r2 = r1 + kPitch;
#pragma unroll
        for (int g = 0; g < 8; g++) {
          const int c = (g < 4) ? g : g + 4;              // true columns s..s+3 and s+8..s+11 survive :504-507
          int p00 = r0[c - 1], p01 = r0[c], p02 = r0[c + 1];
          int p10 = r1[c - 1], p12 = r1[c + 1];
          int p20 = r2[c - 1], p21 = r2[c], p22 = r2[c + 1];
          int a = (int)ninth(p00 + p20 + 2 * p10), b = (int)ninth(p02 + p22 + 2 * p12);
          int cc = (int)ninth(p00 + p02 + 2 * p01), d = (int)ninth(p20 + p22 + 2 * p21);
          int sum = (a - b) * (a - b) + (cc - d) * (cc - d);   // <= 25538, no int16 wrap / saturation
          if (sum > args.thr2) m |= 3u << (2 * g);            // lane duplication: outputs 2g, 2g+1
        }
      }
      if (kDebugOut && args.grad_out && gy < H && gxs < W) {
        uint32_t wv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint32_t nib = (m >> (4 * k)) & 15u;
          wv[k] = ((nib & 1u) * 0xffu) | ((nib & 2u) * (0xff00u >> 1)) | ((nib & 4u) * (0xff0000u >> 2)) |
                  ((nib & 8u) * (0xff000000u >> 3));
        }
        *reinterpret_cast<uint4*>(args.grad_out + img_off + (size_t)gy * W + gxs) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      }
      // candidate border (inference.hpp:322): 13 <= x < W-13, 13 <= y < H-13
      if (gy < kRadius || gy >= H - kRadius) m = 0;
      else {
        const int lowcut = kRadius - gxs;                 // columns below 13
        if (lowcut >= 16) m = 0; else if (lowcut > 0) m &= ~((1u << lowcut) - 1u);
        const int keep = W - kRadius - gxs;               // columns below W-13
        if (keep <= 0) m = 0; else if (keep < 16) m &= (1u << keep) - 1u;
      }
      cand[sr] = (uint16_t)m;
    }
  }
  __syncthreads();                                         // copy 0 and cand complete, raw tile dead

  // ---- phase 2b: copies 1..3 = copy 0 shifted left by 1..3 bytes -----------------------------
  for (int i = tid; i < kSmRows * kPitchW; i += kThreadsA) {
    const int q = i % kPitchW;
    const uint32_t lo = x32[i], hi = (q + 1 < kPitchW) ? x32[i + 1] : 0u;
    x32[i + 1 * (kCopyBytes / 4)] = __funnelshift_r(lo, hi, 8);
    x32[i + 2 * (kCopyBytes / 4)] = __funnelshift_r(lo, hi, 16);
    x32[i + 3 * (kCopyBytes / 4)] = __funnelshift_r(lo, hi, 24);
  }
  __syncthreads();

  // ---- phase 3: fern tests, 4 pixels per step ------------------------------------------------
  {
    const int qx = tid % kQuadsX;
    const int gx = x0 + 4 * qx;
    uint32_t* __restrict__ hash = args.hash + img_off;
    const uint32_t m8 = (gx & 4) ? kMsb : 0x80808000u;          // test #8: byte lanes x%8==0 dropped (filter.hpp:582)
    const int T = forest.n_tests;
    const int n_groups = (T <= 9) ? 1 : (T <= 17) ? 2 : (T <= 25) ? 3 : 4;
#pragma unroll 1
    for (int ry = tid / kQuadsX; ry < kTileH; ry += kThreadsA / kQuadsX) {
      const int gy = y0 + ry;
      const bool inside = gx < W && gy < H;                      // cand is 0 outside the image
      const uint32_t cm = (cand[ry * (kTileW / 16) + (qx >> 2)] >> ((qx & 3) * 4)) & 15u;
      uint32_t st[4] = {0u, 0u, 0u, 0u};
      if (cm != 0u && gy >= kRadius && gy < H - 15) {            // hashed rows (filter.hpp:601-604)
        const uint8_t* base = smem + (ry + kRadius) * kPitch + 16 + 4 * qx;
        unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};
        acc[0] = eval_group<kTau, 0>(base, forest, m8);
        if (n_groups > 1) acc[1] = eval_group<kTau, 1>(base, forest, m8);      // uniform branches
        if (n_groups > 2) acc[2] = eval_group<kTau, 2>(base, forest, m8);
        if (n_groups > 3) acc[3] = eval_group<kTau, 3>(base, forest, m8);
        // byte j of (acc[g] >> 7) = state byte g of pixel j; 4x4 byte transpose -> one state per pixel
        uint32_t w[4];
#pragma unroll
        for (int g = 0; g < 4; g++) w[g] = (uint32_t)(acc[g] >> 7);
        uint32_t lo01 = __byte_perm(w[0], w[1], 0x5140), hi01 = __byte_perm(w[0], w[1], 0x7362);
        uint32_t lo23 = __byte_perm(w[2], w[3], 0x5140), hi23 = __byte_perm(w[2], w[3], 0x7362);
        st[0] = __byte_perm(lo01, lo23, 0x5410);
        st[1] = __byte_perm(lo01, lo23, 0x7632);
        st[2] = __byte_perm(hi01, hi23, 0x5410);
        st[3] = __byte_perm(hi01, hi23, 0x7632);
      }
      uint4 o;
      o.x = (cm & 1u) ? (st[0] | kCandFlag) : 0u;
      o.y = (cm & 2u) ? (st[1] | kCandFlag) : 0u;
      o.z = (cm & 4u) ? (st[2] | kCandFlag) : 0u;
      o.w = (cm & 8u) ? (st[3] | kCandFlag) : 0u;
      if (inside) *reinterpret_cast<uint4*>(hash + (size_t)gy * W + gx) = o;
      // per-row candidate counts (a warp covers 32 consecutive quads of one row)
      int cnt = __reduce_add_sync(0xffffffffu, __popc(cm));
      if ((tid & 31) == 0 && cnt > 0) {
        atomicAdd(args.rowcnt + (size_t)img * H + gy, cnt);
        atomicMax(args.lastrow + img, gy);
      }
    }
  }
}

size_t preprocess_smem_bytes() {
  return (size_t)4 * kCopyBytes + (size_t)kTileH * (kTileW / 16) * sizeof(uint16_t);
}

template <int kMode, bool kTau>
static cudaError_t configure_one(int smem) {
  return cudaFuncSetAttribute(preprocess_hash_kernel<kMode, kTau>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

cudaError_t configure_preprocess_hash() {   // per device: opt in to > 48 KB dynamic shared memory
  int smem = (int)preprocess_smem_bytes();
  cudaError_t e = configure_one<0, false>(smem);
  if (e == cudaSuccess) e = configure_one<0, true>(smem);
  if (e == cudaSuccess) e = configure_one<1, false>(smem);
  if (e == cudaSuccess) e = configure_one<1, true>(smem);
  if (e == cudaSuccess) e = configure_one<2, false>(smem);
  if (e == cudaSuccess) e = configure_one<2, true>(smem);
  return e;
}

template <int kMode>
static void launch_mode(const PreprocessArgs& args, const ForestDev& forest, dim3 grid, size_t smem, cudaStream_t stream) {
  if (forest.type != 0)
    preprocess_hash_kernel<kMode, true><<<grid, kThreadsA, smem, stream>>>(args, forest);
  else
    preprocess_hash_kernel<kMode, false><<<grid, kThreadsA, smem, stream>>>(args, forest);
}

cudaError_t launch_preprocess_hash(const PreprocessArgs& args, const ForestDev& forest, int n_img,
                                   int mode, cudaStream_t stream) {
  size_t smem = preprocess_smem_bytes();
  dim3 grid((args.W + kTileW - 1) / kTileW, (args.H + kTileH - 1) / kTileH, n_img);
  if (mode == 2) launch_mode<2>(args, forest, grid, smem, stream);
  else if (mode == 1) launch_mode<1>(args, forest, grid, smem, stream);
  else launch_mode<0>(args, forest, grid, smem, stream);
  return cudaGetLastError();
}

}  // namespace gpc

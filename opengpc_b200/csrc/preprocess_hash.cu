// preprocess_hash.cu -- kernel A: box blur + Sobel candidates + fern hashing, one tile per CTA.
//
// Replaces, fused and for a whole batch of images, the reference's
//   ndb::box + Buffer::clearBoundary   (filter.hpp:293-392, buffer.hpp:630-654)
//   ndb::sobel                         (filter.hpp:404-519, incl. the lane duplication at :504-507)
//   ndb::arr2ind + border lambda       (filter.hpp:60-87, inference.hpp:318-330)
//   ndb::gpcFilter / gpcFilterTau      (filter.hpp:547-606, :619-683)
// Nothing here is a translation of the SSE code: a CTA stages a (32+28) x (256+32) raw tile in
// shared memory, derives the smoothed tile with dp4a row sums, evaluates the Sobel predicate
// per 16-pixel segment, and then evaluates all fern tests 4 pixels at a time with byte-SIMD
// integer arithmetic (funnel-shifted unaligned loads from the smoothed tile).  Each pixel's
// state is written once to the hash image (bit 31 = candidate).
#include "gpc_device.cuh"

namespace gpc {

__device__ __forceinline__ uint32_t third(uint32_t s) { return __umulhi(s, 21846u << 16); }   // (s*21846)>>16
__device__ __forceinline__ uint32_t ninth(uint32_t s) { return __umulhi(s, 7282u << 16); }    // (s*7282)>>16

// msb of each byte = (a > b) unsigned; other bits are garbage.
__device__ __forceinline__ uint32_t gtu4_msb(uint32_t a, uint32_t b) {
  uint32_t t = (a & 0x7f7f7f7fu) + (~b & 0x7f7f7f7fu);
  return (a & ~b) | (~(a ^ b) & t);
}

// Horizontal floor-thirds of 4 consecutive pixels: h[k] = (p[x+k-1] + p[x+k] + p[x+k+1]) / 3.
__device__ __forceinline__ void hthirds(uint32_t wm1, uint32_t w, uint32_t wp1, uint32_t h[4]) {
  h[0] = third(__dp4a(__funnelshift_r(wm1, w, 24), 0x00010101u, 0u));
  h[1] = third(__dp4a(w, 0x00010101u, 0u));
  h[2] = third(__dp4a(w, 0x01010100u, 0u));
  h[3] = third(__dp4a(__funnelshift_r(w, wp1, 16), 0x00010101u, 0u));
}

// One fern test on 4 horizontally adjacent pixels: msb of byte j = test result of pixel j.
//   zero forest (filter.hpp:575):  a > b                      (unsigned bytes)
//   tau forest  (filter.hpp:647-652): a > sat_int8(b - tau)   (signed saturating subtract of the byte
//   reinterpreted as int8, then an unsigned compare).  In the biased domain x = b ^ 0x80 the
//   saturating subtract is clamp(x - tau, 0, 255), done on two 16-bit lanes per register with the
//   native VIADDMNMX (DPX) instruction; the result is compared without un-biasing it.
__device__ __forceinline__ uint32_t eval_test(const uint8_t* base, const ForestDev& forest, const int t) {
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(base + forest.off_a[t]);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(base + forest.off_b[t]);
  uint32_t a = pa[0], b = pb[0];
  if (forest.sh_a[t]) a = __funnelshift_r(a, pa[1], forest.sh_a[t]);       // uniform branches
  if (forest.sh_b[t]) b = __funnelshift_r(b, pb[1], forest.sh_b[t]);
  const uint32_t mt = forest.mtau2[t];
  if (mt == 0u) return gtu4_msb(a, b);
  const uint32_t x = b ^ 0x80808080u;
  const uint32_t lo = __viaddmin_s16x2_relu(__byte_perm(x, 0u, 0x4140), mt, 0x00ff00ffu);
  const uint32_t hi = __viaddmin_s16x2_relu(__byte_perm(x, 0u, 0x4342), mt, 0x00ff00ffu);
  const uint32_t c = __byte_perm(lo, hi, 0x6420);                          // clamp(x - tau, 0, 255); b' = c ^ 0x80
  const uint32_t s = (a & 0x7f7f7f7fu) + (~c & 0x7f7f7f7fu);               // low-7-bit compare (bit 7 unaffected by ^0x80)
  return (a & c) | ((a ^ c) & s);                                          // msb: a7 > b'7, or equal and low7(a) > low7(b')
}

// kMode 0: product path.  1: also writes smooth / grad (stage-parity seam gpc_preprocess).
// 2: evalFastMaskOnSubsetSSE seam (gpc_hash_smooth): args.raw is an already smoothed image and
//    args.flags a u8 image whose non-zero bytes mark the pixels to hash; phases 1-2 are skipped.
template <int kMode>
__global__ void __launch_bounds__(kThreadsA)
preprocess_hash_kernel(const PreprocessArgs args, const ForestDev forest) {
  constexpr bool kDebugOut = (kMode == 1);
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* raw32 = reinterpret_cast<uint32_t*>(smem);                          // [kRawRows][kPitchW]
  uint32_t* sm32 = raw32 + kRawRows * kPitchW;                                  // [kSmRows][kPitchW]
  uint16_t* cand = reinterpret_cast<uint16_t*>(sm32 + kSmRows * kPitchW);       // [kTileH][kTileW/16]

  const int W = args.W, H = args.H;
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int tid = threadIdx.x;
  const size_t img_off = (size_t)img * W * H;
  const uint8_t* __restrict__ raw = args.raw + img_off;

  // ---- phase 0: stage the raw tile (zero outside the image) --------------------------------
  if (kMode == 2) {
    constexpr int kChunks = kPitch / 16;
    uint4* dst = reinterpret_cast<uint4*>(sm32);
    for (int c = tid; c < kSmRows * kChunks; c += kThreadsA) {
      int r = c / kChunks, k = c - r * kChunks;
      int gy = y0 - kRadius + r, gx = x0 - 16 + 16 * k;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(reinterpret_cast<const uint4*>(raw + (size_t)gy * W + gx));
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
    uint4* dst = reinterpret_cast<uint4*>(smem);
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

  // ---- phase 1: smoothed tile -----------------------------------------------------------------
  // thread = (quad column, one of 3 row segments); walks down its rows with a 3-row window of
  // horizontal thirds held in registers.
  if (kMode != 2) {
    constexpr int kSegs = 3;
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
        sm32[j * kPitchW + q] = v;
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
    const uint8_t* raw8 = smem;
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
  __syncthreads();

  // ---- phase 3: fern tests, 4 pixels per step ------------------------------------------------
  {
    const int qx = tid & 63;
    const int gx = x0 + 4 * qx;
    uint32_t* __restrict__ hash = args.hash + img_off;
    const uint32_t m8 = (gx & 4) ? 0x01010101u : 0x01010100u;   // test #8: byte lanes x%8==0 dropped (filter.hpp:582)
    const int T = forest.n_tests;
#pragma unroll 1
    for (int ry = tid >> 6; ry < kTileH; ry += kThreadsA / 64) {
      const int gy = y0 + ry;
      const bool inside = gx < W && gy < H;                      // cand is 0 outside the image
      const uint32_t cm = (cand[ry * (kTileW / 16) + (qx >> 2)] >> ((qx & 3) * 4)) & 15u;
      uint32_t st[4] = {0u, 0u, 0u, 0u};
      if (cm != 0u && gy >= kRadius && gy < H - 15) {            // hashed rows (filter.hpp:601-604)
        const uint8_t* base = reinterpret_cast<const uint8_t*>(sm32 + (ry + kRadius) * kPitchW + 4 + qx);
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int t = 0; t < kMaxTests; t++) {
          if (t < T) {                                           // uniform
            const uint32_t r = eval_test(base, forest, t);
            // bit placement of filter.hpp:574-584: t<8 -> bit t; t==8 -> bit 0 (masked); t>=9 -> bit t-1
            if (t < 8) acc[0] |= (r >> (7 - t)) & (0x01010101u << t);
            else if (t == 8) acc[0] |= (r >> 7) & m8;
            else { const int p = t - 1; acc[p >> 3] |= (r >> (7 - (p & 7))) & (0x01010101u << (p & 7)); }
          }
        }
        // 4x4 byte transpose: state of pixel j = byte j of acc[0..3]
        uint32_t lo01 = __byte_perm(acc[0], acc[1], 0x5140), hi01 = __byte_perm(acc[0], acc[1], 0x7362);
        uint32_t lo23 = __byte_perm(acc[2], acc[3], 0x5140), hi23 = __byte_perm(acc[2], acc[3], 0x7362);
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
  return (size_t)(kRawRows + kSmRows) * kPitch + (size_t)kTileH * (kTileW / 16) * sizeof(uint16_t);
}

cudaError_t configure_preprocess_hash() {   // per device: opt in to > 48 KB dynamic shared memory
  int smem = (int)preprocess_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(preprocess_hash_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(preprocess_hash_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(preprocess_hash_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  return e;
}

cudaError_t launch_preprocess_hash(const PreprocessArgs& args, const ForestDev& forest, int n_img,
                                   int mode, cudaStream_t stream) {
  size_t smem = preprocess_smem_bytes();
  dim3 grid((args.W + kTileW - 1) / kTileW, (args.H + kTileH - 1) / kTileH, n_img);
  if (mode == 2)
    preprocess_hash_kernel<2><<<grid, kThreadsA, smem, stream>>>(args, forest);
  else if (mode == 1)
    preprocess_hash_kernel<1><<<grid, kThreadsA, smem, stream>>>(args, forest);
  else
    preprocess_hash_kernel<0><<<grid, kThreadsA, smem, stream>>>(args, forest);
  return cudaGetLastError();
}

}  // namespace gpc

// smooth_sobel.cu -- kernel A1: box blur + Sobel candidate masks for a whole batch of images.
//
// Replaces the reference's
//   ndb::box + Buffer::clearBoundary   (filter.hpp:293-392, buffer.hpp:630-654)
//   ndb::sobel                         (filter.hpp:404-519, incl. the lane duplication at :504-507)
//   ndb::arr2ind + border lambda       (filter.hpp:60-87, inference.hpp:318-330)
// One CTA stages a (64+2) x (256+32) raw tile in shared memory and produces, for its 256 x 64
// pixels, (a) the smoothed image in BIASED form (s ^ 0x80, the form kernel A2's signed byte
// compares want) and (b) one 16-bit candidate mask per 16-pixel segment.  Every pixel of both
// outputs is written exactly once, borders included, so kernel A2 can fetch arbitrary tiles of
// the smoothed image with TMA.  Nothing here is a translation of the SSE code: the horizontal
// floor-thirds are dp4a row sums, the vertical pass and the Sobel pass slide 3-row register windows, and every
// floor-third / floor-ninth is the byte 2 of a plain 32-bit product (no multiply-high: IMAD.HI issues at less than
// half the rate of IMAD on this machine, scripts/micro/int_pipe_bench.cu).
#include <cuda.h>

#include "gpc_device.cuh"

namespace gpc {

// IMAD.HI is a slow instruction on this machine (measured: replacing 136 of them per thread by plain IMADs took 13.5 %
// off the kernel), so a floor-third that a PRMT picks up anyway is left in BYTE 2 of a plain product:
// s <= 765 -> s * 21846 < 2^24, byte 2 = (s * 21846) >> 16 = the reference's mulhi third, byte 3 = 0.
__device__ __forceinline__ uint32_t third_b2(uint32_t s) { return s * 21846u; }
constexpr uint32_t kThirdByte = 2;
// The same for the ninths of the Sobel sums: s <= 1020 -> s * 7282 < 2^24, byte 2 = (s * 7282) >> 16 <= 113.
__device__ __forceinline__ uint32_t ninth_b2(uint32_t s) { return s * 7282u; }
// |a - b|^2 + |c - d|^2 for four ninths left in byte 2 of pa, pb, pc, pd: two PRMTs pack (a, c, 0, 0) and (b, d, 0, 0),
// VABSDIFF4 takes both differences at once and the dot product of the word with itself sums their squares.
__device__ __forceinline__ uint32_t sq_diff2_b2(uint32_t pa, uint32_t pb, uint32_t pc, uint32_t pd) {
  const uint32_t z = __vabsdiffu4(__byte_perm(pa, pc, 0x7362), __byte_perm(pb, pd, 0x7362));
  return __dp4a(z, z, 0u);
}
__device__ __forceinline__ uint32_t ninth(uint32_t s) { return __umulhi(s, 7282u << 16); }    // (s*7282)>>16

constexpr int kPreW = 256, kPreH = 64;             // output tile
constexpr int kPrePitch = kPreW + 32;              // image cols x0-16 .. x0+kPreW+15
constexpr int kPrePitchW = kPrePitch / 4;
constexpr int kPreRows = kPreH + 2;                // image rows y0-1 .. y0+kPreH
constexpr int kPreThreads = 256;
// TMA staging (kTma): the tile is fetched as two half tiles of 128 output columns, each with its own 16-byte aligned
// halo -- image columns x0 + 128 k - 16 .. x0 + 128 k + 143 -- because a TMA box is at most 256 elements wide.  A
// thread only ever reads the half tile its own output columns belong to.
constexpr int kPreHalfPitch = 128 + 32;            // bytes per row of a half tile
constexpr int kPreHalfPitchW = kPreHalfPitch / 4;
constexpr int kPreHalfBytes = (kPreRows * kPreHalfPitch + 127) / 128 * 128;   // TMA destinations are 128-byte aligned
constexpr int kPreHalfWords = kPreHalfBytes / 4;
constexpr int kPreSmemWords = (2 * kPreHalfWords > kPreRows * kPrePitchW) ? 2 * kPreHalfWords : kPreRows * kPrePitchW;
constexpr int kPreSegRows = kPreH / (kPreThreads / (kPreW / 4));   // rows per thread in the vertical pass

// Horizontal floor-thirds of 4 consecutive pixels: h[k] = (p[x+k-1] + p[x+k] + p[x+k+1]) / 3.
__device__ __forceinline__ void hthirds(uint32_t wm1, uint32_t w, uint32_t wp1, uint32_t h[4]) {
  h[0] = third_b2(__dp4a(__funnelshift_r(wm1, w, 24), 0x00010101u, 0u));      // the third sits in byte kThirdByte
  h[1] = third_b2(__dp4a(w, 0x00010101u, 0u));
  h[2] = third_b2(__dp4a(w, 0x01010100u, 0u));
  h[3] = third_b2(__dp4a(__funnelshift_r(w, wp1, 16), 0x00010101u, 0u));
}

// Sobel predicate (filter.hpp:466-503) for the 4 pixels of one word (image columns x0..x0+3).  With P = raw pixel,
//   A = ninth(col(x-1)), B = ninth(col(x+1)), col(x) = P[y-1][x] + 2 P[y][x] + P[y+1][x]
//   C = ninth(row(y-1)), D = ninth(row(y+1)), row(y) = P[y][x-1] + 2 P[y][x] + P[y][x+1]
// and bit j of the result = ((A-B)^2 + (C-D)^2 > thr2) for pixel x0+j.  Row sums are dp4a on (funnel-shifted)
// words, column sums are formed for two pixels at a time in 16-bit lanes.  A thread walks CONSECUTIVE rows with a
// sliding 3-row window:
// per raw row and quad everything that depends on that row alone is computed once (SobelRow) and used by the three
// pixel rows it touches: the ninths of its horizontal [1 2 1] sums (C of the row below, D of the row above) and its
// pixels split into 16-bit lanes for the vertical [1 2 1] column sums.
struct SobelRow {
  uint32_t rn[4];        // ninth(row(y)) for the quad's four columns
  uint32_t E, O, LR;     // lanes (P[x0], P[x0+2]), (P[x0+1], P[x0+3]), (P[x0-1], P[x0+4])
};
__device__ __forceinline__ SobelRow sobel_row(uint32_t wl, uint32_t w, uint32_t wr) {
  SobelRow r;
  const uint32_t t0 = __funnelshift_r(wl, w, 24), t3 = __funnelshift_r(w, wr, 16);
  r.rn[0] = ninth_b2(__dp4a(t0, 0x00010201u, 0u)); r.rn[1] = ninth_b2(__dp4a(w, 0x00010201u, 0u));
  r.rn[2] = ninth_b2(__dp4a(w, 0x01020100u, 0u)); r.rn[3] = ninth_b2(__dp4a(t3, 0x00010201u, 0u));
  r.E = w & 0x00ff00ffu;
  r.O = __byte_perm(w, 0u, 0x4341);
  r.LR = __byte_perm(__funnelshift_l(wl, wr, 8), 0u, 0x4140);     // (wl byte 3, wr byte 0)
  return r;
}
// rows y-1, y, y+1 of one quad -> bits kBit + 2 j and kBit + 2 j + 1 = predicate of pixel x0 + j (the segment mask
// carries every surviving column twice, filter.hpp:504-507)
template <int kBit>
__device__ __forceinline__ uint32_t sobel_window(const SobelRow& a, const SobelRow& b, const SobelRow& c, int thr2) {
  const uint32_t colE = a.E + c.E + 2u * b.E, colO = a.O + c.O + 2u * b.O, colLR = a.LR + c.LR + 2u * b.LR;   // lanes <= 1020
  const uint32_t n[6] = {ninth_b2(colLR & 0xffffu), ninth_b2(colE & 0xffffu), ninth_b2(colO & 0xffffu),
                         ninth_b2(colE >> 16), ninth_b2(colO >> 16), ninth_b2(colLR >> 16)};
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if ((int)sq_diff2_b2(n[j], n[j + 2], a.rn[j], c.rn[j]) > thr2) m |= 3u << (kBit + 2 * j);     // <= 25538, no int16 wrap / saturation
  }
  return m;
}

// candidate border (inference.hpp:322): 13 <= x < W-13, 13 <= y < H-13, applied to a segment mask
__device__ __forceinline__ uint32_t border_mask(uint32_t m, int gy, int gxs, int W, int H) {
  if (gy < kRadius || gy >= H - kRadius) return 0u;
  const int lowcut = kRadius - gxs;                 // columns below 13
  if (lowcut >= 16) return 0u;
  if (lowcut > 0) m &= ~((1u << lowcut) - 1u);
  const int keep = W - kRadius - gxs;               // columns below W-13
  if (keep <= 0) return 0u;
  if (keep < 16) m &= (1u << keep) - 1u;
  return m;
}

// Word that holds image column x0 + 4 * qw of tile row r (qw = 0 .. kPreW/4 - 1); [-1] and [+1 .. +4] are valid too.
template <bool kTma>
__device__ __forceinline__ const uint32_t* pre_word(const uint32_t* raw32, int r, int qw) {
  if (kTma) return raw32 + (qw >> 5) * kPreHalfWords + r * kPreHalfPitchW + 4 + (qw & 31);
  return raw32 + r * kPrePitchW + 4 + qw;
}

// kDebugOut: also writes the unbiased smooth image and the 0/255 grad image (gpc_preprocess seam).
// kTma: the raw tile arrives as two TMA tensor copies (zero fill outside the image) instead of a staging loop.
template <bool kDebugOut, bool kTma>
__device__ __forceinline__ void smooth_sobel_body(const PreprocessArgs& args, const CUtensorMap* tmap) {
  __shared__ __align__(128) uint32_t raw32[kPreSmemWords];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int cta_last;                                   // largest row of this tile with a candidate
  const int W = args.W, H = args.H;
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * kPreW, y0 = blockIdx.y * kPreH;
  const int tid = threadIdx.x;
  const size_t img_off = (size_t)img * W * H;
  const uint8_t* __restrict__ raw = args.raw + img_off;

  if (threadIdx.x == 0) cta_last = -1;
  // ---- stage the raw tile (zero outside the image) ---------------------------------------------
  if (kTma) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * kPreRows * kPreHalfPitch) : "memory");
#pragma unroll
      for (int k = 0; k < 2; k++)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"((uint32_t)__cvta_generic_to_shared(raw32 + k * kPreHalfWords)), "l"(tmap), "r"(x0 + 128 * k - 16), "r"(y0 - 1), "r"(img), "r"(bar)
            : "memory");
    }
    __syncthreads();                                         // the initialised barrier is visible to all waiters
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar) : "memory");
  } else {
    constexpr int kChunks = kPrePitch / 16;
    uint4* dst = reinterpret_cast<uint4*>(raw32);
    for (int c = tid; c < kPreRows * kChunks; c += kPreThreads) {
      const int r = c / kChunks, k = c - r * kChunks;
      const int gy = y0 - 1 + r, gx = x0 - 16 + 16 * k;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(reinterpret_cast<const uint4*>(raw + (size_t)gy * W + gx));
      dst[c] = v;
    }
    __syncthreads();
  }

  // ---- smoothed pixels: thread = (quad column, band of kPreSegRows rows), 3-row register window ---
  {
    const int q = tid % (kPreW / 4), band = tid / (kPreW / 4);
    const int gxq = x0 + 4 * q;
    if (gxq < W) {
      const int last_written = (H & 1) ? H - 3 : H - 4;    // box writes rows 1..last (filter.hpp:307,388)
      uint32_t colmask = 0xffffffffu;                      // clearBoundary: columns 0,1 and W-1
      if (gxq == 0) colmask = 0xffff0000u;
      if (gxq == W - 4) colmask &= 0x00ffffffu;
      // 3-row window per pixel column, PACKED: byte 0..2 of win[k] = horizontal thirds of rows j-1, j, j+1.  One PRMT
      // shifts the window and inserts the new row (picking the third out of byte 2 of its product, see third_b2), one
      // dp4a sums it.  (History: with the thirds as multiply-highs in separate registers ptxas fused each one into an
      // IMAD.HI with a 64-bit addend and re-computed it for each of the three sums it appears in -- 364 instead of 248
      // IMAD.HI per thread; the packed window fixed that, the plain products then removed the IMAD.HI altogether.)
      uint32_t win[4] = {0u, 0u, 0u, 0u};
      auto push_row = [&](int r) {                         // raw-tile row r = image row y0 - 1 + r
        const uint32_t* row = pre_word<kTma>(raw32, r, q);
        uint32_t h[4];
        hthirds(row[-1], row[0], row[1], h);
#pragma unroll
        for (int k = 0; k < 4; k++) win[k] = __byte_perm(win[k], h[k], 0x7421u + (kThirdByte << 8));   // (b1, b2, h, 0)
      };
      const int j0 = band * kPreSegRows;
      push_row(j0);
      push_row(j0 + 1);
      // one 64-bit address per thread; the rows of its band are 32-bit word offsets from it
      uint32_t* const dst = reinterpret_cast<uint32_t*>(args.smooth_x + img_off + (size_t)(y0 + j0) * W + gxq);
      uint32_t* const dbg = (kDebugOut && args.smooth_out)
                                ? reinterpret_cast<uint32_t*>(args.smooth_out + img_off + (size_t)(y0 + j0) * W + gxq) : nullptr;
      const uint32_t roww = (uint32_t)W / 4u;
      const int rows_here = H - (y0 + j0);                 // rows of the band inside the image
#pragma unroll
      for (int jj = 0; jj < kPreSegRows; jj++) {
        push_row(j0 + jj + 2);
        uint32_t t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = third_b2(__dp4a(win[k], 0x00010101u, 0u));
        constexpr uint32_t kPick = 0x0040u + kThirdByte * 0x11u;       // byte kThirdByte of both operands
        uint32_t v = __byte_perm(__byte_perm(t[0], t[1], kPick), __byte_perm(t[2], t[3], kPick), 0x5410);
        const int gy = y0 + j0 + jj;
        v = (gy < 1 || gy > last_written) ? 0u : (v & colmask);
        if (jj < rows_here) {
          dst[(uint32_t)jj * roww] = v ^ 0x80808080u;
          if (kDebugOut && dbg) dbg[(uint32_t)jj * roww] = v;
        }
      }
    }
  }

  // ---- Sobel predicate per 16-pixel segment -> candidate bit masks, per-row candidate counts ----------
  // thread = (segment sg of the tile row, tile rows 4 * (tid / 16) + it): everything that does not depend on the row is
  // computed once, the rows are compile-time offsets from one shared-memory and one global address
  {
    static_assert((kPreH * (kPreW / 16)) % kPreThreads == 0 && kPreW / 16 == 16, "uniform trip count, 16 segments per tile row");
    constexpr int kSobelIters = kPreH * (kPreW / 16) / kPreThreads;
    const int segs_per_row = W / 16;
    const int sg = tid & 15, ry0 = (tid >> 4) * kSobelIters;
    const int gxs = x0 + 16 * sg;
    const bool col_ok = gxs < W;
    const uint32_t colkeep = border_mask(0xffffu, kRadius, gxs, W, 2 * kRadius + 1);   // columns 13 .. W-14 of this segment
    const uint32_t* const seg0 = pre_word<kTma>(raw32, ry0, 4 * sg);
    constexpr int kRowPitchW = kTma ? kPreHalfPitchW : kPrePitchW;
    uint16_t* const cand0 = args.cand + ((size_t)img * H + y0 + ry0) * segs_per_row + (gxs >> 4);
    int32_t* const rowcnt0 = args.rowcnt + (size_t)img * H + y0 + ry0;
    int my_last = -1;
    SobelRow wa[3], wb[3];                                 // window of quad a (columns s..s+3) and quad b (s+8..s+11)
    auto load_row = [&](int k, int slot) {                 // tile row ry0 + k = image row y0 + ry0 + k - 1
      const uint32_t* row = seg0 + k * kRowPitchW;
      const uint4 v = *reinterpret_cast<const uint4*>(row);
      wa[slot] = sobel_row(row[-1], v.x, v.y);
      wb[slot] = sobel_row(v.y, v.z, v.w);
    };
    load_row(0, 0);
    load_row(1, 1);
#pragma unroll
    for (int it = 0; it < kSobelIters; it++) {
      const int gy = y0 + ry0 + it;
      const bool valid = col_ok && gy < H;
      uint32_t m = 0;
      // rows the reference writes (filter.hpp:517); without the gradient image only the candidate rows are ever looked at
      const bool row_needed = kDebugOut ? (gy >= 1 && gy < H - 3) : (gy >= kRadius && gy < H - kRadius);
      load_row(it + 2, 2);
      if (valid && row_needed) {
        // true columns s..s+3 and s+8..s+11 survive the lane duplication (filter.hpp:504-507): output bits 2g, 2g+1
        m = sobel_window<0>(wa[0], wa[1], wa[2], args.thr2) | sobel_window<8>(wb[0], wb[1], wb[2], args.thr2);
      }
      wa[0] = wa[1]; wa[1] = wa[2]; wb[0] = wb[1]; wb[1] = wb[2];
      if (kDebugOut && args.grad_out && valid) {
        uint32_t wv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint32_t nib = (m >> (4 * k)) & 15u;
          wv[k] = ((nib & 1u) * 0xffu) | ((nib & 2u) * (0xff00u >> 1)) | ((nib & 4u) * (0xff0000u >> 2)) |
                  ((nib & 8u) * (0xff000000u >> 3));
        }
        *reinterpret_cast<uint4*>(args.grad_out + img_off + (size_t)gy * W + gxs) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      }
      const uint32_t bm = (valid && gy >= kRadius && gy < H - kRadius) ? (m & colkeep) : 0u;   // candidate border (inference.hpp:322)
      if (valid) cand0[(uint32_t)(it) * (uint32_t)segs_per_row] = (uint16_t)bm;
      // the 16 segments of one tile row sit in 16 consecutive lanes: one global atomic per tile row
      int cnt = __popc(bm);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
      if (sg == 0 && cnt > 0) {
        atomicAdd(rowcnt0 + it, cnt);
        my_last = gy;                                      // rows ascend with it
      }
    }
    if (my_last >= 0) atomicMax(&cta_last, my_last);
  }
  __syncthreads();
  if (tid == 0 && cta_last >= 0 && cta_last > *reinterpret_cast<volatile int32_t*>(args.lastrow + img))
    atomicMax(args.lastrow + img, cta_last);              // a stale read only costs a redundant atomic
}

#ifndef GPC_A1_MINB
#define GPC_A1_MINB 4       // 64 registers: the sliding Sobel window wants them (measured: 1 -> 0.181, 4 -> 0.171, 5 -> 0.184, 6 -> 0.188 ms)
#endif
template <bool kDebugOut>
__global__ void __launch_bounds__(kPreThreads, GPC_A1_MINB)
smooth_sobel_kernel(const PreprocessArgs args) {
  smooth_sobel_body<kDebugOut, false>(args, nullptr);
}

template <bool kDebugOut>
__global__ void __launch_bounds__(kPreThreads, GPC_A1_MINB)
smooth_sobel_tma_kernel(const __grid_constant__ CUtensorMap tmap, const PreprocessArgs args) {
  smooth_sobel_body<kDebugOut, true>(args, &tmap);
}

// ---- the reference's SSE=OFF build: boxNaive + clearBoundary, sobelNaive (filter.hpp:157-231) --------------------
// Same tile, same outputs, different arithmetic: S = (sum of the 3x3 neighbourhood) / 9 on rows 1..H-3, columns
// 2..W-2; Sobel responses are signed, divided by 9 towards zero, compared in int, one result per pixel (no lane
// duplication).  The reference walks the image as one linear array, so the neighbours of the first / last column
// are the adjacent rows' bytes: the tile is staged by LINEAR address (zero outside the image) to reproduce that in
// the gradient image.  The smoothed image is written UNBIASED (kernel A2's naive modes compare unsigned bytes).
__device__ __forceinline__ int div9_toward_zero(int v) {                     // |v| <= 1020
  const int q = (int)ninth((uint32_t)abs(v));
  return v < 0 ? -q : q;
}

template <bool kDebugOut>
__global__ void __launch_bounds__(kPreThreads)
smooth_sobel_naive_kernel(const PreprocessArgs args) {
  __shared__ __align__(16) uint32_t raw32[kPreRows * kPrePitchW];
  __shared__ int cta_last;
  const int W = args.W, H = args.H;
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * kPreW, y0 = blockIdx.y * kPreH;
  const int tid = threadIdx.x;
  const size_t img_off = (size_t)img * W * H;
  const uint8_t* __restrict__ raw = args.raw + img_off;
  const long long n_pix = (long long)W * H;

  if (threadIdx.x == 0) cta_last = -1;
  {
    constexpr int kChunks = kPrePitch / 16;
    uint4* dst = reinterpret_cast<uint4*>(raw32);
    for (int c = tid; c < kPreRows * kChunks; c += kPreThreads) {
      const int r = c / kChunks, k = c - r * kChunks;
      const long long idx = (long long)(y0 - 1 + r) * W + (x0 - 16 + 16 * k);      // linear memory, as the reference reads it
      uint4 v = make_uint4(0, 0, 0, 0);
      if (idx >= 0 && idx + 16 <= n_pix) v = __ldg(reinterpret_cast<const uint4*>(raw + idx));
      dst[c] = v;
    }
  }
  __syncthreads();
  const uint8_t* raw8 = reinterpret_cast<const uint8_t*>(raw32);

  // ---- smoothed pixels: thread = (quad column, band of rows), 3-row window of horizontal 3-sums ------------------
  {
    const int q = tid % (kPreW / 4), band = tid / (kPreW / 4);
    const int gxq = x0 + 4 * q;
    if (gxq < W) {
      uint32_t colmask = 0xffffffffu;                      // clearBoundary: columns 0,1 and W-1
      if (gxq == 0) colmask = 0xffff0000u;
      if (gxq == W - 4) colmask &= 0x00ffffffu;
      uint32_t ha[4], hb[4], hc[4];
      auto load_h = [&](int r, uint32_t h[4]) {            // raw-tile row r = image row y0 - 1 + r
        const uint32_t* row = raw32 + r * kPrePitchW + 4 + q;
        const uint32_t wm1 = row[-1], w = row[0], wp1 = row[1];
        h[0] = __dp4a(__funnelshift_r(wm1, w, 24), 0x00010101u, 0u);
        h[1] = __dp4a(w, 0x00010101u, 0u);
        h[2] = __dp4a(w, 0x01010100u, 0u);
        h[3] = __dp4a(__funnelshift_r(w, wp1, 16), 0x00010101u, 0u);
      };
      const int j0 = band * kPreSegRows;
      load_h(j0, ha);
      load_h(j0 + 1, hb);
#pragma unroll 4
      for (int j = j0; j < j0 + kPreSegRows; j++) {
        load_h(j + 2, hc);
        uint32_t v = ninth(ha[0] + hb[0] + hc[0]) | (ninth(ha[1] + hb[1] + hc[1]) << 8) |
                     (ninth(ha[2] + hb[2] + hc[2]) << 16) | (ninth(ha[3] + hb[3] + hc[3]) << 24);   // sums <= 2295: exact / 9
        const int gy = y0 + j;
        if (gy < 1 || gy > H - 3) v = 0u; else v &= colmask;      // clearBoundary: row 0, rows H-2, H-1
        if (gy < H) {
          *reinterpret_cast<uint32_t*>(args.smooth_x + img_off + (size_t)gy * W + gxq) = v;
          if (kDebugOut && args.smooth_out)
            *reinterpret_cast<uint32_t*>(args.smooth_out + img_off + (size_t)gy * W + gxq) = v;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) { ha[k] = hb[k]; hb[k] = hc[k]; }
      }
    }
  }

  // ---- Sobel predicate per pixel -> candidate bit masks per 16-pixel segment, per-row candidate counts -------------
  {
    const int segs_per_row = W / 16;
    for (int sr = tid; sr < kPreH * (kPreW / 16); sr += kPreThreads) {
      const int ry = sr / (kPreW / 16), sg = sr - ry * (kPreW / 16);
      const int gy = y0 + ry, gxs = x0 + 16 * sg;
      const bool valid = gy < H && gxs < W;
      uint32_t m = 0;
      // written outputs: linear positions W+1 .. (H-1)*W (filter.hpp:171-185): rows 1..H-2 except (1,0), plus (H-1,0)
      if (valid && gy >= 1 && (gy <= H - 2 || (gy == H - 1 && gxs == 0))) {
        const uint8_t* r0 = raw8 + (ry + 0) * kPrePitch + 16 + 16 * sg;    // image rows gy-1, gy, gy+1 at column gxs
        const uint8_t* r1 = r0 + kPrePitch;
        const uint8_t* r2 = r1 + kPrePitch;
        const int npx = (gy == H - 1) ? 1 : 16;
#pragma unroll 4
        for (int j = 0; j < npx; j++) {
          const int p11 = r0[j - 1], p12 = r0[j], p13 = r0[j + 1];
          const int p21 = r1[j - 1], p23 = r1[j + 1];
          const int p31 = r2[j - 1], p32 = r2[j], p33 = r2[j + 1];
          const int sx = div9_toward_zero(p11 + p31 + 2 * p21 - p13 - 2 * p23 - p33);
          const int sy = div9_toward_zero(p11 + p13 + 2 * p12 - p31 - 2 * p32 - p33);
          if (sx * sx + sy * sy > args.thr2) m |= 1u << j;
        }
        if (gy == 1 && gxs == 0) m &= ~1u;                                 // (1,0) is never written
      }
      if (kDebugOut && args.grad_out && valid) {
        uint32_t wv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint32_t nib = (m >> (4 * k)) & 15u;
          wv[k] = ((nib & 1u) * 0xffu) | ((nib & 2u) * (0xff00u >> 1)) | ((nib & 4u) * (0xff0000u >> 2)) |
                  ((nib & 8u) * (0xff000000u >> 3));
        }
        *reinterpret_cast<uint4*>(args.grad_out + img_off + (size_t)gy * W + gxs) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      }
      const uint32_t bm = valid ? border_mask(m, gy, gxs, W, H) : 0u;
      if (valid) args.cand[((size_t)img * H + gy) * segs_per_row + (gxs >> 4)] = (uint16_t)bm;
      int cnt = __popc(bm);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
      if ((tid & 15) == 0 && cnt > 0) {
        atomicAdd(args.rowcnt + (size_t)img * H + gy, cnt);
        atomicMax(&cta_last, gy);
      }
    }
  }
  __syncthreads();
  if (tid == 0 && cta_last >= 0 && cta_last > *reinterpret_cast<volatile int32_t*>(args.lastrow + img))
    atomicMax(args.lastrow + img, cta_last);
}

// box of one half tile of kernel A1 (see kPreHalfPitch): what the raw-image tensor map is encoded with
void smooth_sobel_tma_box(int* box_w, int* box_h) { *box_w = kPreHalfPitch; *box_h = kPreRows; }

// raw_tmap: tensor map over args.raw ([n_img][H][W] u8, box = smooth_sobel_tma_box) or nullptr (staging loop: the
// raw pointer is not 16-byte aligned, or the naive result mode, whose tile is staged by linear address).
cudaError_t launch_smooth_sobel(const PreprocessArgs& args, int n_img, bool debug_out, const void* raw_tmap, cudaStream_t stream) {
  dim3 grid((args.W + kPreW - 1) / kPreW, (args.H + kPreH - 1) / kPreH, n_img);
  if (args.naive) {
    if (debug_out) smooth_sobel_naive_kernel<true><<<grid, kPreThreads, 0, stream>>>(args);
    else smooth_sobel_naive_kernel<false><<<grid, kPreThreads, 0, stream>>>(args);
    return cudaGetLastError();
  }
  if (raw_tmap) {
    const CUtensorMap& tmap = *reinterpret_cast<const CUtensorMap*>(raw_tmap);
    if (debug_out) smooth_sobel_tma_kernel<true><<<grid, kPreThreads, 0, stream>>>(tmap, args);
    else smooth_sobel_tma_kernel<false><<<grid, kPreThreads, 0, stream>>>(tmap, args);
    return cudaGetLastError();
  }
  if (debug_out) smooth_sobel_kernel<true><<<grid, kPreThreads, 0, stream>>>(args);
  else smooth_sobel_kernel<false><<<grid, kPreThreads, 0, stream>>>(args);
  return cudaGetLastError();
}

// gpc_hash_smooth seam: the caller provides the smoothed image and a u8 flag image; produce the
// biased smooth image and the candidate masks kernel A2 consumes.  One thread per 16-pixel segment.
__global__ void __launch_bounds__(256)
prep_from_smooth_kernel(const uint8_t* __restrict__ smooth, const uint8_t* __restrict__ flags, uint8_t* __restrict__ smooth_x,
                        uint16_t* __restrict__ cand, int32_t* __restrict__ rowcnt, int32_t* __restrict__ lastrow, int W, int H,
                        uint32_t bias) {
  const int segs_per_row = W / 16;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= segs_per_row * H) return;
  const int gy = i / segs_per_row, gxs = 16 * (i - gy * segs_per_row);
  uint4 v = __ldg(reinterpret_cast<const uint4*>(smooth + (size_t)gy * W + gxs));
  v.x ^= bias; v.y ^= bias; v.z ^= bias; v.w ^= bias;       // 0x80808080, or 0 in the naive result mode
  *reinterpret_cast<uint4*>(smooth_x + (size_t)gy * W + gxs) = v;
  const uint4 f = __ldg(reinterpret_cast<const uint4*>(flags + (size_t)gy * W + gxs));
  const uint32_t wv[4] = {f.x, f.y, f.z, f.w};
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int b = 0; b < 4; b++)
      if ((wv[k] >> (8 * b)) & 0xffu) m |= 1u << (4 * k + b);
  const uint32_t bm = border_mask(m, gy, gxs, W, H);
  cand[i] = (uint16_t)bm;
  if (bm) { atomicAdd(rowcnt + gy, __popc(bm)); atomicMax(lastrow, gy); }
}

cudaError_t launch_prep_from_smooth(const uint8_t* smooth, const uint8_t* flags, uint8_t* smooth_x, uint16_t* cand, int32_t* rowcnt,
                                    int32_t* lastrow, int W, int H, int naive, cudaStream_t stream) {
  const int n = (W / 16) * H;
  prep_from_smooth_kernel<<<(n + 255) / 256, 256, 0, stream>>>(smooth, flags, smooth_x, cand, rowcnt, lastrow, W, H,
                                                               naive ? 0u : 0x80808080u);
  return cudaGetLastError();
}

}  // namespace gpc

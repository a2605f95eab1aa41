// gpc_device.cuh -- shared device-side definitions of the B200 Global Patch Collider path.
//
// Data layout in HBM (all resident in the context, see gpc_capi.cu):
//   raw    uint8  [n_img][H][W]        input images, image 2p = left, 2p+1 = right of pair p
//   smooth uint8  [n_img][H][W]        biased smoothed images (s ^ 0x80), kernel A1 -> TMA -> kernel A2
//   cand   uint16 [n_img][H][W/16]     candidate bit masks per 16-pixel segment
//   hash   uint32 [n_img][H][W]        bit 31 = "candidate", bits 0..30 = fern state
//   rowcnt int32  [n_img][H]           candidates per image row
//   lastrow int32 [n_img]              largest row with a candidate (-1 if none)
//   stage  uint32 [n_pair][H][W]       per-row match lists, xL<<16 | xR, sorted by state
//   rowmatch int32 [n_pair][H]         matches per row
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#include <cuda_runtime.h>
#else   // NVRTC (forest-specialised kernel A2, see jit.cu): no host headers
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
typedef unsigned long size_t;
#endif

namespace gpc {

constexpr int kRadius = 13;            // patch radius / candidate border (inference.hpp:322)
constexpr int kMaxTests = 32;          // inference.hpp:426
constexpr uint32_t kCandFlag = 0x80000000u;

// ---- kernel A2 (hash_tiles.cu) tile geometry ---------------------------------------------------
#ifndef GPC_TILE_W
#define GPC_TILE_W 128
#endif
#ifndef GPC_TILE_H
#define GPC_TILE_H 64
#endif
#ifndef GPC_THREADS_A
#define GPC_THREADS_A 256
#endif
constexpr int kTileW = GPC_TILE_W;             // output pixels per tile row; kTileW + 32 <= 256 (TMA box limit)
constexpr int kTileH = GPC_TILE_H;             // output rows per tile
constexpr int kPitch = kTileW + 32;            // smem row pitch in bytes: image cols x0-16 .. x0+kTileW+15
constexpr int kPitchW = kPitch / 4;            // ... in 32-bit words
constexpr int kSmRows = kTileH + 2 * kRadius;  // image rows y0-13 .. y0+kTileH+12
constexpr int kCopyBytes = (kSmRows * kPitch + 127) / 128 * 128;   // one copy of the smoothed tile (TMA destination: 128-byte aligned)
constexpr int kThreadsA = GPC_THREADS_A;
constexpr int kQuadsX = kTileW / 4;            // quads (4 pixels) per tile row
static_assert(kPitch <= 256 && kSmRows <= 256, "TMA box dimensions are limited to 256");
static_assert(kThreadsA % kQuadsX == 0 && kTileH % (kThreadsA / kQuadsX) == 0, "thread block must tile the rows evenly");

// Forest baked for the kernel's shared-memory layout (replaces the per-width baking of
// inference.hpp:427-428).  The smoothed tile is kept four times, copy k shifted left by k bytes,
// so that the 4-pixel operand of every test is ONE aligned 32-bit load: imm = k * kCopyBytes +
// (dy * kPitch + dx - k) with k = (dy * kPitch + dx) mod 4.  Passed by value as a kernel
// parameter: with the test loop fully unrolled every field is read from a fixed constant-bank
// address (uniform datapath), never through an indexed load.
// Result modes.  kResultsSse: the reference's default build (-D_INTRINSICS_SSE).  kResultsNaive: its SSE=OFF build
// (the *Naive functions of filter.hpp: 3x3 sum / 9, signed Sobel, plain int tau compare, first test in the
// highest bit, every candidate row hashed); the smoothed image is then kept UNBIASED and the fern tests are
// assigned to the kernel's test slots by bit position.
constexpr int kResultsSse = 0, kResultsNaive = 1;

struct ForestDev {
  int32_t n_tests;               // test slots in use (naive mode: highest slot + 1)
  int32_t type;                  // 0: a > b ; 1: a > sat_int8(b - tau)  (naive mode: a > b - tau in int)
  int32_t naive;                 // kResultsSse / kResultsNaive
  int32_t imm_a[kMaxTests];      // byte offset of operand a's word relative to the quad's word in copy 0
  int32_t imm_b[kMaxTests];
  uint32_t mtau2[kMaxTests];     // -tau (int8 tau) as int16 replicated into both 16-bit lanes; 0 = no tau
                                 // (naive mode: 32768 - tau in both lanes, tau clamped to +-256)
};

struct PreprocessArgs {     // kernel A1; all pointers already offset to the first image of the launch
  const uint8_t* raw;      // [n_img][H][W]
  uint8_t* smooth_x;       // [n_img][H][W] biased smoothed image (s ^ 0x80)
  uint16_t* cand;          // [n_img][H][W/16] candidate bit masks per 16-pixel segment
  int32_t* rowcnt;         // [n_img][H] candidates per row (zeroed before the launch)
  int32_t* lastrow;        // [n_img] largest row with a candidate (-1 before the launch)
  uint8_t* smooth_out;     // optional [n_img][H][W] (debug seam: unbiased)
  uint8_t* grad_out;       // optional [n_img][H][W] (debug seam: 0 / 255)
  int32_t W, H;
  int32_t thr2;            // (int16)(thr*thr), filter.hpp:418  (naive mode: thr*thr as int, filter.hpp:159)
  int32_t naive;           // kResultsNaive: boxNaive / sobelNaive semantics, smooth_x unbiased
};

struct HashArgs {           // kernel A2; pointers are buffer BASES, img0 = first image of the launch
  const uint16_t* cand;    // [..][H][W/16]
  uint32_t* hash;          // [..][H][W]
  int32_t W, H;
  int32_t img0;
  int32_t hash_y_end;      // rows kRadius <= y < hash_y_end are hashed: H - 15 (filter.hpp:601-604), naive mode H - 13
};

struct MatchArgs {
  const uint32_t* hash;    // [2*n_pair][H][W]
  const int32_t* lastrow;  // [2*n_pair]
  const int32_t* rowcnt;   // [2*n_pair][H] candidates per image row (kernel A1)
  uint32_t* stage;         // [n_pair][H][W]
  int32_t* rowmatch;       // [n_pair][H]
  int32_t W, H;
  int32_t disp_high, vertical_tolerance;
  int32_t table_log2;      // log2 of the per-row bucket count (>= 256, >= 2 * W, ~ 4 * candidates per side)
  int32_t x_bits;          // bits needed for a column index (ceil(log2 W))
  int32_t wcap;            // per-side candidate capacity used to carve shared memory
  int32_t pow2cap;         // wcap rounded up to a power of two
  int32_t key_bits;        // number of significant state bits (forest dependent)
  // fast row matcher (match_rows_fast_kernel)
  int32_t nib_log2;        // log2 of the words of the 4-bit bucket table (8 buckets per word)
  int32_t slot_log2;       // log2 of the slot table; x_bits <= slot_log2 <= nib_log2
  uint32_t* fb_hdr;        // rows handed to the general kernel: hdr[0] count, hdr[1] CTAs done; ent = (pair, row) words
  uint32_t* fb_ent;
  uint32_t* big_hdr;       // rows handed to the block-wide ordering kernel, same layout
  uint32_t* big_ent;
  unsigned long long* mrec;  // [n_pair][H][W] unordered match records of a row: key << 32 | xl << 16 | xr
  uint32_t* ovbuf;         // [n_pair][H][4][kOvCap] overflow lists of a row: left v, left x, right v, right x
  int32_t* rowhdr;         // [n_pair][H][4] matches, left overflow, right overflow, done (1: nothing left for the tail kernel)
};

}  // namespace gpc

// gpc_device.cuh -- shared device-side definitions of the B200 Global Patch Collider path.
//
// Data layout in HBM (all resident in the context, see gpc_capi.cu):
//   raw    uint8  [n_img][H][W]        input images, image 2p = left, 2p+1 = right of pair p
//   hash   uint32 [n_img][H][W]        bit 31 = "candidate", bits 0..30 = fern state
//   rowcnt int32  [n_img][H]           candidates per image row
//   lastrow int32 [n_img]              largest row with a candidate (-1 if none)
//   stage  uint32 [n_pair][H][W]       per-row match lists, xL<<16 | xR, sorted by state
//   rowmatch int32 [n_pair][H]         matches per row
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gpc {

constexpr int kRadius = 13;            // patch radius / candidate border (inference.hpp:322)
constexpr int kMaxTests = 32;          // inference.hpp:426
constexpr uint32_t kCandFlag = 0x80000000u;

// ---- kernel A tile geometry -------------------------------------------------------------
#ifndef GPC_TILE_W
#define GPC_TILE_W 256
#endif
#ifndef GPC_TILE_H
#define GPC_TILE_H 32
#endif
#ifndef GPC_THREADS_A
#define GPC_THREADS_A 256
#endif
constexpr int kTileW = GPC_TILE_W;             // output pixels per tile row (multiple of 128)
constexpr int kTileH = GPC_TILE_H;             // output rows per tile
constexpr int kPitch = kTileW + 32;            // smem row pitch in bytes: image cols x0-16 .. x0+kTileW+15
constexpr int kPitchW = kPitch / 4;            // ... in 32-bit words
constexpr int kRawRows = kTileH + 2 * kRadius + 2;   // image rows y0-14 .. y0+kTileH+13
constexpr int kSmRows = kTileH + 2 * kRadius;        // image rows y0-13 .. y0+kTileH+12
constexpr int kCopyBytes = kSmRows * kPitch;   // one copy of the smoothed tile
constexpr int kThreadsA = GPC_THREADS_A;
constexpr int kQuadsX = kTileW / 4;            // quads (4 pixels) per tile row
static_assert(kThreadsA % kQuadsX == 0 && kThreadsA >= kPitchW, "thread block must cover whole quad rows");

// Forest baked for the kernel's shared-memory layout (replaces the per-width baking of
// inference.hpp:427-428).  The smoothed tile is kept four times, copy k shifted left by k bytes,
// so that the 4-pixel operand of every test is ONE aligned 32-bit load: imm = k * kCopyBytes +
// (dy * kPitch + dx - k) with k = (dy * kPitch + dx) mod 4.  Passed by value as a kernel
// parameter: with the test loop fully unrolled every field is read from a fixed constant-bank
// address (uniform datapath), never through an indexed load.
struct ForestDev {
  int32_t n_tests;
  int32_t type;                  // 0: a > b ; 1: a > sat_int8(b - tau)
  int32_t imm_a[kMaxTests];      // byte offset of operand a's word relative to the quad's word in copy 0
  int32_t imm_b[kMaxTests];
  uint32_t mtau2[kMaxTests];     // -tau (int8 tau) as int16 replicated into both 16-bit lanes; 0 = no tau
  uint32_t pmul[kMaxTests];      // 1 << (bit position of the test inside its state byte); kept in the constant
                                 // bank so that the accumulate stays one IMAD.WIDE (an immediate power of two
                                 // is strength-reduced to five ALU-pipe instructions)
};

struct PreprocessArgs {
  const uint8_t* raw;      // [n_img][H][W]
  uint32_t* hash;          // [n_img][H][W]
  int32_t* rowcnt;         // [n_img][H]
  int32_t* lastrow;        // [n_img]
  uint8_t* smooth_out;     // optional [n_img][H][W]
  uint8_t* grad_out;       // optional [n_img][H][W]
  const uint8_t* flags;    // mode 2 only: [n_img][H][W], non-zero = hash this pixel
  int32_t W, H;
  int32_t thr2;            // (int16)(thr*thr), filter.hpp:418
};

struct MatchArgs {
  const uint32_t* hash;    // [2*n_pair][H][W]
  const int32_t* lastrow;  // [2*n_pair]
  uint32_t* stage;         // [n_pair][H][W]
  int32_t* rowmatch;       // [n_pair][H]
  int32_t W, H;
  int32_t disp_high, vertical_tolerance;
  int32_t table_log2;      // log2 of the per-row bucket count (>= 256, >= 2 * W, ~ 4 * candidates per side)
  int32_t x_bits;          // bits needed for a column index (ceil(log2 W))
  int32_t wcap;            // per-side candidate capacity used to carve shared memory
  int32_t pow2cap;         // wcap rounded up to a power of two
  int32_t key_bits;        // number of significant state bits (forest dependent)
};

}  // namespace gpc

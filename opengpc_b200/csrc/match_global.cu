// match_global.cu -- device-wide matcher: LSD radix sort of (key, side, index) records followed
// by a segmented scan that keeps keys occurring exactly once on each side.
//
// Replaces Forest::findCorrespondences (inference.hpp:227-254: two std::sort + merge scan) for
// the cases the per-row matcher of match_rows.cu does not cover:
//   * global mode, epipolarMode(false): key = 32-bit state (inference.hpp:184-202), result
//     filtered by |dy| <= verticalTolerance and |dx| <= dispHigh (inference.hpp:384-391);
//   * explicit descriptor lists with arbitrary 64-bit keys (the C++ API's findCorrespondences).
//
// Record order before the sort is "all left records in input (raster) order, then all right
// records"; the sort is stable, so inside a run of equal keys left precedes right and each side
// keeps its input order.  That makes the reference's rules local tests on the sorted array:
//   k != tmax : match iff the run is exactly {L, R}
//   k == tmax : (tmax = largest right key) match iff the run is exactly {L, R, R}; the partner is
//               the first R (the reference's one implementation-defined case);
//               a single R at the very tail never matches (inference.hpp:243-249).
// Output order = position in the sorted array = ascending key, as std::sort gives the reference.
//
// useHashtable(true) (inference.hpp:204-225, ndb::Hashmatch hashmatch.hpp:48-272) runs on the same machinery:
// the sort key is the bucket index key % 214673, the stable sort leaves each bucket's records in insertion
// order (all left, then all right, each in raster order), and one thread replays the bucket: the first 10
// records in stable ascending key order, walked by the reference's pairing rules.
//
// Pre-filter (hash-image input, sort-path semantics): a record can only take part in a match, or spoil one, if
// its key also occurs on the OTHER side.  Each side's keys are marked in a bit table (2^22 bits per side and pair,
// indexed by a hash of the key), and only records whose bit is set in the other side's table are gathered and
// sorted -- a superset of every run that holds both sides, complete for each such key, in the same raster order,
// so the sorted runs the rules above look at are unchanged while ~85 % of the records never enter the sort.
// The largest right key is taken over ALL right records (per-row maxima, then a reduction), not from the
// filtered array.  The hashtable matcher does not use the filter: its buckets depend on every record.
//
// Every kernel carries a pair dimension (blockIdx.y): a chunk of independent pairs is sorted by the
// same launches, each pair in its own slice of the workspace.
#include <algorithm>
#include <cstdlib>

#include "gpc_device.cuh"

namespace gpc {

constexpr int kSortThreads = 512;           // 16 warps; two to four blocks per SM
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRounds = 8;                  // keys per thread: a block owns a tile of kRounds * 512 keys; warp w takes the
                                            // w-th run of 32 * kRounds consecutive keys, 32 per round
constexpr int kTile = kSortThreads * kRounds;
constexpr int kDigits = 256;
constexpr uint32_t kSideBit = 0x80000000u;
constexpr uint32_t kHtBuckets = 214673u;    // inference.hpp:212
constexpr int kHtDepth = 10;                // hashmatch.hpp:95: a bucket keeps the first 10 elements offered to it
#ifndef GPC_BLOOM_LOG2
#define GPC_BLOOM_LOG2 22
#endif
constexpr int kBloomLog2 = GPC_BLOOM_LOG2;  // pre-filter: bits per side and pair (2^22: 512 KB; a chunk of pairs stays L2 resident)
constexpr int kBloomWords = (1 << kBloomLog2) / 32;

// Workspace of a chunk of pairs (device pointers; slice `pair` starts at pair * stride of each array).
template <typename KeyT>
struct SortWs {
  KeyT* keys[2];                 // [n_pairs][rec_stride] ping-pong
  uint32_t* vals[2];             // [n_pairs][rec_stride] side << 31 | index
  uint32_t* blockhist;           // [n_pairs][kDigits][nb_max]
  uint32_t* digit_tot;           // [n_pairs][kDigits]
  int32_t* blockcount;           // [n_pairs][nb_max + 1]
  int32_t* n_side;               // [n_pairs][2] left / right record counts
  unsigned long long* tmax;      // [n_pairs] largest right key
  uint32_t* tmax_has;            // [n_pairs] 1 if there is a right record (a key of 2^64 - 1 is legal, so no sentinel value)
  int32_t* rowoff;               // [n_pairs][2][H] candidate offsets (hash-image input only)
  int32_t* rowcnt_f;             // [n_pairs][2][H] records per row that pass the pre-filter
  unsigned long long* rowmax;    // [n_pairs][H] largest right key + 1 of the row (0: none)
  uint32_t* bloom;               // [n_pairs][2][kBloomWords] key bit tables of the two sides
  uint32_t* keep;                // [n_pairs][2][H][keep_words] pre-filter result, one bit per pixel
  int keep_words;                // ceil(W / 32)
  long long rec_stride;
  int nb_max;
};

// ---- record gathering from hash images ----------------------------------------------------------
// rowoff[pair][side][y] = exclusive prefix of the row's candidate count inside its image.
__global__ void __launch_bounds__(1024)
global_rowoff_kernel(const int32_t* __restrict__ rowcnt, int H, int32_t* __restrict__ rowoff_all, int32_t* __restrict__ n_side_all) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int side = blockIdx.x, pair = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int32_t* cnt = rowcnt + ((size_t)2 * pair + side) * H;
  int32_t* rowoff = rowoff_all + ((size_t)2 * pair + side) * H;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int y0 = 0; y0 < H; y0 += 1024) {
    const int y = y0 + tid;
    const int v = (y < H) ? cnt[y] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sums[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    const int base = carry + warp_sums[wid];
    if (y < H) rowoff[y] = base + incl - v;
    __syncthreads();
    if (tid == 1023) carry = base + incl;
    __syncthreads();
  }
  if (tid == 0) n_side_all[2 * pair + side] = carry;
}

// ---- pre-filter ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bloom_hash(unsigned long long key) {
  const uint32_t h = (uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA6Bu;
  return h >> (32 - kBloomLog2);
}

// One warp per (side, row): mark the row's keys in the side's bit table; per-row maximum of the right keys.
__global__ void __launch_bounds__(128)
bloom_build_kernel(const uint32_t* __restrict__ hash, int W, int H, int epipolar, const SortWs<uint32_t> ws) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, pair = blockIdx.y;
  if (warp >= 2 * H) return;
  const int side = warp / H, y = warp - side * H;
  const uint32_t* row = hash + ((size_t)(2 * pair + side) * H + y) * W;
  uint32_t* table = ws.bloom + ((size_t)2 * pair + side) * kBloomWords;
  unsigned long long kmax = 0;
  for (int x0 = 0; x0 < W; x0 += 128) {
    uint32_t v4[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { const int x = x0 + 32 * q + lane; v4[q] = (x < W) ? row[x] : 0u; }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (v4[q] >> 31) {
        unsigned long long k = v4[q] & 0x7fffffffu;
        if (epipolar) k |= (unsigned long long)y << 32;
        const uint32_t b = bloom_hash(k), bit = 1u << (b & 31);
        // test first: repeated keys (flat images) must not pile atomics on one word
        if (!(*reinterpret_cast<volatile uint32_t*>(table + (b >> 5)) & bit)) atomicOr(table + (b >> 5), bit);
        kmax = max(kmax, k + 1ull);
      }
    }
  }
  if (side == 1) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
    if (lane == 0) ws.rowmax[(size_t)pair * H + y] = kmax;
  }
}

// tmax[pair] = largest right key over all rows, tmax_has[pair] = there is a right record
__global__ void __launch_bounds__(32)
rowmax_reduce_kernel(const SortWs<uint32_t> ws, int H) {
  const int pair = blockIdx.x, lane = threadIdx.x;
  unsigned long long m = 0;
  for (int y = lane; y < H; y += 32) m = max(m, ws.rowmax[(size_t)pair * H + y]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if (lane == 0) { ws.tmax[pair] = m ? m - 1ull : 0ull; ws.tmax_has[pair] = m ? 1u : 0u; }
}

// One warp per (side, row): keep[...] = candidates whose key is marked in the OTHER side's table; rowcnt_f = their count.
__global__ void __launch_bounds__(128)
bloom_filter_kernel(const uint32_t* __restrict__ hash, int W, int H, int epipolar, const SortWs<uint32_t> ws) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, pair = blockIdx.y;
  if (warp >= 2 * H) return;
  const int side = warp / H, y = warp - side * H;
  const uint32_t* row = hash + ((size_t)(2 * pair + side) * H + y) * W;
  const uint32_t* other = ws.bloom + ((size_t)2 * pair + (side ^ 1)) * kBloomWords;
  uint32_t* keep = ws.keep + (((size_t)2 * pair + side) * H + y) * ws.keep_words;
  int count = 0;
  for (int x0 = 0; x0 < W; x0 += 128) {
    uint32_t v4[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { const int x = x0 + 32 * q + lane; v4[q] = (x < W) ? row[x] : 0u; }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      bool c = false;
      if (v4[q] >> 31) {
        unsigned long long k = v4[q] & 0x7fffffffu;
        if (epipolar) k |= (unsigned long long)y << 32;
        const uint32_t b = bloom_hash(k);
        c = (__ldg(other + (b >> 5)) >> (b & 31)) & 1u;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, c);
      if (x0 + 32 * q < W && lane == 0) keep[(x0 >> 5) + q] = m;
      count += __popc(m);
    }
  }
  if (lane == 0) ws.rowcnt_f[((size_t)2 * pair + side) * H + y] = count;
}

// One warp per (side, row): candidates in raster order.
template <typename KeyT>
__global__ void __launch_bounds__(128)
global_gather_kernel(const uint32_t* __restrict__ hash, int W, int H, int epipolar, int hashtable, int filtered, const SortWs<KeyT> ws) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, pair = blockIdx.y;
  if (warp >= 2 * H) return;
  const int side = warp / H, y = warp - side * H;
  const uint32_t* row = hash + ((size_t)(2 * pair + side) * H + y) * W;
  const int32_t* n_side = ws.n_side + 2 * pair;
  KeyT* keys = ws.keys[0] + (size_t)pair * ws.rec_stride;
  uint32_t* vals = ws.vals[0] + (size_t)pair * ws.rec_stride;
  int off = ws.rowoff[((size_t)2 * pair + side) * H + y] + (side ? n_side[0] : 0);
  const uint32_t* keepw = filtered ? ws.keep + (((size_t)2 * pair + side) * H + y) * ws.keep_words : nullptr;
  for (int x0 = 0; x0 < W; x0 += 128) {                  // four 32-pixel groups per trip, their loads issued together
    uint32_t v4[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { const int x = x0 + 32 * q + lane; v4[q] = (x < W) ? row[x] : 0u; }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int x = x0 + 32 * q + lane;
      const uint32_t v = v4[q];
      const uint32_t b = filtered ? ((x0 + 32 * q < W) ? keepw[(x0 >> 5) + q] : 0u) : __ballot_sync(0xffffffffu, (v >> 31) != 0u);
      const bool c = (b >> lane) & 1u;
      if (c) {
        const int p = off + __popc(b & ((1u << lane) - 1u));
        unsigned long long k = v & 0x7fffffffu;
        if (epipolar) k |= (unsigned long long)y << 32;
        if (hashtable) k %= kHtBuckets;                  // buffer.hpp:84-86
        keys[p] = (KeyT)k;
        vals[p] = (side ? kSideBit : 0u) | (uint32_t)(y * W + x);
      }
      off += __popc(b);
    }
  }
}

// Lanes of the warp (among `amask`) holding the same 8-bit digit.  Eight ballots instead of match.any, which
// iterates once per distinct value (a warp of hash digits has ~30): measured 153 vs 183 us per scatter pass.
__device__ __forceinline__ uint32_t digit_peers(uint32_t amask, uint32_t d) {
  uint32_t peers = amask;
#pragma unroll
  for (int b = 0; b < 8; b++) {
    const bool bit = (d >> b) & 1u;
    const uint32_t bal = __ballot_sync(amask, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

// kRounds consecutive keys from a 16-byte aligned address.
__device__ __forceinline__ void load_keys16(const uint32_t* p, uint32_t (&k)[kRounds]) {
#pragma unroll
  for (int q = 0; q < kRounds / 4; q++) {
    const uint4 v = reinterpret_cast<const uint4*>(p)[q];
    k[4 * q] = v.x; k[4 * q + 1] = v.y; k[4 * q + 2] = v.z; k[4 * q + 3] = v.w;
  }
}
__device__ __forceinline__ void load_keys16(const unsigned long long* p, unsigned long long (&k)[kRounds]) {
#pragma unroll
  for (int q = 0; q < kRounds / 2; q++) {
    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(p)[q];
    k[2 * q] = v.x; k[2 * q + 1] = v.y;
  }
}

// ---- LSD radix sort, 8-bit digits; a block owns a tile of kTile keys, visited in kRounds rounds of 1024
// consecutive keys (round-major order = input order, which keeps the sort stable) ------------------------
template <typename KeyT>
__device__ __forceinline__ void radix_hist_kernel_tile(const SortWs<KeyT> ws, int cur, int shift, const int tile) {
  __shared__ uint32_t tot[kDigits];
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kTile - 1) / kTile;
  if (tile >= nb) return;
  const KeyT* keys = (cur ? ws.keys[1] : ws.keys[0]) + (size_t)pair * ws.rec_stride;
  const int tid = threadIdx.x;
  if (tid < kDigits) tot[tid] = 0u;
  // the histogram does not care about order: each thread takes kRounds consecutive keys with 16-byte loads
  const int i0 = tile * kTile + kRounds * tid;
  KeyT k[kRounds];
  const bool full = i0 + kRounds <= n && (reinterpret_cast<uintptr_t>(keys + i0) & 15u) == 0u;
  if (full) {
    load_keys16(keys + i0, k);
  } else {
#pragma unroll
    for (int r = 0; r < kRounds; r++) k[r] = (i0 + r < n) ? keys[i0 + r] : (KeyT)0;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const bool active = i0 + r < n;
    // plain shared-memory atomics: measured 30 us per 26 M keys against 76 us with one atomic per distinct digit
    // (the peer search costs more ALU work than the conflicts it avoids); a warp-uniform digit -- key bits the
    // forest never sets -- would serialise 32-fold and is counted by one lane instead
    const uint32_t d = (uint32_t)(k[r] >> shift) & 0xffu;
    int uniform = 0;
    __match_all_sync(0xffffffffu, active ? d : 0x100u, &uniform);
    if (uniform) { if ((tid & 31) == 0 && active) atomicAdd(&tot[d], 32u); }
    else if (active) atomicAdd(&tot[d], 1u);
  }
  __syncthreads();
  if (tid < kDigits) ws.blockhist[((size_t)pair * kDigits + tid) * ws.nb_max + tile] = tot[tid];
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const SortWs<KeyT> ws, int cur, int shift) {
  // a block walks the tiles tile = blockIdx.x, blockIdx.x + gridDim.x, ... of its pair
  for (int tile = blockIdx.x;; tile += gridDim.x) {
    const int n = ws.n_side[2 * blockIdx.y] + ws.n_side[2 * blockIdx.y + 1];
    if (tile * kTile >= n) return;
    radix_hist_kernel_tile<KeyT>(ws, cur, shift, tile);
    __syncthreads();
  }
}

// per-digit exclusive scan over the nb active blocks (one warp per digit, coalesced 32-wide
// chunks); digit_tot[d] = number of keys with digit d.  The cross-digit base is added by the
// scatter kernel.  Grid (kDigits / 32, n_pairs).
template <typename KeyT>
__global__ void __launch_bounds__(1024)
radix_scan_kernel(const SortWs<KeyT> ws) {
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kTile - 1) / kTile;
  const int d = blockIdx.x * 32 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  uint32_t* row = ws.blockhist + ((size_t)pair * kDigits + d) * ws.nb_max;
  uint32_t carry = 0;
  for (int b0 = 0; b0 < nb; b0 += 32) {
    const int b = b0 + lane;
    const uint32_t c = (b < nb) ? row[b] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, k); if (lane >= k) incl += t; }
    if (b < nb) row[b] = carry + incl - c;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) ws.digit_tot[(size_t)pair * kDigits + d] = carry;
}

// Shared memory of the scatter kernel: a region used first as per-warp digit counters, then as the staging
// area of the tile sorted by digit, followed by four 256-entry tables.
template <typename KeyT>
__host__ __device__ constexpr int scatter_region_bytes() {
  return (kSortWarps * kDigits * 4 > kTile * ((int)sizeof(KeyT) + 4)) ? kSortWarps * kDigits * 4 : kTile * ((int)sizeof(KeyT) + 4);
}
template <typename KeyT>
__host__ __device__ constexpr int scatter_smem_bytes() { return scatter_region_bytes<KeyT>() + (kSortWarps / 8 + 3) * kDigits * 4; }

template <typename KeyT>
__device__ __forceinline__ void radix_scatter_kernel_tile(const SortWs<KeyT> ws, int cur, int shift, const int tile) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int kGroups = kSortWarps / 8;                // warps are prefix-summed in groups of eight
  static_assert(kGroups * kDigits == kSortThreads, "one thread per (digit, group of warps)");
  uint32_t (*whist)[kDigits] = reinterpret_cast<uint32_t (*)[kDigits]>(smem);          // [warp][digit] counts -> exclusive prefix
  KeyT* skeys = reinterpret_cast<KeyT*>(smem);                                            // later: tile sorted by digit
  uint32_t* svals = reinterpret_cast<uint32_t*>(smem + kTile * sizeof(KeyT));
  uint32_t* gtot = reinterpret_cast<uint32_t*>(smem + scatter_region_bytes<KeyT>());      // [group][digit]
  uint32_t* racc = gtot + kGroups * kDigits;             // the tile's keys with digit d
  uint32_t* loff = racc + kDigits;                       // exclusive scan of racc
  uint32_t* dbase = loff + kDigits;                      // global position of the tile's first key with digit d
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kTile - 1) / kTile;
  if (tile >= nb) return;
  const KeyT* keys_in = (cur ? ws.keys[1] : ws.keys[0]) + (size_t)pair * ws.rec_stride;
  const uint32_t* vals_in = (cur ? ws.vals[1] : ws.vals[0]) + (size_t)pair * ws.rec_stride;
  KeyT* keys_out = (cur ? ws.keys[0] : ws.keys[1]) + (size_t)pair * ws.rec_stride;
  uint32_t* vals_out = (cur ? ws.vals[0] : ws.vals[1]) + (size_t)pair * ws.rec_stride;
  const uint32_t* digit_tot = ws.digit_tot + (size_t)pair * kDigits;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tile0 = tile * kTile;
  const int w0 = tile0 + wid * (32 * kRounds) + lane;    // the thread's record of round r is w0 + 32 r
  // all of the thread's records first, so that every load is in flight at once
  KeyT k[kRounds];
  uint32_t v[kRounds], lr[kRounds];
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const int i = w0 + 32 * r;
    k[r] = (i < n) ? keys_in[i] : (KeyT)0;
    v[r] = (i < n) ? vals_in[i] : 0u;
  }
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
#pragma unroll
    for (int q = 0; q < kSortWarps * kDigits / 4 / kSortThreads; q++) z[q * kSortThreads + tid] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (wid == kSortWarps - 1) {                           // one warp: exclusive scan of the 256 digit totals
    uint32_t c[8], sum = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { c[q] = digit_tot[8 * lane + q]; sum += c[q]; }
    uint32_t incl = sum;
#pragma unroll
    for (int q = 1; q < 32; q <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, q); if (lane >= q) incl += t; }
    uint32_t run = incl - sum;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      dbase[8 * lane + q] = run + ws.blockhist[((size_t)pair * kDigits + 8 * lane + q) * ws.nb_max + tile];
      run += c[q];
    }
  }
  __syncthreads();
  // ranking inside the warp's own run of keys: the warp's counters need no block-wide barrier
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const bool active = w0 + 32 * r < n;
    const uint32_t amask = __ballot_sync(0xffffffffu, active);
    const uint32_t d = (uint32_t)(k[r] >> shift) & 0xffu;
    uint32_t rank = 0, prev = 0, peers = 0;
    if (active) {
      peers = digit_peers(amask, d);
      rank = __popc(peers & ((1u << lane) - 1u));
      prev = whist[wid][d];
    }
    __syncwarp();
    if (active && rank == 0) whist[wid][d] = prev + __popc(peers);
    __syncwarp();
    lr[r] = prev + rank;                                 // rank among the warp's records with this digit
  }
  __syncthreads();
  const int pd = tid & (kDigits - 1), pg = tid >> 8;     // digit pd, warps 8 * pg .. 8 * pg + 7
  {                                                      // exclusive prefix over the warps, in two levels
    uint32_t acc = 0;
#pragma unroll
    for (int w = 8 * pg; w < 8 * pg + 8; w++) { const uint32_t c = whist[w][pd]; whist[w][pd] = acc; acc += c; }
    gtot[pg * kDigits + pd] = acc;
  }
  __syncthreads();
  {
    uint32_t total = 0, base = 0;
#pragma unroll
    for (int g = 0; g < kGroups; g++) { const uint32_t c = gtot[g * kDigits + pd]; if (g < pg) base += c; total += c; }
    if (base)
#pragma unroll
      for (int w = 8 * pg; w < 8 * pg + 8; w++) whist[w][pd] += base;
    if (pg == 0) racc[pd] = total;
  }
  __syncthreads();
  if (wid == 0) {                                        // exclusive scan of the tile's digit counts
    uint32_t c[8], sum = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) { c[q] = racc[8 * lane + q]; sum += c[q]; }
    uint32_t incl = sum;
#pragma unroll
    for (int q = 1; q < 32; q <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, q); if (lane >= q) incl += t; }
    uint32_t run = incl - sum;
#pragma unroll
    for (int q = 0; q < 8; q++) { loff[8 * lane + q] = run; run += c[q]; }
  }
#pragma unroll
  for (int r = 0; r < kRounds; r++) lr[r] += whist[wid][(uint32_t)(k[r] >> shift) & 0xffu];
  __syncthreads();                                       // whist is dead from here: the region becomes the staging area
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    if (w0 + 32 * r < n) {
      const uint32_t pos = loff[(uint32_t)(k[r] >> shift) & 0xffu] + lr[r];
      skeys[pos] = k[r];
      svals[pos] = v[r];
    }
  }
  __syncthreads();
  const int tile_n = min(kTile, n - tile0);
#pragma unroll
  for (int r = 0; r < kRounds; r++) {                    // runs of equal digits leave as contiguous stores
    const int j = r * kSortThreads + tid;
    if (j < tile_n) {
      const KeyT key = skeys[j];
      const uint32_t d = (uint32_t)(key >> shift) & 0xffu;
      const uint32_t pos = dbase[d] + ((uint32_t)j - loff[d]);
      keys_out[pos] = key;
      vals_out[pos] = svals[j];
    }
  }
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads, 2)
radix_scatter_kernel(const SortWs<KeyT> ws, int cur, int shift) {
  // a block walks the tiles tile = blockIdx.x, blockIdx.x + gridDim.x, ... of its pair
  for (int tile = blockIdx.x;; tile += gridDim.x) {
    const int n = ws.n_side[2 * blockIdx.y] + ws.n_side[2 * blockIdx.y + 1];
    if (tile * kTile >= n) return;
    radix_scatter_kernel_tile<KeyT>(ws, cur, shift, tile);
    __syncthreads();
  }
}

// tmax[pair] = the largest key carried by a right record = the key of the last right record of the
// sorted array; tmax_has[pair] = 0 if there is no right record.  One warp per pair, scanning backwards; it normally stops at once.
template <typename KeyT>
__global__ void __launch_bounds__(32)
global_tmax_kernel(const SortWs<KeyT> ws, int cur) {
  const int pair = blockIdx.x, lane = threadIdx.x;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const KeyT* keys = (cur ? ws.keys[1] : ws.keys[0]) + (size_t)pair * ws.rec_stride;
  const uint32_t* vals = (cur ? ws.vals[1] : ws.vals[0]) + (size_t)pair * ws.rec_stride;
  for (int hi = n; hi > 0; hi -= 32) {
    const int i = hi - 1 - lane;
    const bool right = i >= 0 && (vals[i] & kSideBit) != 0u;
    const uint32_t b = __ballot_sync(0xffffffffu, right);
    if (b) {
      if (lane == __ffs(b) - 1) { ws.tmax[pair] = (unsigned long long)keys[i]; ws.tmax_has[pair] = 1u; }
      return;
    }
  }
  if (lane == 0) { ws.tmax[pair] = 0ull; ws.tmax_has[pair] = 0u; }
}

// ---- segmented scan over the sorted records ---------------------------------------------------------
struct GlobalEmitArgs {
  int32_t W;
  int32_t disp_high, vertical_tolerance;
  int32_t mode;                      // 0 supports (filtered), 1 correspondences (unfiltered), 2 index pairs
  void* out;                         // strided: pair p writes at out + p * out_stride records, at most cap of them
  long long out_stride, cap;         // packed (pair_base != nullptr): pair p writes from out + pair_base[p], cap = total capacity
  const long long* pair_base;        // [n_pairs + 1] exclusive prefix of n_out (filled between the count and the emit step)
  int32_t* n_out;                    // [n_pairs]
};

// Matches among the kRounds consecutive sorted records i0 .. i0 + kRounds - 1 (i0 a multiple of kRounds): bit j of
// the result is set when record i0 + j is the left record of a match; V[j] / V[j + 1] are then its value and
// its partner's.  The window (one record before, three after) is loaded with 16-byte loads where it can be.
template <typename KeyT>
__device__ __forceinline__ uint32_t window_matches(const KeyT* __restrict__ keys, const uint32_t* __restrict__ vals, int i0, int n,
                                                   unsigned long long tmax, bool has_tmax, const GlobalEmitArgs& a, uint32_t (&V)[kRounds + 1]) {
  if (i0 >= n) return 0u;
  KeyT K[kRounds + 4];                                   // K[j] = key of record i0 - 1 + j
  const bool full = i0 + kRounds <= n && (reinterpret_cast<uintptr_t>(keys + i0) & 15u) == 0u &&
                    (reinterpret_cast<uintptr_t>(vals + i0) & 15u) == 0u;
  if (full) {
    KeyT kk[kRounds];
    load_keys16(keys + i0, kk);
    uint32_t vv[kRounds];
    load_keys16(vals + i0, vv);
#pragma unroll
    for (int j = 0; j < kRounds; j++) { K[j + 1] = kk[j]; V[j] = vv[j]; }
  } else {
#pragma unroll
    for (int j = 0; j < kRounds; j++) { const int i = min(i0 + j, n - 1); K[j + 1] = keys[i]; V[j] = vals[i]; }
  }
  K[0] = keys[max(i0 - 1, 0)];
#pragma unroll
  for (int j = kRounds; j < kRounds + 3; j++) K[j + 1] = keys[min(i0 + j, n - 1)];
  V[kRounds] = vals[min(i0 + kRounds, n - 1)];
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < kRounds; j++) {
    const int i = i0 + j;
    const KeyT k = K[j + 1];
    bool m = i < n && !(V[j] & kSideBit);                                  // a left record ...
    m = m && !(i > 0 && K[j] == k);                                        // ... the first of its run ...
    m = m && i + 1 < n && K[j + 2] == k && (V[j + 1] & kSideBit);          // ... followed by a right record
    const bool is_tail = has_tmax && ((unsigned long long)k == tmax);
    const bool k2 = i + 2 < n && K[j + 3] == k, k3 = i + 3 < n && K[j + 4] == k;
    m = m && (is_tail ? (k2 && !k3) : !k2);              // tail key: exactly {L, R, R}; elsewhere exactly {L, R}
    if (m && a.mode == 0) {
      const uint32_t vl = V[j], vr = V[j + 1] & ~kSideBit;
      const int xl = (int)(vl % (uint32_t)a.W), yl = (int)(vl / (uint32_t)a.W);
      const int xr = (int)(vr % (uint32_t)a.W), yr = (int)(vr / (uint32_t)a.W);
      const int dx = xl - xr, dy = yl - yr;
      m = dy <= a.vertical_tolerance && -dy <= a.vertical_tolerance && dx <= a.disp_high && -dx <= a.disp_high;
    }
    bits |= (m ? 1u : 0u) << j;
  }
  return bits;
}

template <typename KeyT>
__device__ __forceinline__ void global_count_kernel_tile(const SortWs<KeyT> ws, int cur, const GlobalEmitArgs a, const int tile) {
  __shared__ int cnt;
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kTile - 1) / kTile;
  if (tile >= nb) return;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  uint32_t V[kRounds + 1];
  const uint32_t bits = window_matches((cur ? ws.keys[1] : ws.keys[0]) + (size_t)pair * ws.rec_stride,
                                       (cur ? ws.vals[1] : ws.vals[0]) + (size_t)pair * ws.rec_stride,
                                       tile * kTile + kRounds * threadIdx.x, n, ws.tmax[pair], ws.tmax_has[pair] != 0u, a, V);
  const int mine = __reduce_add_sync(0xffffffffu, __popc(bits));
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) ws.blockcount[(size_t)pair * (ws.nb_max + 1) + tile] = cnt;
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
global_count_kernel(const SortWs<KeyT> ws, int cur, const GlobalEmitArgs a) {
  // a block walks the tiles tile = blockIdx.x, blockIdx.x + gridDim.x, ... of its pair
  for (int tile = blockIdx.x;; tile += gridDim.x) {
    const int n = ws.n_side[2 * blockIdx.y] + ws.n_side[2 * blockIdx.y + 1];
    if (tile * kTile >= n) return;
    global_count_kernel_tile<KeyT>(ws, cur, a, tile);
    __syncthreads();
  }
}

// exclusive scan of a pair's block counts (one warp per pair), total -> n_out[pair]
template <typename KeyT>
__global__ void __launch_bounds__(32)
global_blockscan_kernel(const SortWs<KeyT> ws, const GlobalEmitArgs a) {
  const int pair = blockIdx.x, lane = threadIdx.x;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kTile - 1) / kTile;
  int32_t* bc = ws.blockcount + (size_t)pair * (ws.nb_max + 1);
  int carry = 0;
  for (int b0 = 0; b0 < nb; b0 += 32) {
    const int b = b0 + lane;
    const int c = (b < nb) ? bc[b] : 0;
    int incl = c;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, k); if (lane >= k) incl += t; }
    if (b < nb) bc[b] = carry + incl - c;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) a.n_out[pair] = carry;
}

// packed output: pair_base[p + 1] = pair_base[p] + n_out[p]; `first` starts the prefix at 0, otherwise it continues
// from pair_base[0] (left there by the previous chunk of pairs)
__global__ void global_pairbase_kernel(const int32_t* __restrict__ n_out, int n_pairs, long long* __restrict__ pair_base, int first) {
  if (threadIdx.x == 0) {
    long long acc = first ? 0ll : pair_base[0];
    for (int p = 0; p < n_pairs; p++) { pair_base[p] = acc; acc += n_out[p]; }
    pair_base[n_pairs] = acc;
  }
}

template <typename KeyT>
__device__ __forceinline__ void global_emit_kernel_tile(const SortWs<KeyT> ws, int cur, const GlobalEmitArgs a, const int tile) {
  __shared__ int warp_base[kSortWarps];
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kTile - 1) / kTile;
  if (tile >= nb) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t V[kRounds + 1];
  const uint32_t bits = window_matches((cur ? ws.keys[1] : ws.keys[0]) + (size_t)pair * ws.rec_stride,
                                       (cur ? ws.vals[1] : ws.vals[0]) + (size_t)pair * ws.rec_stride,
                                       tile * kTile + kRounds * tid, n, ws.tmax[pair], ws.tmax_has[pair] != 0u, a, V);
  // thread order = record order: exclusive scan of the per-thread match counts
  const int c = __popc(bits);
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) warp_base[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int v = lane < kSortWarps ? warp_base[lane] : 0;
    int wi = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
    if (lane < kSortWarps) warp_base[lane] = wi - v;
  }
  __syncthreads();
  long long k = (long long)ws.blockcount[(size_t)pair * (ws.nb_max + 1) + tile] + warp_base[wid] + (incl - c);
#pragma unroll
  for (int j = 0; j < kRounds; j++) {
    if (!((bits >> j) & 1u)) continue;
    const long long slot = k++;
    const long long idx = a.pair_base ? a.pair_base[pair] + slot : (long long)pair * a.out_stride + slot;
    if ((a.pair_base ? idx : slot) >= a.cap) continue;
    const uint32_t vl = V[j], vr = V[j + 1] & ~kSideBit;
    if (a.mode == 2) {
      int32_t* o = reinterpret_cast<int32_t*>(a.out) + 2 * idx;
      o[0] = (int32_t)vl; o[1] = (int32_t)vr;
    } else {
      const int xl = (int)(vl % (uint32_t)a.W), yl = (int)(vl / (uint32_t)a.W);
      const int xr = (int)(vr % (uint32_t)a.W), yr = (int)(vr / (uint32_t)a.W);
      if (a.mode == 0) {
        float* o = reinterpret_cast<float*>(a.out) + 3 * idx;
        o[0] = __int_as_float(xl); o[1] = __int_as_float(yl); o[2] = (float)(xl - xr);
      } else {
        int32_t* o = reinterpret_cast<int32_t*>(a.out) + 4 * idx;
        o[0] = xl; o[1] = yl; o[2] = xr; o[3] = yr;
      }
    }
  }
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
global_emit_kernel(const SortWs<KeyT> ws, int cur, const GlobalEmitArgs a) {
  // a block walks the tiles tile = blockIdx.x, blockIdx.x + gridDim.x, ... of its pair
  for (int tile = blockIdx.x;; tile += gridDim.x) {
    const int n = ws.n_side[2 * blockIdx.y] + ws.n_side[2 * blockIdx.y + 1];
    if (tile * kTile >= n) return;
    global_emit_kernel_tile<KeyT>(ws, cur, a, tile);
    __syncthreads();
  }
}

// ---- useHashtable(true): replay of the reference's buckets -------------------------------------------
struct HtArgs {
  const uint32_t* hash;              // hash images of the chunk's first pair, or
  const unsigned long long* keys64;  // explicit keys (src then tar); vals then index this list
  int32_t W, H, epipolar, n_src;
};

__device__ __forceinline__ void write_match(const GlobalEmitArgs& a, int pair, long long slot, uint32_t vl, uint32_t vr) {
  const long long idx = a.pair_base ? a.pair_base[pair] + slot : (long long)pair * a.out_stride + slot;
  if ((a.pair_base ? idx : slot) >= a.cap) return;
  if (a.mode == 2) {
    int32_t* o = reinterpret_cast<int32_t*>(a.out) + 2 * idx;
    o[0] = (int32_t)vl; o[1] = (int32_t)vr;
    return;
  }
  const int xl = (int)(vl % (uint32_t)a.W), yl = (int)(vl / (uint32_t)a.W);
  const int xr = (int)(vr % (uint32_t)a.W), yr = (int)(vr / (uint32_t)a.W);
  if (a.mode == 0) {
    float* o = reinterpret_cast<float*>(a.out) + 3 * idx;
    o[0] = __int_as_float(xl); o[1] = __int_as_float(yl); o[2] = (float)(xl - xr);
  } else {
    int32_t* o = reinterpret_cast<int32_t*>(a.out) + 4 * idx;
    o[0] = xl; o[1] = yl; o[2] = xr; o[3] = yr;
  }
}

__device__ __forceinline__ bool passes_filter(const GlobalEmitArgs& a, uint32_t vl, uint32_t vr) {
  if (a.mode != 0) return true;
  const int xl = (int)(vl % (uint32_t)a.W), yl = (int)(vl / (uint32_t)a.W);
  const int xr = (int)(vr % (uint32_t)a.W), yr = (int)(vr / (uint32_t)a.W);
  const int dx = xl - xr, dy = yl - yr;
  return dy <= a.vertical_tolerance && -dy <= a.vertical_tolerance && dx <= a.disp_high && -dx <= a.disp_high;
}

// The 64-bit key of every bucket-sorted record, gathered once (one thread per record) into the idle half of the
// ping-pong key buffer, so that the bucket replay below reads consecutive memory.
__device__ __forceinline__ void ht_keys_kernel_tile(const SortWs<uint32_t> ws, int cur, const HtArgs h, const int tile) {
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const uint32_t* vals = (cur ? ws.vals[1] : ws.vals[0]) + (size_t)pair * ws.rec_stride;
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(cur ? ws.keys[0] : ws.keys[1]) + (size_t)pair * ws.rec_stride;
  const uint32_t* img = h.keys64 ? nullptr : h.hash + (size_t)(2 * pair) * h.H * h.W;
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const int i = tile * kTile + r * kSortThreads + threadIdx.x;
    if (i >= n) continue;
    const uint32_t v = vals[i], pix = v & ~kSideBit;
    unsigned long long k;
    if (h.keys64) {
      k = h.keys64[(v >> 31) ? h.n_src + pix : pix];
    } else {
      k = img[(size_t)(v >> 31) * h.H * h.W + pix] & 0x7fffffffu;
      if (h.epipolar) k |= (unsigned long long)(pix / (uint32_t)h.W) << 32;
    }
    skey[i] = k;
  }
}

__global__ void __launch_bounds__(kSortThreads)
ht_keys_kernel(const SortWs<uint32_t> ws, int cur, const HtArgs h) {
  for (int tile = blockIdx.x;; tile += gridDim.x) {
    const int n = ws.n_side[2 * blockIdx.y] + ws.n_side[2 * blockIdx.y + 1];
    if (tile * kTile >= n) return;
    ht_keys_kernel_tile(ws, cur, h, tile);
    __syncthreads();
  }
}

// Record i of the bucket-sorted array: if it opens a bucket, replay that bucket and return its number of
// matches (written from `slot` on when kWrite).  keys = bucket indices, skey = full keys, vals = side << 31 | index.
template <bool kWrite>
__device__ int ht_bucket(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                         const unsigned long long* __restrict__ skey, int i, int n, int pair, const GlobalEmitArgs& a, long long slot) {
  if (i >= n) return 0;
  const uint32_t b = keys[i];
  if (i > 0 && keys[i - 1] == b) return 0;
  unsigned long long lk[kHtDepth];                       // the bucket's list: keys ascending, ties in insertion order
  uint32_t lv[kHtDepth];
  int m = 0;
  for (int j = i; j < n && m < kHtDepth && keys[j] == b; j++) {        // hashmatch.hpp:101: a full bucket drops the rest
    const uint32_t v = vals[j];
    const unsigned long long k = skey[j];
    int pos = m;                                         // :112-116: behind every element with key <= k
    while (pos > 0 && lk[pos - 1] > k) { lk[pos] = lk[pos - 1]; lv[pos] = lv[pos - 1]; pos--; }
    lk[pos] = k; lv[pos] = v;
    m++;
  }
  int found = 0, j = 0;
  while (j < m) {                                        // getDuplicates, hashmatch.hpp:162-198
    const int p = j;
    j = j + 1;
    if (j < m && lk[p] == lk[j]) {
      if ((lv[p] ^ lv[j]) & kSideBit) {                  // from different images
        bool emit, stop = false;
        if (j + 1 < m) { emit = lk[j + 1] != lk[j]; stop = j + 2 >= m; }     // :174-180
        else emit = true;                                                     // :181-183
        if (emit) {
          const uint32_t vl = lv[p] & ~kSideBit, vr = lv[j] & ~kSideBit;      // left records precede equal right ones
          if (passes_filter(a, vl, vr)) {
            if (kWrite) write_match(a, pair, slot + found, vl, vr);
            found++;
          }
        }
        if (stop) break;
      } else if (j + 1 < m && ((lv[j] ^ lv[j + 1]) & kSideBit)) {
        j = j + 1;                                       // :189-193: step over the false pair
      }
    }
  }
  return found;
}

// Matches per record (non-zero only where a bucket opens) into the idle half of the value buffer, and per tile.
__device__ __forceinline__ void ht_count_kernel_tile(const SortWs<uint32_t> ws, int cur, const GlobalEmitArgs a, const int tile) {
  __shared__ int cnt;
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kTile - 1) / kTile;
  if (tile >= nb) return;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const uint32_t* keys = (cur ? ws.keys[1] : ws.keys[0]) + (size_t)pair * ws.rec_stride;
  const uint32_t* vals = (cur ? ws.vals[1] : ws.vals[0]) + (size_t)pair * ws.rec_stride;
  const unsigned long long* skey = reinterpret_cast<const unsigned long long*>(cur ? ws.keys[0] : ws.keys[1]) + (size_t)pair * ws.rec_stride;
  uint8_t* found = reinterpret_cast<uint8_t*>((cur ? ws.vals[0] : ws.vals[1]) + (size_t)pair * ws.rec_stride);
  int mine = 0;
#pragma unroll 1
  for (int r = 0; r < kRounds; r++) {                    // round-major: neighbouring lanes replay neighbouring buckets
    const int i = tile * kTile + r * kSortThreads + threadIdx.x;
    const int c = ht_bucket<false>(keys, vals, skey, i, n, pair, a, 0);
    if (i < n) found[i] = (uint8_t)c;
    mine += c;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) ws.blockcount[(size_t)pair * (ws.nb_max + 1) + tile] = cnt;
}

__global__ void __launch_bounds__(kSortThreads)
ht_count_kernel(const SortWs<uint32_t> ws, int cur, const GlobalEmitArgs a) {
  for (int tile = blockIdx.x;; tile += gridDim.x) {
    const int n = ws.n_side[2 * blockIdx.y] + ws.n_side[2 * blockIdx.y + 1];
    if (tile * kTile >= n) return;
    ht_count_kernel_tile(ws, cur, a, tile);
    __syncthreads();
  }
}

__device__ __forceinline__ void ht_emit_kernel_tile(const SortWs<uint32_t> ws, int cur, const GlobalEmitArgs a, const int tile) {
  __shared__ int warp_base[kSortWarps];
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kTile - 1) / kTile;
  if (tile >= nb) return;
  const uint32_t* keys = (cur ? ws.keys[1] : ws.keys[0]) + (size_t)pair * ws.rec_stride;
  const uint32_t* vals = (cur ? ws.vals[1] : ws.vals[0]) + (size_t)pair * ws.rec_stride;
  const unsigned long long* skey = reinterpret_cast<const unsigned long long*>(cur ? ws.keys[0] : ws.keys[1]) + (size_t)pair * ws.rec_stride;
  const uint8_t* found = reinterpret_cast<const uint8_t*>((cur ? ws.vals[0] : ws.vals[1]) + (size_t)pair * ws.rec_stride);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int i0 = tile * kTile + kRounds * tid;     // thread order = record order = bucket order
  int cr[kRounds], c = 0;
#pragma unroll
  for (int r = 0; r < kRounds; r++) { cr[r] = (i0 + r < n) ? found[i0 + r] : 0; c += cr[r]; }
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) warp_base[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int v = lane < kSortWarps ? warp_base[lane] : 0;
    int wi = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
    if (lane < kSortWarps) warp_base[lane] = wi - v;
  }
  __syncthreads();
  long long slot = (long long)ws.blockcount[(size_t)pair * (ws.nb_max + 1) + tile] + warp_base[wid] + (incl - c);
#pragma unroll 1
  for (int r = 0; r < kRounds; r++)
    if (cr[r]) slot += ht_bucket<true>(keys, vals, skey, i0 + r, n, pair, a, slot);      // only buckets that hold a match
}

__global__ void __launch_bounds__(kSortThreads)
ht_emit_kernel(const SortWs<uint32_t> ws, int cur, const GlobalEmitArgs a) {
  for (int tile = blockIdx.x;; tile += gridDim.x) {
    const int n = ws.n_side[2 * blockIdx.y] + ws.n_side[2 * blockIdx.y + 1];
    if (tile * kTile >= n) return;
    ht_emit_kernel_tile(ws, cur, a, tile);
    __syncthreads();
  }
}

// ---- host-side launch sequences -----------------------------------------------------------------------
static size_t pad256(size_t b) { return (b + 255) / 256 * 256; }

// Blocks per pair of the tile kernels: the number of tiles is only known on the device (and small after the
// pre-filter), so a block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; enough blocks to fill the GPU
// for several waves when few pairs share a launch (the hashtable replay is latency bound and wants them), no more
// than that when many do.
static int tile_grid_x(int nb_capacity, int n_pairs) { return std::max(1, std::min(nb_capacity, std::max(48, 2400 / std::max(n_pairs, 1)))); }

// bytes of workspace for n_pairs pairs of up to max_records records each, images of W x H (0 x 0 for explicit keys)
size_t global_workspace_bytes(long long max_records, int n_pairs, int H, int W) {
  const size_t nb = (size_t)((max_records + kTile - 1) / kTile + 1);
  const size_t np = (size_t)n_pairs, rows = (size_t)std::max(H, 1), kw = (size_t)(W + 31) / 32;
  size_t b = pad256(np * 8) + pad256(np * 4) + pad256(np * 2 * 4) + 2 * pad256(np * (size_t)max_records * 8) + 2 * pad256(np * (size_t)max_records * 4) +
             pad256(np * kDigits * nb * 4) + pad256(np * kDigits * 4) + pad256(np * (nb + 1) * 4) + pad256(np * 2 * rows * 4) + 256;
  if (W > 0) b += pad256(np * 2 * rows * 4) + pad256(np * rows * 8) + pad256(np * 2 * (size_t)kBloomWords * 4) + pad256(np * 2 * rows * kw * 4);
  return b;
}

template <typename KeyT>
static SortWs<KeyT> carve(void* ws, long long max_records, int n_pairs, int H, int W = 0) {
  SortWs<KeyT> w;
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  auto take = [&p](size_t bytes) { uint8_t* r = p; p += pad256(bytes); return r; };
  const size_t np = (size_t)n_pairs;
  w.nb_max = (int)((max_records + kTile - 1) / kTile + 1);
  w.rec_stride = max_records;
  w.tmax = reinterpret_cast<unsigned long long*>(take(np * 8));
  w.tmax_has = reinterpret_cast<uint32_t*>(take(np * 4));
  w.n_side = reinterpret_cast<int32_t*>(take(np * 2 * 4));
  w.keys[0] = reinterpret_cast<KeyT*>(take(np * (size_t)max_records * 8));        // sized for 64-bit keys either way
  w.keys[1] = reinterpret_cast<KeyT*>(take(np * (size_t)max_records * 8));
  w.vals[0] = reinterpret_cast<uint32_t*>(take(np * (size_t)max_records * 4));
  w.vals[1] = reinterpret_cast<uint32_t*>(take(np * (size_t)max_records * 4));
  w.blockhist = reinterpret_cast<uint32_t*>(take(np * kDigits * (size_t)w.nb_max * 4));
  w.digit_tot = reinterpret_cast<uint32_t*>(take(np * kDigits * 4));
  w.blockcount = reinterpret_cast<int32_t*>(take(np * ((size_t)w.nb_max + 1) * 4));
  w.rowoff = reinterpret_cast<int32_t*>(take(np * 2 * (size_t)std::max(H, 1) * 4));
  w.rowcnt_f = nullptr; w.rowmax = nullptr; w.bloom = nullptr; w.keep = nullptr;
  w.keep_words = (W + 31) / 32;
  if (W > 0) {                                           // pre-filter buffers (hash-image input)
    const size_t rows = (size_t)std::max(H, 1);
    w.rowcnt_f = reinterpret_cast<int32_t*>(take(np * 2 * rows * 4));
    w.rowmax = reinterpret_cast<unsigned long long*>(take(np * rows * 8));
    w.bloom = reinterpret_cast<uint32_t*>(take(np * 2 * (size_t)kBloomWords * 4));
    w.keep = reinterpret_cast<uint32_t*>(take(np * 2 * rows * (size_t)w.keep_words * 4));
  }
  return w;
}

// The scatter kernel's dynamic shared memory: set once per device, when a context is created.
cudaError_t configure_match_global() {
  cudaError_t e = cudaFuncSetAttribute(radix_scatter_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_smem_bytes<uint32_t>());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(radix_scatter_kernel<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_smem_bytes<unsigned long long>());
  return e;
}

template <typename KeyT>
static cudaError_t sort_passes(SortWs<KeyT>& w, long long max_records, int n_pairs, int key_bits, cudaStream_t stream, int* launches,
                               int* cur_out) {
  const int nb = (int)((max_records + kTile - 1) / kTile);
  const dim3 grid(tile_grid_x(nb, n_pairs), n_pairs);
  int cur = 0;
  for (int shift = 0; shift < key_bits; shift += 8) {
    radix_hist_kernel<KeyT><<<grid, kSortThreads, 0, stream>>>(w, cur, shift);
    radix_scan_kernel<KeyT><<<dim3(kDigits / 32, n_pairs), 1024, 0, stream>>>(w);
    radix_scatter_kernel<KeyT><<<grid, kSortThreads, scatter_smem_bytes<KeyT>(), stream>>>(w, cur, shift);
    cur ^= 1;
    *launches += 3;
  }
  *cur_out = cur;
  return cudaGetLastError();
}

template <typename KeyT>
static cudaError_t sort_and_emit(SortWs<KeyT>& w, long long max_records, int n_pairs, int key_bits, const GlobalEmitArgs& ea,
                                 cudaStream_t stream, int* launches, int first_chunk = 1, bool tmax_known = false) {
  const int nb = (int)((max_records + kTile - 1) / kTile);
  if (nb <= 0 || n_pairs <= 0) return cudaSuccess;
  const dim3 grid(tile_grid_x(nb, n_pairs), n_pairs);
  int cur = 0;
  {
    cudaError_t e = sort_passes(w, max_records, n_pairs, key_bits, stream, launches, &cur);
    if (e != cudaSuccess) return e;
  }
  if (!tmax_known) global_tmax_kernel<KeyT><<<n_pairs, 32, 0, stream>>>(w, cur);      // else: taken over all right records
  global_count_kernel<KeyT><<<grid, kSortThreads, 0, stream>>>(w, cur, ea);
  global_blockscan_kernel<KeyT><<<n_pairs, 32, 0, stream>>>(w, ea);
  if (ea.pair_base) { global_pairbase_kernel<<<1, 32, 0, stream>>>(ea.n_out, n_pairs, const_cast<long long*>(ea.pair_base), first_chunk); *launches += 1; }
  global_emit_kernel<KeyT><<<grid, kSortThreads, 0, stream>>>(w, cur, ea);
  *launches += 4;
  return cudaGetLastError();
}

// Hash images of n_pairs pairs ([2 * n_pairs][H][W], rowcnt [2 * n_pairs][H]) -> ordered supports (mode 0)
// or correspondences (mode 1); pair p writes at out + p * out_stride records (pair_base == nullptr) or packed
// from out + pair_base[p] (pair_base filled here; first_chunk = 0 continues the prefix of an earlier call), count to
// n_out[p], candidate counts to n_cand[2p], n_cand[2p+1] (optional).  ws from global_workspace_bytes(max_records, n_pairs, H, W).
cudaError_t launch_match_global(const uint32_t* hash, const int32_t* rowcnt, int W, int H, int n_pairs, int epipolar, int key_bits,
                                int disp_high, int vertical_tolerance, int mode, void* ws, long long max_records, void* out,
                                long long out_stride, long long cap, int32_t* n_out, int32_t* n_cand, cudaStream_t stream,
                                int* launches, int hashtable, long long* pair_base, int first_chunk) {
  GlobalEmitArgs ea{};
  ea.W = W; ea.disp_high = disp_high; ea.vertical_tolerance = vertical_tolerance; ea.mode = mode;
  ea.out = out; ea.out_stride = out_stride; ea.cap = cap; ea.n_out = n_out; ea.pair_base = pair_base;
  const dim3 gather_grid((2 * H * 32 + 127) / 128, n_pairs);
  cudaError_t e;
  if (hashtable) {                                       // inference.hpp:204-225: sort by bucket, then replay the buckets
    SortWs<uint32_t> w = carve<uint32_t>(ws, max_records, n_pairs, H, W);
    global_rowoff_kernel<<<dim3(2, n_pairs), 1024, 0, stream>>>(rowcnt, H, w.rowoff, w.n_side);
    global_gather_kernel<uint32_t><<<gather_grid, 128, 0, stream>>>(hash, W, H, epipolar, 1, 0, w);
    *launches += 2;
    const int nb = (int)((max_records + kTile - 1) / kTile);
    if (nb <= 0 || n_pairs <= 0) return cudaGetLastError();
    int cur = 0;
    if ((e = sort_passes(w, max_records, n_pairs, 18, stream, launches, &cur)) != cudaSuccess) return e;   // 214673 < 2^18
    HtArgs h{hash, nullptr, W, H, epipolar, 0};
    const dim3 grid(tile_grid_x(nb, n_pairs), n_pairs);
    ht_keys_kernel<<<grid, kSortThreads, 0, stream>>>(w, cur, h);
    ht_count_kernel<<<grid, kSortThreads, 0, stream>>>(w, cur, ea);
    global_blockscan_kernel<uint32_t><<<n_pairs, 32, 0, stream>>>(w, ea);
    if (ea.pair_base) { global_pairbase_kernel<<<1, 32, 0, stream>>>(n_out, n_pairs, pair_base, first_chunk); *launches += 1; }
    ht_emit_kernel<<<grid, kSortThreads, 0, stream>>>(w, cur, ea);
    *launches += 4;
    e = cudaGetLastError();
    if (e == cudaSuccess && n_cand) e = cudaMemcpyAsync(n_cand, w.n_side, (size_t)n_pairs * 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
    return e;
  }
  // sort-path semantics: pre-filter, gather the surviving records, sort, scan
  static const bool prefilter = !(std::getenv("GPC_GLOBAL_PREFILTER") && std::atoi(std::getenv("GPC_GLOBAL_PREFILTER")) == 0);
  SortWs<uint32_t> w32 = carve<uint32_t>(ws, max_records, n_pairs, H, W);
  const int32_t* counts = rowcnt;
  if (prefilter) {
    if ((e = cudaMemsetAsync(w32.bloom, 0, (size_t)n_pairs * 2 * kBloomWords * sizeof(uint32_t), stream)) != cudaSuccess) return e;
    if (n_cand) {                                        // the reported candidate counts are the unfiltered ones
      global_rowoff_kernel<<<dim3(2, n_pairs), 1024, 0, stream>>>(rowcnt, H, w32.rowoff, n_cand);
      *launches += 1;
    }
    bloom_build_kernel<<<gather_grid, 128, 0, stream>>>(hash, W, H, epipolar, w32);
    rowmax_reduce_kernel<<<n_pairs, 32, 0, stream>>>(w32, H);
    bloom_filter_kernel<<<gather_grid, 128, 0, stream>>>(hash, W, H, epipolar, w32);
    *launches += 3;
    counts = w32.rowcnt_f;
  }
  if (epipolar) {
    SortWs<unsigned long long> w = carve<unsigned long long>(ws, max_records, n_pairs, H, W);
    global_rowoff_kernel<<<dim3(2, n_pairs), 1024, 0, stream>>>(counts, H, w.rowoff, w.n_side);
    global_gather_kernel<unsigned long long><<<gather_grid, 128, 0, stream>>>(hash, W, H, 1, 0, prefilter ? 1 : 0, w);
    *launches += 2;
    int hb = 1; while ((1 << hb) < H) hb++;
    e = sort_and_emit(w, max_records, n_pairs, 32 + hb, ea, stream, launches, first_chunk, prefilter);
    if (e == cudaSuccess && n_cand && !prefilter) e = cudaMemcpyAsync(n_cand, w.n_side, (size_t)n_pairs * 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
    return e;
  }
  global_rowoff_kernel<<<dim3(2, n_pairs), 1024, 0, stream>>>(counts, H, w32.rowoff, w32.n_side);
  global_gather_kernel<uint32_t><<<gather_grid, 128, 0, stream>>>(hash, W, H, 0, 0, prefilter ? 1 : 0, w32);
  *launches += 2;
  e = sort_and_emit(w32, max_records, n_pairs, key_bits, ea, stream, launches, first_chunk, prefilter);
  if (e == cudaSuccess && n_cand && !prefilter) e = cudaMemcpyAsync(n_cand, w32.n_side, (size_t)n_pairs * 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
  return e;
}

// Explicit key lists: the caller has copied n_src + n_tar 64-bit keys (src then tar) into global_key_buffer().
__global__ void keys_prepare_kernel(SortWs<unsigned long long> ws, int ns, int nt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { ws.n_side[0] = ns; ws.n_side[1] = nt; }
  if (i >= ns + nt) return;
  const bool tar = i >= ns;
  ws.vals[0][i] = tar ? (kSideBit | (uint32_t)(i - ns)) : (uint32_t)i;
}

cudaError_t launch_match_keys(void* ws, long long max_records, int ns, int nt, int key_bits, int32_t* out_pairs, long long cap,
                              int32_t* n_out, cudaStream_t stream, int* launches) {
  SortWs<unsigned long long> w = carve<unsigned long long>(ws, max_records, 1, 0);
  const int n = ns + nt;
  keys_prepare_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, ns, nt);
  *launches += 1;
  GlobalEmitArgs ea{};
  ea.W = 1; ea.mode = 2; ea.out = out_pairs; ea.out_stride = 0; ea.cap = cap; ea.n_out = n_out;
  return sort_and_emit(w, n, 1, key_bits, ea, stream, launches);
}

// ndb::Hashmatch on explicit keys: d_keys64 holds n_src + n_tar keys (src then tar), outside the workspace.
__global__ void ht_keys_prepare_kernel(SortWs<uint32_t> ws, const unsigned long long* __restrict__ keys64, int ns, int nt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { ws.n_side[0] = ns; ws.n_side[1] = nt; }
  if (i >= ns + nt) return;
  ws.keys[0][i] = (uint32_t)(keys64[i] % kHtBuckets);
  ws.vals[0][i] = i >= ns ? (kSideBit | (uint32_t)(i - ns)) : (uint32_t)i;
}

cudaError_t launch_hashmatch_keys(void* ws, long long max_records, const unsigned long long* d_keys64, int ns, int nt, int32_t* out_pairs,
                                  long long cap, int32_t* n_out, cudaStream_t stream, int* launches) {
  SortWs<uint32_t> w = carve<uint32_t>(ws, max_records, 1, 0);
  const int n = ns + nt;
  ht_keys_prepare_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, d_keys64, ns, nt);
  *launches += 1;
  GlobalEmitArgs ea{};
  ea.W = 1; ea.mode = 2; ea.out = out_pairs; ea.out_stride = 0; ea.cap = cap; ea.n_out = n_out;
  const int nb = (int)((n + kTile - 1) / kTile);
  int cur = 0;
  cudaError_t e = sort_passes(w, n, 1, 18, stream, launches, &cur);
  if (e != cudaSuccess) return e;
  HtArgs h{nullptr, d_keys64, 1, 1, 0, ns};
  ht_keys_kernel<<<dim3(nb, 1), kSortThreads, 0, stream>>>(w, cur, h);
  ht_count_kernel<<<dim3(nb, 1), kSortThreads, 0, stream>>>(w, cur, ea);
  global_blockscan_kernel<uint32_t><<<1, 32, 0, stream>>>(w, ea);
  ht_emit_kernel<<<dim3(nb, 1), kSortThreads, 0, stream>>>(w, cur, ea);
  *launches += 4;
  return cudaGetLastError();
}

void* global_key_buffer(void* ws, long long max_records) {
  return carve<unsigned long long>(ws, max_records, 1, 0).keys[0];
}

// ------------------------------------------------------------------------------------------------------------------
// Wide states (forests of more than 32 tests, gpc_capi.cu: run_wide_pair).  A candidate's state is a tuple of 32-bit
// words, one hash plane per group of 32 tests.  Two planes are folded into one by replacing the word pair (hi, lo) of
// every candidate of BOTH images of the pair by its dense rank among all pairs that occur: equal tuples get equal
// ranks, and the rank order is the lexicographic order (hi, lo) -- an exact, order-preserving compression, so after
// folding all planes the ordinary matchers run on a plain 31-bit "state" again.
//   wide_gather : every candidate pixel appends key = hi << 32 | lo and its pixel id (image << 30.. see below)
//   sort        : the 64-bit LSD radix sort above
//   wide_rank   : run starts counted per tile, scanned, and rank | candidate flag scattered to the output planes
// Pixel id = image * P + y * W + x (two images, P <= 2^30).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
wide_gather_kernel(const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo, long long n_pix, SortWs<unsigned long long> ws) {
  // order inside the key array is irrelevant (it is sorted next): one warp-aggregated reservation per warp
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t l = 0, h = 0;
  if (i < n_pix) { l = lo[i]; h = hi ? hi[i] : 0u; }
  const bool c = (l >> 31) != 0u;
  const uint32_t b = __ballot_sync(0xffffffffu, c);
  if (b == 0u) return;
  int base = 0;
  if (lane == __ffs(b) - 1) base = atomicAdd(&ws.n_side[0], __popc(b));
  base = __shfl_sync(0xffffffffu, base, __ffs(b) - 1);
  if (c) {
    const int p = base + __popc(b & ((1u << lane) - 1u));
    ws.keys[0][p] = ((unsigned long long)(h & 0x7fffffffu) << 32) | (unsigned long long)(l & 0x7fffffffu);
    ws.vals[0][p] = (uint32_t)i;
  }
}

constexpr int kWidePer = kTile / 256;            // consecutive sorted keys per thread: one CTA of 256 threads per sort tile

__global__ void __launch_bounds__(256)
wide_count_kernel(const SortWs<unsigned long long> ws, int cur) {          // run starts per tile of kTile sorted keys
  const int n = ws.n_side[0];
  const unsigned long long* keys = cur ? ws.keys[1] : ws.keys[0];
  const int i0 = blockIdx.x * kTile + threadIdx.x * kWidePer;
  int c = 0;
  for (int k = 0; k < kWidePer; k++) {
    const int i = i0 + k;
    c += (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
  }
  __shared__ int tot;
  if (threadIdx.x == 0) tot = 0;
  __syncthreads();
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&tot, c);
  __syncthreads();
  if (threadIdx.x == 0) ws.blockcount[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024)
wide_scan_kernel(const SortWs<unsigned long long> ws, int n_tiles) {       // exclusive prefix of the tile counts, in place
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n_tiles; i0 += 1024) {
    const int i = i0 + tid;
    const int v = (i < n_tiles) ? ws.blockcount[i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int w = warp_sums[lane];
      int wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    const int base = carry + warp_sums[wid];
    if (i < n_tiles) ws.blockcount[i] = base + incl - v;
    __syncthreads();
    if (tid == 1023) carry = base + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
wide_rank_kernel(const SortWs<unsigned long long> ws, int cur, uint32_t* __restrict__ out) {
  __shared__ int warp_sums[8];
  const int n = ws.n_side[0];
  const unsigned long long* keys = cur ? ws.keys[1] : ws.keys[0];
  const uint32_t* vals = cur ? ws.vals[1] : ws.vals[0];
  const int i0 = blockIdx.x * kTile + threadIdx.x * kWidePer, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int c = 0;
  for (int k = 0; k < kWidePer; k++) {
    const int i = i0 + k;
    c += (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
  }
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  int run = ws.blockcount[blockIdx.x] + incl - c;                          // run starts before this thread's first key
  for (int w = 0; w < wid; w++) run += warp_sums[w];
  for (int k = 0; k < kWidePer; k++) {
    const int i = i0 + k;
    if (i >= n) break;
    run += (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
    out[vals[i]] = kCandFlag | (uint32_t)(run - 1);                        // rank of the run this key belongs to
  }
}

size_t wide_workspace_bytes(long long max_records) { return global_workspace_bytes(max_records, 1, 0, 0); }

// out[pixel] = flag | dense rank of (hi[pixel], lo[pixel]) among the candidate pixels (flag = bit 31 of lo), 0 elsewhere
// is the caller's job (out is expected to be zeroed).  hi == nullptr ranks lo alone.  key_bits: significant bits of the
// 64-bit key hi << 32 | lo.
cudaError_t launch_wide_rank(void* ws, long long max_records, const uint32_t* hi, const uint32_t* lo, long long n_pix, int key_bits,
                             uint32_t* out, cudaStream_t stream, int* launches) {
  SortWs<unsigned long long> w = carve<unsigned long long>(ws, max_records, 1, 0);
  cudaError_t e = cudaMemsetAsync(w.n_side, 0, 2 * sizeof(int32_t), stream);
  if (e != cudaSuccess) return e;
  wide_gather_kernel<<<(unsigned)((n_pix + 255) / 256), 256, 0, stream>>>(hi, lo, n_pix, w);
  int cur = 0;
  e = sort_passes(w, max_records, 1, key_bits, stream, launches, &cur);
  if (e != cudaSuccess) return e;
  const int n_tiles = (int)((max_records + kTile - 1) / kTile);
  wide_count_kernel<<<n_tiles, 256, 0, stream>>>(w, cur);
  wide_scan_kernel<<<1, 1024, 0, stream>>>(w, n_tiles);
  wide_rank_kernel<<<n_tiles, 256, 0, stream>>>(w, cur, out);
  *launches += 4;
  return cudaGetLastError();
}

}  // namespace gpc

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
// Every kernel carries a pair dimension (blockIdx.y): a chunk of independent pairs is sorted by the
// same launches, each pair in its own slice of the workspace.
#include <algorithm>

#include "gpc_device.cuh"

namespace gpc {

constexpr int kSortThreads = 1024;          // one key per thread, 32 warps
constexpr int kDigits = 256;
constexpr uint32_t kSideBit = 0x80000000u;

// Workspace of a chunk of pairs (device pointers; slice `pair` starts at pair * stride of each array).
template <typename KeyT>
struct SortWs {
  KeyT* keys[2];                 // [n_pairs][rec_stride] ping-pong
  uint32_t* vals[2];             // [n_pairs][rec_stride] side << 31 | index
  uint32_t* blockhist;           // [n_pairs][kDigits][nb_max]
  uint32_t* digit_tot;           // [n_pairs][kDigits]
  int32_t* blockcount;           // [n_pairs][nb_max + 1]
  int32_t* n_side;               // [n_pairs][2] left / right record counts
  unsigned long long* tmax;      // [n_pairs] largest right key + 1 (0: no right record)
  int32_t* rowoff;               // [n_pairs][2][H] candidate offsets (hash-image input only)
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

// One warp per (side, row): candidates in raster order.
template <typename KeyT>
__global__ void __launch_bounds__(128)
global_gather_kernel(const uint32_t* __restrict__ hash, int W, int H, int epipolar, const SortWs<KeyT> ws) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, pair = blockIdx.y;
  if (warp >= 2 * H) return;
  const int side = warp / H, y = warp - side * H;
  const uint32_t* row = hash + ((size_t)(2 * pair + side) * H + y) * W;
  const int32_t* n_side = ws.n_side + 2 * pair;
  KeyT* keys = ws.keys[0] + (size_t)pair * ws.rec_stride;
  uint32_t* vals = ws.vals[0] + (size_t)pair * ws.rec_stride;
  int off = ws.rowoff[((size_t)2 * pair + side) * H + y] + (side ? n_side[0] : 0);
  unsigned long long kmax = 0;
  bool any = false;
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const uint32_t v = (x < W) ? row[x] : 0u;
    const bool c = (v >> 31) != 0u;
    const uint32_t b = __ballot_sync(0xffffffffu, c);
    if (c) {
      const int p = off + __popc(b & ((1u << lane) - 1u));
      unsigned long long k = v & 0x7fffffffu;
      if (epipolar) k |= (unsigned long long)y << 32;
      keys[p] = (KeyT)k;
      vals[p] = (side ? kSideBit : 0u) | (uint32_t)(y * W + x);
      kmax = k > kmax ? k : kmax;
      any = true;
    }
    off += __popc(b);
  }
  if (side == 1 && any) atomicMax(ws.tmax + pair, kmax + 1ull);     // stored as key+1 so that 0 means "no right record"
}

// ---- LSD radix sort, 8-bit digits, one key per thread ---------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const SortWs<KeyT> ws, int cur, int shift) {
  __shared__ uint32_t hist[kDigits];
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  if ((int)blockIdx.x >= nb) return;
  const KeyT* keys = ws.keys[cur] + (size_t)pair * ws.rec_stride;
  if (threadIdx.x < kDigits) hist[threadIdx.x] = 0u;
  __syncthreads();
  const int i = blockIdx.x * kSortThreads + threadIdx.x;
  if (i < n) {
    const uint32_t d = (uint32_t)(keys[i] >> shift) & 0xffu;
    const uint32_t old = atomicAdd(&hist[d], 1u);
    if (old == 0xffffffffu) __trap();                    // value-returning form (see match_rows.cu note)
  }
  __syncthreads();
  if (threadIdx.x < kDigits)
    ws.blockhist[((size_t)pair * kDigits + threadIdx.x) * ws.nb_max + blockIdx.x] = hist[threadIdx.x];
}

// per-digit exclusive scan over the nb active blocks (one warp per digit, coalesced 32-wide
// chunks); digit_tot[d] = number of keys with digit d.  The cross-digit base is added by the
// scatter kernel.  Grid (kDigits / 32, n_pairs).
template <typename KeyT>
__global__ void __launch_bounds__(1024)
radix_scan_kernel(const SortWs<KeyT> ws) {
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
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

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const SortWs<KeyT> ws, int cur, int shift) {
  __shared__ uint32_t whist[32][kDigits];                // per-warp digit counts -> exclusive prefix over warps
  __shared__ uint32_t dbase[kDigits];                    // exclusive prefix of the digit totals
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  if ((int)blockIdx.x >= nb) return;
  const KeyT* keys_in = ws.keys[cur] + (size_t)pair * ws.rec_stride;
  const uint32_t* vals_in = ws.vals[cur] + (size_t)pair * ws.rec_stride;
  KeyT* keys_out = ws.keys[cur ^ 1] + (size_t)pair * ws.rec_stride;
  uint32_t* vals_out = ws.vals[cur ^ 1] + (size_t)pair * ws.rec_stride;
  const uint32_t* digit_tot = ws.digit_tot + (size_t)pair * kDigits;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int k = tid; k < 32 * kDigits; k += kSortThreads) (&whist[0][0])[k] = 0u;
  __syncthreads();
  const int i = blockIdx.x * kSortThreads + tid;
  const bool active = i < n;
  KeyT key = 0;
  uint32_t val = 0, d = 0, rank = 0;
  const uint32_t amask = __ballot_sync(0xffffffffu, active);
  if (active) {
    key = keys_in[i]; val = vals_in[i];
    d = (uint32_t)(key >> shift) & 0xffu;
    const uint32_t peers = __match_any_sync(amask, d);
    rank = __popc(peers & ((1u << lane) - 1u));
    if (rank == 0) whist[wid][d] = __popc(peers);
  }
  __syncthreads();
  if (tid < kDigits) {
    uint32_t acc = 0;
    for (int w = 0; w < 32; w++) { const uint32_t c = whist[w][tid]; whist[w][tid] = acc; acc += c; }
  } else if (wid == 8) {                                 // one warp: exclusive scan of the 256 digit totals
    uint32_t c[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { c[k] = digit_tot[8 * lane + k]; sum += c[k]; }
    uint32_t incl = sum;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, k); if (lane >= k) incl += t; }
    uint32_t run = incl - sum;
#pragma unroll
    for (int k = 0; k < 8; k++) { dbase[8 * lane + k] = run; run += c[k]; }
  }
  __syncthreads();
  if (active) {
    const uint32_t pos = dbase[d] + ws.blockhist[((size_t)pair * kDigits + d) * ws.nb_max + blockIdx.x] + whist[wid][d] + rank;
    keys_out[pos] = key;
    vals_out[pos] = val;
  }
}

// ---- segmented scan over the sorted records ---------------------------------------------------------
struct GlobalEmitArgs {
  int32_t W;
  int32_t disp_high, vertical_tolerance;
  int32_t mode;                      // 0 supports (filtered), 1 correspondences (unfiltered), 2 index pairs
  void* out;                         // pair p writes at out + p * out_stride records
  long long out_stride, cap;         // records per pair region, capacity of one region
  int32_t* n_out;                    // [n_pairs]
};

template <typename KeyT>
__device__ __forceinline__ bool is_match(const KeyT* __restrict__ keys, const uint32_t* __restrict__ vals, int i, int n,
                                         unsigned long long tmax1, const GlobalEmitArgs& a, uint32_t* vl, uint32_t* vr) {
  if (i >= n) return false;
  const uint32_t v = vals[i];
  if (v & kSideBit) return false;
  const KeyT k = keys[i];
  if (i > 0 && keys[i - 1] == k) return false;
  if (i + 1 >= n || keys[i + 1] != k || !(vals[i + 1] & kSideBit)) return false;
  const bool is_tail = (tmax1 != 0ull) && ((unsigned long long)k == tmax1 - 1ull);
  if (!is_tail) {
    if (i + 2 < n && keys[i + 2] == k) return false;
  } else {
    if (i + 2 >= n || keys[i + 2] != k) return false;     // a single right record at the tail never matches
    if (i + 3 < n && keys[i + 3] == k) return false;      // three or more: duplicates
  }
  *vl = v; *vr = vals[i + 1] & ~kSideBit;
  if (a.mode == 0) {
    const int xl = (int)(*vl % (uint32_t)a.W), yl = (int)(*vl / (uint32_t)a.W);
    const int xr = (int)(*vr % (uint32_t)a.W), yr = (int)(*vr / (uint32_t)a.W);
    const int dx = xl - xr, dy = yl - yr;
    if (!(dy <= a.vertical_tolerance && -dy <= a.vertical_tolerance && dx <= a.disp_high && -dx <= a.disp_high)) return false;
  }
  return true;
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
global_count_kernel(const SortWs<KeyT> ws, int cur, const GlobalEmitArgs a) {
  __shared__ int cnt;
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  if ((int)blockIdx.x >= nb) return;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  uint32_t vl, vr;
  const bool m = is_match(ws.keys[cur] + (size_t)pair * ws.rec_stride, ws.vals[cur] + (size_t)pair * ws.rec_stride,
                          blockIdx.x * kSortThreads + threadIdx.x, n, ws.tmax[pair], a, &vl, &vr);
  const uint32_t b = __ballot_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && b) { if (atomicAdd(&cnt, __popc(b)) < 0) __trap(); }
  __syncthreads();
  if (threadIdx.x == 0) ws.blockcount[(size_t)pair * (ws.nb_max + 1) + blockIdx.x] = cnt;
}

// exclusive scan of a pair's block counts (one warp per pair), total -> n_out[pair]
template <typename KeyT>
__global__ void __launch_bounds__(32)
global_blockscan_kernel(const SortWs<KeyT> ws, const GlobalEmitArgs a) {
  const int pair = blockIdx.x, lane = threadIdx.x;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
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

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
global_emit_kernel(const SortWs<KeyT> ws, int cur, const GlobalEmitArgs a) {
  __shared__ int warp_base[32];
  const int pair = blockIdx.y;
  const int n = ws.n_side[2 * pair] + ws.n_side[2 * pair + 1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  if ((int)blockIdx.x >= nb) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t vl = 0, vr = 0;
  const bool m = is_match(ws.keys[cur] + (size_t)pair * ws.rec_stride, ws.vals[cur] + (size_t)pair * ws.rec_stride,
                          blockIdx.x * kSortThreads + tid, n, ws.tmax[pair], a, &vl, &vr);
  const uint32_t b = __ballot_sync(0xffffffffu, m);
  if (lane == 0) warp_base[wid] = __popc(b);
  __syncthreads();
  if (wid == 0) {
    int v = warp_base[lane], incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    warp_base[lane] = incl - v;
  }
  __syncthreads();
  if (!m) return;
  const long long k = (long long)ws.blockcount[(size_t)pair * (ws.nb_max + 1) + blockIdx.x] + warp_base[wid] + __popc(b & ((1u << lane) - 1u));
  if (k >= a.cap) return;
  const long long idx = (long long)pair * a.out_stride + k;
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

// ---- host-side launch sequences -----------------------------------------------------------------------
static size_t pad256(size_t b) { return (b + 255) / 256 * 256; }

// bytes of workspace for n_pairs pairs of up to max_records records each, H rows (0 for explicit keys)
size_t global_workspace_bytes(long long max_records, int n_pairs, int H) {
  const size_t nb = (size_t)((max_records + kSortThreads - 1) / kSortThreads + 1);
  const size_t np = (size_t)n_pairs;
  return pad256(np * 8) + pad256(np * 2 * 4) + 2 * pad256(np * (size_t)max_records * 8) + 2 * pad256(np * (size_t)max_records * 4) +
         pad256(np * kDigits * nb * 4) + pad256(np * kDigits * 4) + pad256(np * (nb + 1) * 4) + pad256(np * 2 * (size_t)std::max(H, 1) * 4) + 256;
}

template <typename KeyT>
static SortWs<KeyT> carve(void* ws, long long max_records, int n_pairs, int H) {
  SortWs<KeyT> w;
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  auto take = [&p](size_t bytes) { uint8_t* r = p; p += pad256(bytes); return r; };
  const size_t np = (size_t)n_pairs;
  w.nb_max = (int)((max_records + kSortThreads - 1) / kSortThreads + 1);
  w.rec_stride = max_records;
  w.tmax = reinterpret_cast<unsigned long long*>(take(np * 8));
  w.n_side = reinterpret_cast<int32_t*>(take(np * 2 * 4));
  w.keys[0] = reinterpret_cast<KeyT*>(take(np * (size_t)max_records * 8));        // sized for 64-bit keys either way
  w.keys[1] = reinterpret_cast<KeyT*>(take(np * (size_t)max_records * 8));
  w.vals[0] = reinterpret_cast<uint32_t*>(take(np * (size_t)max_records * 4));
  w.vals[1] = reinterpret_cast<uint32_t*>(take(np * (size_t)max_records * 4));
  w.blockhist = reinterpret_cast<uint32_t*>(take(np * kDigits * (size_t)w.nb_max * 4));
  w.digit_tot = reinterpret_cast<uint32_t*>(take(np * kDigits * 4));
  w.blockcount = reinterpret_cast<int32_t*>(take(np * ((size_t)w.nb_max + 1) * 4));
  w.rowoff = reinterpret_cast<int32_t*>(take(np * 2 * (size_t)std::max(H, 1) * 4));
  return w;
}

template <typename KeyT>
static cudaError_t sort_and_emit(SortWs<KeyT>& w, long long max_records, int n_pairs, int key_bits, const GlobalEmitArgs& ea,
                                 cudaStream_t stream, int* launches) {
  const int nb = (int)((max_records + kSortThreads - 1) / kSortThreads);
  if (nb <= 0 || n_pairs <= 0) return cudaSuccess;
  const dim3 grid(nb, n_pairs);
  int cur = 0;
  for (int shift = 0; shift < key_bits; shift += 8) {
    radix_hist_kernel<KeyT><<<grid, kSortThreads, 0, stream>>>(w, cur, shift);
    radix_scan_kernel<KeyT><<<dim3(kDigits / 32, n_pairs), 1024, 0, stream>>>(w);
    radix_scatter_kernel<KeyT><<<grid, kSortThreads, 0, stream>>>(w, cur, shift);
    cur ^= 1;
    *launches += 3;
  }
  global_count_kernel<KeyT><<<grid, kSortThreads, 0, stream>>>(w, cur, ea);
  global_blockscan_kernel<KeyT><<<n_pairs, 32, 0, stream>>>(w, ea);
  global_emit_kernel<KeyT><<<grid, kSortThreads, 0, stream>>>(w, cur, ea);
  *launches += 3;
  return cudaGetLastError();
}

// Hash images of n_pairs pairs ([2 * n_pairs][H][W], rowcnt [2 * n_pairs][H]) -> ordered supports (mode 0)
// or correspondences (mode 1); pair p writes at out + p * out_stride records, count to n_out[p], candidate
// counts to n_cand[2p], n_cand[2p+1] (optional).  ws from global_workspace_bytes(max_records, n_pairs, H).
cudaError_t launch_match_global(const uint32_t* hash, const int32_t* rowcnt, int W, int H, int n_pairs, int epipolar, int key_bits,
                                int disp_high, int vertical_tolerance, int mode, void* ws, long long max_records, void* out,
                                long long out_stride, long long cap, int32_t* n_out, int32_t* n_cand, cudaStream_t stream,
                                int* launches) {
  GlobalEmitArgs ea{};
  ea.W = W; ea.disp_high = disp_high; ea.vertical_tolerance = vertical_tolerance; ea.mode = mode;
  ea.out = out; ea.out_stride = out_stride; ea.cap = cap; ea.n_out = n_out;
  const dim3 gather_grid((2 * H * 32 + 127) / 128, n_pairs);
  cudaError_t e;
  if (epipolar) {
    SortWs<unsigned long long> w = carve<unsigned long long>(ws, max_records, n_pairs, H);
    if ((e = cudaMemsetAsync(w.tmax, 0, (size_t)n_pairs * 8, stream)) != cudaSuccess) return e;
    global_rowoff_kernel<<<dim3(2, n_pairs), 1024, 0, stream>>>(rowcnt, H, w.rowoff, w.n_side);
    global_gather_kernel<unsigned long long><<<gather_grid, 128, 0, stream>>>(hash, W, H, 1, w);
    *launches += 2;
    int hb = 1; while ((1 << hb) < H) hb++;
    e = sort_and_emit(w, max_records, n_pairs, 32 + hb, ea, stream, launches);
    if (e == cudaSuccess && n_cand) e = cudaMemcpyAsync(n_cand, w.n_side, (size_t)n_pairs * 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
    return e;
  }
  SortWs<uint32_t> w = carve<uint32_t>(ws, max_records, n_pairs, H);
  if ((e = cudaMemsetAsync(w.tmax, 0, (size_t)n_pairs * 8, stream)) != cudaSuccess) return e;
  global_rowoff_kernel<<<dim3(2, n_pairs), 1024, 0, stream>>>(rowcnt, H, w.rowoff, w.n_side);
  global_gather_kernel<uint32_t><<<gather_grid, 128, 0, stream>>>(hash, W, H, 0, w);
  *launches += 2;
  e = sort_and_emit(w, max_records, n_pairs, key_bits, ea, stream, launches);
  if (e == cudaSuccess && n_cand) e = cudaMemcpyAsync(n_cand, w.n_side, (size_t)n_pairs * 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
  return e;
}

// Explicit key lists: the caller has copied n_src + n_tar 64-bit keys (src then tar) into global_key_buffer().
__global__ void keys_prepare_kernel(SortWs<unsigned long long> ws, int ns, int nt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { ws.n_side[0] = ns; ws.n_side[1] = nt; }
  if (i >= ns + nt) return;
  const bool tar = i >= ns;
  ws.vals[0][i] = tar ? (kSideBit | (uint32_t)(i - ns)) : (uint32_t)i;
  if (tar) atomicMax(ws.tmax, ws.keys[0][i] + 1ull);
}

cudaError_t launch_match_keys(void* ws, long long max_records, int ns, int nt, int key_bits, int32_t* out_pairs, long long cap,
                              int32_t* n_out, cudaStream_t stream, int* launches) {
  SortWs<unsigned long long> w = carve<unsigned long long>(ws, max_records, 1, 0);
  cudaError_t e = cudaMemsetAsync(w.tmax, 0, 8, stream);
  if (e != cudaSuccess) return e;
  const int n = ns + nt;
  keys_prepare_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, ns, nt);
  *launches += 1;
  GlobalEmitArgs ea{};
  ea.W = 1; ea.mode = 2; ea.out = out_pairs; ea.out_stride = 0; ea.cap = cap; ea.n_out = n_out;
  return sort_and_emit(w, n, 1, key_bits, ea, stream, launches);
}

void* global_key_buffer(void* ws, long long max_records) {
  return carve<unsigned long long>(ws, max_records, 1, 0).keys[0];
}

}  // namespace gpc

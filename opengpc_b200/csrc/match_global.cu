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
#include "gpc_device.cuh"

namespace gpc {

constexpr int kSortThreads = 1024;          // one key per thread, 32 warps
constexpr int kDigits = 256;
constexpr uint32_t kSideBit = 0x80000000u;

// ---- record gathering from hash images ----------------------------------------------------------
// One warp per (side, row): candidates in raster order.  rowoff[side][y] = exclusive prefix of the
// row's candidate count inside its image; n_side[0..1] = totals.
__global__ void __launch_bounds__(1024)
global_rowoff_kernel(const int32_t* __restrict__ rowcnt, int H, int32_t* __restrict__ rowoff, int32_t* __restrict__ n_side) {
  // blockIdx.x = side; single block scan over H rows
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int side = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int y0 = 0; y0 < H; y0 += 1024) {
    const int y = y0 + tid;
    const int v = (y < H) ? rowcnt[(size_t)side * H + y] : 0;
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
    if (y < H) rowoff[(size_t)side * H + y] = base + incl - v;
    __syncthreads();
    if (tid == 1023) carry = base + incl;
    __syncthreads();
  }
  if (tid == 0) n_side[side] = carry;
}

template <typename KeyT>
__global__ void __launch_bounds__(128)
global_gather_kernel(const uint32_t* __restrict__ hash_l, const uint32_t* __restrict__ hash_r, const int32_t* __restrict__ rowoff,
                     const int32_t* __restrict__ n_side, int W, int H, int epipolar, KeyT* __restrict__ keys,
                     uint32_t* __restrict__ vals, unsigned long long* __restrict__ tmax) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * H) return;
  const int side = warp / H, y = warp - side * H;
  const uint32_t* row = (side ? hash_r : hash_l) + (size_t)y * W;
  int off = rowoff[(size_t)side * H + y] + (side ? n_side[0] : 0);
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
  if (side == 1 && any) atomicMax(tmax, kmax + 1ull);     // stored as key+1 so that 0 means "no right record"
}

// ---- LSD radix sort, 8-bit digits, one key per thread ---------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const KeyT* __restrict__ keys, const int32_t* __restrict__ n_ptr, int shift, int nb_max,
                  uint32_t* __restrict__ blockhist) {
  __shared__ uint32_t hist[kDigits];
  const int n = n_ptr[0] + n_ptr[1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  if ((int)blockIdx.x >= nb) return;
  if (threadIdx.x < kDigits) hist[threadIdx.x] = 0u;
  __syncthreads();
  const int i = blockIdx.x * kSortThreads + threadIdx.x;
  if (i < n) {
    const uint32_t d = (uint32_t)(keys[i] >> shift) & 0xffu;
    const uint32_t old = atomicAdd(&hist[d], 1u);
    if (old == 0xffffffffu) __trap();                    // value-returning form (see match_rows.cu note)
  }
  __syncthreads();
  if (threadIdx.x < kDigits) blockhist[(size_t)threadIdx.x * nb_max + blockIdx.x] = hist[threadIdx.x];
}

// per-digit exclusive scan over the nb active blocks (one warp per digit, coalesced 32-wide
// chunks); digit_tot[d] = number of keys with digit d.  The cross-digit base is added by the
// scatter kernel.
__global__ void __launch_bounds__(1024)
radix_scan_kernel(const int32_t* __restrict__ n_ptr, int nb_max, uint32_t* __restrict__ blockhist,
                  uint32_t* __restrict__ digit_tot) {
  const int n = n_ptr[0] + n_ptr[1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  const int d = blockIdx.x * 32 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  uint32_t* row = blockhist + (size_t)d * nb_max;
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
  if (lane == 0) digit_tot[d] = carry;
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out, const int32_t* __restrict__ n_ptr, int shift, int nb_max,
                     const uint32_t* __restrict__ blockhist, const uint32_t* __restrict__ digit_tot) {
  __shared__ uint32_t whist[32][kDigits];                // per-warp digit counts -> exclusive prefix over warps
  __shared__ uint32_t dbase[kDigits];                    // exclusive prefix of the digit totals
  const int n = n_ptr[0] + n_ptr[1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  if ((int)blockIdx.x >= nb) return;
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
    const uint32_t pos = dbase[d] + blockhist[(size_t)d * nb_max + blockIdx.x] + whist[wid][d] + rank;
    keys_out[pos] = key;
    vals_out[pos] = val;
  }
}

// ---- segmented scan over the sorted records ---------------------------------------------------------
struct GlobalEmitArgs {
  const uint32_t* vals;
  const int32_t* n_ptr;
  const unsigned long long* tmax;    // largest right key + 1 (0: no right record)
  int32_t* blockcount;               // [nb_max + 1]
  int32_t W;
  int32_t disp_high, vertical_tolerance;
  int32_t mode;                      // 0 supports (filtered), 1 correspondences (unfiltered), 2 index pairs
  void* out;
  long long cap;
  int32_t* n_out;
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
global_count_kernel(const KeyT* __restrict__ keys, const GlobalEmitArgs a) {
  __shared__ int cnt;
  const int n = a.n_ptr[0] + a.n_ptr[1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  if ((int)blockIdx.x >= nb) return;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  uint32_t vl, vr;
  const bool m = is_match(keys, a.vals, blockIdx.x * kSortThreads + threadIdx.x, n, *a.tmax, a, &vl, &vr);
  const uint32_t b = __ballot_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && b) { if (atomicAdd(&cnt, __popc(b)) < 0) __trap(); }
  __syncthreads();
  if (threadIdx.x == 0) a.blockcount[blockIdx.x] = cnt;
}

__global__ void global_blockscan_kernel(const GlobalEmitArgs a) {
  if (threadIdx.x != 0) return;
  const int n = a.n_ptr[0] + a.n_ptr[1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  int acc = 0;
  for (int b = 0; b < nb; b++) { const int c = a.blockcount[b]; a.blockcount[b] = acc; acc += c; }
  *a.n_out = acc;
}

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
global_emit_kernel(const KeyT* __restrict__ keys, const GlobalEmitArgs a) {
  __shared__ int warp_base[32];
  const int n = a.n_ptr[0] + a.n_ptr[1];
  const int nb = (n + kSortThreads - 1) / kSortThreads;
  if ((int)blockIdx.x >= nb) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t vl = 0, vr = 0;
  const bool m = is_match(keys, a.vals, blockIdx.x * kSortThreads + tid, n, *a.tmax, a, &vl, &vr);
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
  const long long idx = (long long)a.blockcount[blockIdx.x] + warp_base[wid] + __popc(b & ((1u << lane) - 1u));
  if (idx >= a.cap) return;
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
size_t global_workspace_bytes(long long max_records, int key_bytes) {
  const long long nb = (max_records + kSortThreads - 1) / kSortThreads + 1;
  size_t b = 0;
  b += 2 * (size_t)max_records * key_bytes;       // key ping-pong
  b += 2 * (size_t)max_records * 4;               // value ping-pong
  b += (size_t)kDigits * nb * 4 + kDigits * 4;    // block histograms + digit totals
  b += (size_t)(nb + 1) * 4;                      // block counts
  b += 64;                                        // n_side[2], tmax
  return b + 1024;
}

template <typename KeyT>
struct GlobalWs {
  KeyT* keys[2]; uint32_t* vals[2]; uint32_t* blockhist; uint32_t* digit_tot; int32_t* blockcount; int32_t* n_side;
  unsigned long long* tmax;
  int nb_max;
};

template <typename KeyT>
static GlobalWs<KeyT> carve(void* ws, long long max_records) {
  GlobalWs<KeyT> w;
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  auto take = [&p](size_t bytes) { uint8_t* r = p; p += (bytes + 255) / 256 * 256; return r; };
  w.nb_max = (int)((max_records + kSortThreads - 1) / kSortThreads + 1);
  w.tmax = reinterpret_cast<unsigned long long*>(take(8));
  w.n_side = reinterpret_cast<int32_t*>(take(8));
  w.keys[0] = reinterpret_cast<KeyT*>(take((size_t)max_records * sizeof(KeyT)));
  w.keys[1] = reinterpret_cast<KeyT*>(take((size_t)max_records * sizeof(KeyT)));
  w.vals[0] = reinterpret_cast<uint32_t*>(take((size_t)max_records * 4));
  w.vals[1] = reinterpret_cast<uint32_t*>(take((size_t)max_records * 4));
  w.blockhist = reinterpret_cast<uint32_t*>(take((size_t)kDigits * w.nb_max * 4));
  w.digit_tot = reinterpret_cast<uint32_t*>(take((size_t)kDigits * 4));
  w.blockcount = reinterpret_cast<int32_t*>(take((size_t)(w.nb_max + 1) * 4));
  return w;
}

size_t global_workspace_bytes_padded(long long max_records, int key_bytes) {
  return global_workspace_bytes(max_records, key_bytes) + 10 * 256;
}

template <typename KeyT>
static cudaError_t sort_and_emit(GlobalWs<KeyT>& w, long long max_records, int key_bits, GlobalEmitArgs ea, cudaStream_t stream,
                                 int* launches) {
  const int nb = (int)((max_records + kSortThreads - 1) / kSortThreads);
  if (nb <= 0) return cudaSuccess;
  int cur = 0;
  for (int shift = 0; shift < key_bits; shift += 8) {
    radix_hist_kernel<KeyT><<<nb, kSortThreads, 0, stream>>>(w.keys[cur], w.n_side, shift, w.nb_max, w.blockhist);
    radix_scan_kernel<<<kDigits / 32, 1024, 0, stream>>>(w.n_side, w.nb_max, w.blockhist, w.digit_tot);
    radix_scatter_kernel<KeyT><<<nb, kSortThreads, 0, stream>>>(w.keys[cur], w.vals[cur], w.keys[cur ^ 1], w.vals[cur ^ 1],
                                                               w.n_side, shift, w.nb_max, w.blockhist, w.digit_tot);
    cur ^= 1;
    *launches += 3;
  }
  ea.vals = w.vals[cur]; ea.n_ptr = w.n_side; ea.tmax = w.tmax; ea.blockcount = w.blockcount;
  global_count_kernel<KeyT><<<nb, kSortThreads, 0, stream>>>(w.keys[cur], ea);
  global_blockscan_kernel<<<1, 32, 0, stream>>>(ea);
  global_emit_kernel<KeyT><<<nb, kSortThreads, 0, stream>>>(w.keys[cur], ea);
  *launches += 3;
  return cudaGetLastError();
}

// Hash images (one pair) -> ordered supports (mode 0) or correspondences (mode 1).
// rowcnt = [2][H] candidate counts of the two images; ws from global_workspace_bytes_padded.
cudaError_t launch_match_global(const uint32_t* hash_l, const uint32_t* hash_r, const int32_t* rowcnt, int32_t* rowoff2,
                                int W, int H, int epipolar, int key_bits, int disp_high, int vertical_tolerance, int mode,
                                void* ws, long long max_records, void* out, long long cap, int32_t* n_out,
                                cudaStream_t stream, int* launches) {
  GlobalEmitArgs ea{};
  ea.W = W; ea.disp_high = disp_high; ea.vertical_tolerance = vertical_tolerance; ea.mode = mode;
  ea.out = out; ea.cap = cap; ea.n_out = n_out;
  const int gather_blocks = (2 * H * 32 + 127) / 128;
  cudaError_t e;
  if (epipolar) {
    GlobalWs<unsigned long long> w = carve<unsigned long long>(ws, max_records);
    if ((e = cudaMemsetAsync(w.tmax, 0, 8, stream)) != cudaSuccess) return e;
    global_rowoff_kernel<<<2, 1024, 0, stream>>>(rowcnt, H, rowoff2, w.n_side);
    global_gather_kernel<unsigned long long><<<gather_blocks, 128, 0, stream>>>(hash_l, hash_r, rowoff2, w.n_side, W, H, 1,
                                                                                  w.keys[0], w.vals[0], w.tmax);
    *launches += 2;
    int hb = 1; while ((1 << hb) < H) hb++;
    return sort_and_emit(w, max_records, 32 + hb, ea, stream, launches);
  }
  GlobalWs<uint32_t> w = carve<uint32_t>(ws, max_records);
  if ((e = cudaMemsetAsync(w.tmax, 0, 8, stream)) != cudaSuccess) return e;
  global_rowoff_kernel<<<2, 1024, 0, stream>>>(rowcnt, H, rowoff2, w.n_side);
  global_gather_kernel<uint32_t><<<gather_blocks, 128, 0, stream>>>(hash_l, hash_r, rowoff2, w.n_side, W, H, 0, w.keys[0],
                                                                      w.vals[0], w.tmax);
  *launches += 2;
  return sort_and_emit(w, max_records, key_bits, ea, stream, launches);
}

// Explicit key lists (device copies made by the caller into ws): keys laid out src then tar.
__global__ void keys_prepare_kernel(const unsigned long long* __restrict__ keys, int ns, int nt, uint32_t* __restrict__ vals,
                                    int32_t* __restrict__ n_side, unsigned long long* __restrict__ tmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { n_side[0] = ns; n_side[1] = nt; }
  if (i >= ns + nt) return;
  const bool tar = i >= ns;
  vals[i] = tar ? (kSideBit | (uint32_t)(i - ns)) : (uint32_t)i;
  if (tar) atomicMax(tmax, keys[i] + 1ull);
}

// keys_host_order: device pointer to ns + nt 64-bit keys already copied into the workspace's first key buffer.
cudaError_t launch_match_keys(void* ws, long long max_records, int ns, int nt, int key_bits, int32_t* out_pairs, long long cap,
                              int32_t* n_out, cudaStream_t stream, int* launches) {
  GlobalWs<unsigned long long> w = carve<unsigned long long>(ws, max_records);
  cudaError_t e = cudaMemsetAsync(w.tmax, 0, 8, stream);
  if (e != cudaSuccess) return e;
  const int n = ns + nt;
  keys_prepare_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w.keys[0], ns, nt, w.vals[0], w.n_side, w.tmax);
  *launches += 1;
  GlobalEmitArgs ea{};
  ea.W = 1; ea.mode = 2; ea.out = out_pairs; ea.cap = cap; ea.n_out = n_out;
  return sort_and_emit(w, n, key_bits, ea, stream, launches);
}

int32_t* global_nside_ptr(void* ws) { return carve<uint32_t>(ws, 1).n_side; }

void* global_key_buffer(void* ws, long long max_records) {
  return carve<unsigned long long>(ws, max_records).keys[0];
}

}  // namespace gpc

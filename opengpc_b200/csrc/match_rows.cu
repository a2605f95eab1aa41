// match_rows.cu -- kernel B (epipolar matching, one CTA per image row), the row-offset scan
// and kernel C (ordered emission of ndb::Support records).
//
// Replaces Forest::depthPriorFast's key build (inference.hpp:192-197), findCorrespondences
// (:227-254: two std::sort + merge scan) and rectifiedMatch's filter (:384-391).
//
// In epipolar mode the 64-bit key is (y << 32 | state), so equal keys imply equal rows and the
// global sort decomposes into H-26 independent row problems of <= 2(W-26) 31-bit states.  The
// reference's result is "states that occur exactly once in the left row and exactly once in
// the right row", ordered by (y, state), with two tail rules on the globally largest right key
// (SURVEY.md 8a row M).  A row CTA therefore never sorts its candidates: it builds an
// open-addressing table of the left states in shared memory with a store-then-verify protocol
// (no atomics on the table), probes it with the right states, and sorts only the surviving
// matches (about a tenth of the candidates) with a bitonic network before writing them out.
#include "gpc_device.cuh"

namespace gpc {

constexpr int kThreadsB = 256;
constexpr uint32_t kEmpty16 = 0xffffu;

__device__ __forceinline__ uint32_t slot_of(uint32_t key, int log2) {
  return (key * 0x9E3779B1u) >> (32 - log2);
}

struct RowSmem {
  uint32_t* key_l; uint32_t* key_r;        // [wcap]
  uint16_t* x_l; uint16_t* x_r;            // [wcap]
  uint16_t* cur_l; uint16_t* cur_r;        // [wcap] current / final slot per candidate
  uint16_t* tab_l; uint16_t* tab_r;        // [table] owner candidate per slot
  uint8_t* dup_l; uint8_t* dup_r;          // [table]
  unsigned long long* out;                 // [wcap]  state<<32 | xL<<16 | xR
};

size_t match_smem_bytes(int wcap, int table_log2) {
  size_t ts = (size_t)1 << table_log2;
  size_t pow2 = 1; while ((int)pow2 < wcap) pow2 <<= 1;
  return pow2 * 8 + (size_t)wcap * (4 + 4 + 2 + 2 + 2 + 2) + ts * (2 + 2 + 1 + 1) + 64;
}

// Compact the candidates (bit 31) of one hash row into key[] / x[]; order is irrelevant.
__device__ __forceinline__ void compact_row(const uint32_t* __restrict__ row, int W, uint32_t* key, uint16_t* xs,
                                            int* counter) {
  const int lane = threadIdx.x & 31;
  const int nquads = W / 4;                                           // W % 16 == 0
  for (int q0 = threadIdx.x & ~31; q0 < nquads; q0 += kThreadsB) {    // warp-uniform trip count (ballots below)
    const int q = q0 + lane;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q < nquads) v = __ldg(reinterpret_cast<const uint4*>(row) + q);
    uint32_t vv[4] = {v.x, v.y, v.z, v.w};
    uint32_t b[4];
    int total = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { b[k] = __ballot_sync(0xffffffffu, vv[k] >> 31); total += __popc(b[k]); }
    if (total == 0) continue;   // warp-uniform
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (vv[k] >> 31) {
        int p = base + __popc(b[k] & ((1u << lane) - 1u));
        key[p] = vv[k] & 0x7fffffffu;
        xs[p] = (uint16_t)(4 * q + k);
      }
      base += __popc(b[k]);
    }
  }
}

__global__ void __launch_bounds__(kThreadsB)
match_rows_kernel(const MatchArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int W = args.W, H = args.H, wcap = args.wcap, log2 = args.table_log2;
  const int ts = 1 << log2;
  const int tid = threadIdx.x;
  const int y = kRadius + blockIdx.x, pair = blockIdx.y;

  int pow2cap = 1; while (pow2cap < wcap) pow2cap <<= 1;
  RowSmem s;
  s.out = reinterpret_cast<unsigned long long*>(smem);
  s.key_l = reinterpret_cast<uint32_t*>(s.out + pow2cap);
  s.key_r = s.key_l + wcap;
  s.x_l = reinterpret_cast<uint16_t*>(s.key_r + wcap);
  s.x_r = s.x_l + wcap;
  s.cur_l = s.x_r + wcap;
  s.cur_r = s.cur_l + wcap;
  s.tab_l = s.cur_r + wcap;
  s.tab_r = s.tab_l + ts;
  s.dup_l = reinterpret_cast<uint8_t*>(s.tab_r + ts);
  s.dup_r = s.dup_l + ts;
  __shared__ int n_l, n_r, n_out, pending;
  __shared__ uint32_t kmax_s;
  __shared__ int cmax_s, xmin_s;

  if (tid == 0) { n_l = 0; n_r = 0; n_out = 0; kmax_s = 0; cmax_s = 0; xmin_s = 0x7fffffff; }
  for (int i = tid; i < ts; i += kThreadsB) { s.tab_l[i] = kEmpty16; s.tab_r[i] = kEmpty16; s.dup_l[i] = 0; s.dup_r[i] = 0; }
  __syncthreads();

  const uint32_t* row_l = args.hash + ((size_t)(2 * pair) * H + y) * W;
  const uint32_t* row_r = args.hash + ((size_t)(2 * pair + 1) * H + y) * W;
  compact_row(row_l, W, s.key_l, s.x_l, &n_l);
  compact_row(row_r, W, s.key_r, s.x_r, &n_r);
  __syncthreads();
  const int nl = n_l, nr = n_r;
  int m = 0;

  if (nl > 0 && nr > 0) {
    // ---- left table: store, barrier, verify; losers of a slot probe on ---------------------
    for (int i = tid; i < nl; i += kThreadsB) s.cur_l[i] = (uint16_t)slot_of(s.key_l[i], log2);
    for (;;) {
      for (int i = tid; i < nl; i += kThreadsB) {
        uint32_t c = s.cur_l[i];
        if (!(c & 0x8000u)) s.tab_l[c] = (uint16_t)i;           // bit 15 = settled
      }
      if (tid == 0) pending = 0;
      __syncthreads();
      bool mine = false;
      for (int i = tid; i < nl; i += kThreadsB) {
        uint32_t c = s.cur_l[i];
        if (c & 0x8000u) continue;
        const uint32_t key = s.key_l[i];
        uint32_t w = s.tab_l[c];
        if (w == (uint32_t)i) { s.cur_l[i] = (uint16_t)(c | 0x8000u); continue; }   // owns slot c
        for (;;) {                                              // table is read-only in this phase
          if (s.key_l[w] == key) { s.dup_l[c] = 1; c = 0xffffu; break; }            // duplicate of the owner
          c = (c + 1) & (ts - 1);
          w = s.tab_l[c];
          if (w == kEmpty16) break;
        }
        s.cur_l[i] = (uint16_t)c;                               // 0xffff = settled as a duplicate
        if (c != 0xffffu) mine = true;
      }
      if (mine) pending = 1;
      __syncthreads();
      if (!pending) break;
      __syncthreads();
    }
    // ---- right probes -------------------------------------------------------------------------
    for (int j = tid; j < nr; j += kThreadsB) {
      const uint32_t key = s.key_r[j];
      uint32_t c = slot_of(key, log2);
      for (;;) {
        uint32_t w = s.tab_l[c];
        if (w == kEmpty16) { c = 0xffffu; break; }
        if (s.key_l[w] == key) { s.tab_r[c] = (uint16_t)j; break; }
        c = (c + 1) & (ts - 1);
      }
      s.cur_r[j] = (uint16_t)c;
    }
    __syncthreads();
    for (int j = tid; j < nr; j += kThreadsB) {
      uint32_t c = s.cur_r[j];
      if (c != 0xffffu && s.tab_r[c] != (uint16_t)j) s.dup_r[c] = 1;   // another right pixel has this state
    }
    // ---- tail rules: only the globally last right key (largest row with right candidates) ---
    const bool last_row = (args.lastrow[2 * pair + 1] == y);
    if (last_row) {
      uint32_t km = 0;
      for (int j = tid; j < nr; j += kThreadsB) km = max(km, s.key_r[j]);
      km = __reduce_max_sync(0xffffffffu, km);
      if ((tid & 31) == 0) atomicMax(&kmax_s, km);
      __syncthreads();
      km = kmax_s;
      int c = 0, xm = 0x7fffffff;
      for (int j = tid; j < nr; j += kThreadsB)
        if (s.key_r[j] == km) { c++; xm = min(xm, (int)s.x_r[j]); }
      c = __reduce_add_sync(0xffffffffu, c);
      xm = __reduce_min_sync(0xffffffffu, xm);
      if ((tid & 31) == 0) { atomicAdd(&cmax_s, c); atomicMin(&xmin_s, xm); }
    }
    __syncthreads();
    // ---- emit: left states that own a slot, unique on both sides -------------------------------
    const uint32_t kmax = kmax_s;
    const int cmax = cmax_s, xmin = xmin_s;
    for (int i0 = 0; i0 < nl; i0 += kThreadsB) {
      const int i = i0 + tid;
      bool ok = false;
      unsigned long long rec = 0;
      if (i < nl) {
        uint32_t c = s.cur_l[i];
        if (c != 0xffffu) {
          c &= 0x7fffu;
          const uint32_t j = s.tab_r[c];
          if (!s.dup_l[c] && j != kEmpty16) {
            const uint32_t key = s.key_l[i];
            int xl = s.x_l[i], xr = s.x_r[j];
            ok = !s.dup_r[c];
            if (last_row && key == kmax) {       // inference.hpp:243-249 on the tail of sorted tar
              ok = (cmax == 2);                  // 1: the last element never matches; >=3: duplicates
              xr = xmin;                         // 2: "first of the two" := smaller x (stable order)
            }
            int dx = xl - xr;
            ok = ok && (dx <= args.disp_high && -dx <= args.disp_high) && (0 <= args.vertical_tolerance);
            rec = ((unsigned long long)key << 32) | ((unsigned long long)xl << 16) | (unsigned long long)xr;
          }
        }
      }
      const uint32_t b = __ballot_sync(0xffffffffu, ok);
      if (b) {
        int base = 0;
        if ((tid & 31) == 0) base = atomicAdd(&n_out, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (ok) s.out[base + __popc(b & ((1u << (tid & 31)) - 1u))] = rec;
      }
    }
    __syncthreads();
    m = n_out;
    // ---- order by state (keys are unique): bitonic network over the matches only --------------
    if (m > 1) {
      int p2 = 1; while (p2 < m) p2 <<= 1;
      for (int i = m + tid; i < p2; i += kThreadsB) s.out[i] = ~0ull;
      __syncthreads();
      for (int k = 2; k <= p2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < p2; i += kThreadsB) {
            int l = i ^ j;
            if (l > i) {
              unsigned long long a = s.out[i], b2 = s.out[l];
              bool up = ((i & k) == 0);
              if ((a > b2) == up) { s.out[i] = b2; s.out[l] = a; }
            }
          }
          __syncthreads();
        }
    }
  }
  // ---- stage the row's ordered matches ------------------------------------------------------------
  uint32_t* stage = args.stage + ((size_t)pair * H + y) * W;
  for (int i = tid; i < m; i += kThreadsB) stage[i] = (uint32_t)(s.out[i] & 0xffffffffull);
  if (tid == 0) args.rowmatch[(size_t)pair * H + y] = m;
}

cudaError_t configure_match_rows(int max_smem) {
  return cudaFuncSetAttribute(match_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

cudaError_t launch_match_rows(const MatchArgs& args, int n_pairs, cudaStream_t stream) {
  int rows = args.H - 2 * kRadius;
  if (rows <= 0 || n_pairs <= 0) return cudaSuccess;
  size_t smem = match_smem_bytes(args.wcap, args.table_log2);
  match_rows_kernel<<<dim3(rows, n_pairs), kThreadsB, smem, stream>>>(args);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Row-offset scan: one CTA per pair.  rowoff[pair][y] = exclusive prefix of rowmatch over rows,
// totals[pair] = number of supports, n_cand[pair][2] = candidate counts (sum of rowcnt).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
row_scan_kernel(const int32_t* __restrict__ rowmatch, const int32_t* __restrict__ rowcnt, int H,
                int32_t* __restrict__ rowoff, int32_t* __restrict__ totals, int32_t* __restrict__ n_cand) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int y0 = 0; y0 < H; y0 += 1024) {
    const int y = y0 + tid;
    int v = (y >= kRadius && y < H - kRadius) ? rowmatch[(size_t)pair * H + y] : 0;
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
    if (y < H) rowoff[(size_t)pair * H + y] = base + incl - v;
    __syncthreads();
    if (tid == 1023) carry = base + incl;
    __syncthreads();
  }
  if (tid == 0) totals[pair] = carry;
  if (n_cand != nullptr && wid < 2) {       // warp 0: left image, warp 1: right image
    int sum = 0;
    for (int y = lane; y < H; y += 32) sum += rowcnt[((size_t)(2 * pair + wid)) * H + y];
    sum = __reduce_add_sync(0xffffffffu, sum);
    if (lane == 0) n_cand[2 * pair + wid] = sum;
  }
}

cudaError_t launch_row_scan(const int32_t* rowmatch, const int32_t* rowcnt, int H, int n_pairs, int32_t* rowoff,
                            int32_t* totals, int32_t* n_cand, cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  row_scan_kernel<<<n_pairs, 1024, 0, stream>>>(rowmatch, rowcnt, H, rowoff, totals, n_cand);
  return cudaGetLastError();
}

// Exclusive prefix of the per-pair totals (packed output mode); single CTA, n_pairs is small.
__global__ void pair_scan_kernel(const int32_t* __restrict__ totals, int n_pairs, long long* __restrict__ pair_base) {
  if (threadIdx.x == 0) {
    long long acc = 0;
    for (int p = 0; p < n_pairs; p++) { pair_base[p] = acc; acc += totals[p]; }
    pair_base[n_pairs] = acc;
  }
}

cudaError_t launch_pair_scan(const int32_t* totals, int n_pairs, long long* pair_base, cudaStream_t stream) {
  pair_scan_kernel<<<1, 32, 0, stream>>>(totals, n_pairs, pair_base);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Kernel C: expand the staged rows into ndb::Support records at their final position.
// Output index of match k of row y of pair p = base(p) + rowoff[p][y] + k, i.e. ascending
// (y, state) -- the order std::sort gives the reference (inference.hpp:231).
//   packed mode  (pair_base != nullptr): base(p) = pair_base[p], limit = cap_total
//   strided mode (pair_base == nullptr): base(p) = p * cap_per_pair, limit = base + cap_per_pair
// ------------------------------------------------------------------------------------------------
struct EmitArgs {
  const uint32_t* stage; const int32_t* rowmatch; const int32_t* rowoff; const long long* pair_base;
  float* out;        // gpc_support = {int x, int y, float d}
  long long cap;     // cap_total (packed) or cap_per_pair (strided)
  int32_t W, H;
};

__global__ void __launch_bounds__(128)
emit_supports_kernel(const EmitArgs a) {
  const int y = kRadius + blockIdx.x, pair = blockIdx.y;
  const int m = a.rowmatch[(size_t)pair * a.H + y];
  if (m == 0) return;
  const long long off = a.rowoff[(size_t)pair * a.H + y];
  long long base, limit;
  if (a.pair_base) { base = a.pair_base[pair]; limit = a.cap; }
  else { base = (long long)pair * a.cap; limit = base + a.cap; }
  const uint32_t* stage = a.stage + ((size_t)pair * a.H + y) * a.W;
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    const long long idx = base + off + k;
    if (idx >= limit) break;
    const uint32_t u = stage[k];
    const int xl = (int)(u >> 16), xr = (int)(u & 0xffffu);
    float* o = a.out + 3 * idx;
    o[0] = __int_as_float(xl);
    o[1] = __int_as_float(y);
    o[2] = (float)(xl - xr);                 // Support::d = float(xL - xR), exact (inference.hpp:389-390)
  }
}

cudaError_t launch_emit_supports(const uint32_t* stage, const int32_t* rowmatch, const int32_t* rowoff,
                                 const long long* pair_base, void* out, long long cap, int W, int H, int n_pairs,
                                 cudaStream_t stream) {
  int rows = H - 2 * kRadius;
  if (rows <= 0 || n_pairs <= 0) return cudaSuccess;
  EmitArgs a{stage, rowmatch, rowoff, pair_base, reinterpret_cast<float*>(out), cap, W, H};
  emit_supports_kernel<<<dim3(rows, n_pairs), 128, 0, stream>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Candidate index list (ndb::arr2ind + border lambda output, raster order) from the hash image.
// One warp per row; debug / API-parity path only (PreprocessedImage::mask).
// ------------------------------------------------------------------------------------------------
__global__ void mask_scan_kernel(const int32_t* __restrict__ rowcnt, int H, int32_t* __restrict__ rowoff) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int y = 0; y < H; y++) { rowoff[y] = acc; acc += rowcnt[y]; }
    rowoff[H] = acc;
  }
}

__global__ void __launch_bounds__(32)
mask_rows_kernel(const uint32_t* __restrict__ hash, const int32_t* __restrict__ rowoff, int W, int H,
                 int32_t* __restrict__ mask, int cap) {
  const int y = blockIdx.x, lane = threadIdx.x;
  int off = rowoff[y];
  const uint32_t* row = hash + (size_t)y * W;
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const bool c = (x < W) && (row[x] >> 31);
    const uint32_t b = __ballot_sync(0xffffffffu, c);
    if (c) {
      int p = off + __popc(b & ((1u << lane) - 1u));
      if (p < cap) mask[p] = y * W + x;
    }
    off += __popc(b);
  }
}

cudaError_t launch_mask_list(const uint32_t* hash, const int32_t* rowcnt, int32_t* rowoff, int W, int H, int32_t* mask,
                             int cap, cudaStream_t stream) {
  mask_scan_kernel<<<1, 32, 0, stream>>>(rowcnt, H, rowoff);
  mask_rows_kernel<<<H, 32, 0, stream>>>(hash, rowoff, W, H, mask, cap);
  return cudaGetLastError();
}

}  // namespace gpc

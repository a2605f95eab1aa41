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
// (SURVEY.md 8a row M).  A row CTA never sorts its candidates.  It partitions them by a
// multiplicative hash of the state into ~one-element buckets in shared memory (one counting
// atomic per candidate, one scan, one scatter), lets every left candidate scan its own bucket
// for equal states (a handful of entries, early exit on the second duplicate), and sorts only
// the surviving matches (about a tenth of the candidates) before writing them out.
#include "gpc_device.cuh"

namespace gpc {

constexpr int kBuckets = 256;                 // ordering pass: counting sort on the top 8 state bits
constexpr int kBucketLimit = 64;              // above this the in-bucket rank pass falls back to a bitonic network

constexpr uint32_t kHashMul = 0x9E3779B1u;    // odd: state -> state * kHashMul is a bijection mod 2^32

// Shared-memory atomics in this kernel always consume their return value.  Fire-and-forget
// shared atomics (ATOMS with an RZ destination, addressed through a uniform register into the
// dynamic shared window) were observed on B200 / nvcc 12.9 to land late or at a wrong address
// (lost flag bits, corrupted neighbours, occasional illegal-address faults; reproduced with
// scripts/micro/match_harness.cu) -- the value-returning form does not show it.
__device__ __forceinline__ uint32_t atomic_inc_ret(uint32_t* p) { return atomicAdd(p, 1u); }

// shared memory: out[pow2cap] u64 | cnt[nb] u32 | entry[2*wcap] u32 | bcnt/bstart/bfill
size_t match_smem_bytes(int wcap, int table_log2) {
  size_t nb = (size_t)1 << table_log2;
  size_t pow2 = 1; while ((int)pow2 < wcap) pow2 <<= 1;
  size_t body = ((size_t)2 * wcap * 4 + 15) / 16 * 16;
  if (body < pow2 * 8) body = pow2 * 8;                       // out2 (ordering pass) reuses the entry array
  return pow2 * 8 + nb * 4 + body + 3 * kBuckets * 4 + 64;
}

// Orders the m matches of a row (records key << 32 | xl << 16 | xr in out[]) by state -- the keys are
// unique -- and writes xl << 16 | xr to the row's slice of `stage`.  Counting pass on the top 8 state bits,
// then an exact rank inside each (tiny) bucket; bitonic network for skewed states.
template <int kThreadsB>
__device__ __forceinline__ void order_and_stage(const MatchArgs& args, unsigned long long* out, unsigned long long* out2,
                                                uint32_t* bcnt, uint32_t* bstart, uint32_t* bfill, uint32_t* big_bucket,
                                                int m, int pair, int y) {
  const int tid = threadIdx.x, lane = tid & 31;
  // ---- order the matches by state (unique keys) and stage them ------------------------------------------
  uint32_t* stage = args.stage + ((size_t)pair * args.H + y) * args.W;
  if (m > 1) {
    // counting pass on the top 8 state bits, then an exact rank inside each (tiny) bucket
    const int shift = args.key_bits > 8 ? args.key_bits - 8 : 0;
    uint32_t seen = 0;
    for (int i = tid; i < m; i += kThreadsB) seen |= atomic_inc_ret(&bcnt[(uint32_t)(out[i] >> 32) >> shift]);
    if (seen == 0xffffffffu) __trap();
    __syncthreads();
    if (tid < 32) {                                   // exclusive scan of 256 counters: 8 per lane
      uint32_t c[8], sum = 0, mx = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) { c[k] = bcnt[8 * tid + k]; sum += c[k]; mx = max(mx, c[k]); }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      uint32_t run = incl - sum;
#pragma unroll
      for (int k = 0; k < 8; k++) { bstart[8 * tid + k] = run; run += c[k]; }
      mx = __reduce_max_sync(0xffffffffu, mx);
      if (tid == 0) *big_bucket = (mx > (uint32_t)kBucketLimit) ? 1u : 0u;
    }
    __syncthreads();
    if (!*big_bucket) {
      for (int i = tid; i < m; i += kThreadsB) {      // scatter into bucket segments (arbitrary order inside)
        const unsigned long long rec = out[i];
        const uint32_t b = (uint32_t)(rec >> 32) >> shift;
        out2[bstart[b] + atomic_inc_ret(&bfill[b])] = rec;
      }
      __syncthreads();
      for (int i = tid; i < m; i += kThreadsB) {      // rank inside the bucket, write to the final position
        const unsigned long long rec = out2[i];
        const uint32_t b = (uint32_t)(rec >> 32) >> shift;
        const uint32_t s0 = bstart[b], n = bcnt[b];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; j++) rank += (out2[s0 + j] < rec) ? 1u : 0u;
        stage[s0 + rank] = (uint32_t)(rec & 0xffffffffull);
      }
    } else {                                          // skewed states: bitonic network over all matches
      int p2 = 1; while (p2 < m) p2 <<= 1;
      for (int i = m + tid; i < p2; i += kThreadsB) out[i] = ~0ull;
      __syncthreads();
      for (int k = 2; k <= p2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < p2; i += kThreadsB) {
            const int l = i ^ j;
            if (l > i) {
              const unsigned long long a = out[i], b2 = out[l];
              const bool up = ((i & k) == 0);
              if ((a > b2) == up) { out[i] = b2; out[l] = a; }
            }
          }
          __syncthreads();
        }
      for (int i = tid; i < m; i += kThreadsB) stage[i] = (uint32_t)(out[i] & 0xffffffffull);
    }
  } else if (m == 1 && tid == 0) {
    stage[0] = (uint32_t)(out[0] & 0xffffffffull);
  }
}

// One CTA per (row, pair).  Every thread keeps its 4*KQ left and right pixels of the row in
// registers.  h = state * odd constant (a bijection on 32-bit words): the top log2(nb) bits pick
// the bucket, so inside a bucket two states are equal iff the remaining low bits of h are.  An
// entry therefore fits one word: remainder | side | x  (needs nb >= 2 * W, see launch).
//   count   : one atomicAdd per candidate on its bucket counter; the returned rank is kept
//   scan    : exclusive prefix of the counters -> word = start | count << 16 (bank-conflict free:
//             thread t owns buckets t, t+256, ...; the order of buckets in the entry array is free)
//   scatter : entry stored at start[bucket] + rank
//   resolve : every left candidate reads its bucket (<= 4 entries in one unrolled step, longer
//             buckets in a loop) counting equal states per side; unique on both sides = match
template <int KQ, int kThreadsB>
__global__ void __launch_bounds__(kThreadsB)
match_rows_kernel(const MatchArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int W = args.W, H = args.H, log2nb = args.table_log2, xb = args.x_bits;
  const int nb = 1 << log2nb;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int y = kRadius + blockIdx.x, pair = blockIdx.y;
  const int pow2cap = args.pow2cap;

  unsigned long long* out = reinterpret_cast<unsigned long long*>(smem);           // [pow2cap] emitted matches
  uint32_t* cnt = reinterpret_cast<uint32_t*>(out + pow2cap);                      // [nb] counters -> start | count << 16
  uint32_t* entry = cnt + nb;                                                       // [2*wcap] remainder | side | x
  unsigned long long* out2 = reinterpret_cast<unsigned long long*>(entry);          // ordering pass, reuses entry
  size_t body = ((size_t)2 * args.wcap * 4 + 15) / 16 * 16;
  if (body < (size_t)pow2cap * 8) body = (size_t)pow2cap * 8;
  uint32_t* bcnt = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(entry) + body);   // [kBuckets] x3
  uint32_t* bstart = bcnt + kBuckets;
  uint32_t* bfill = bstart + kBuckets;
  __shared__ int n_out, have_s[2];
  __shared__ uint32_t kmax_s, big_bucket, warp_tot[kThreadsB / 32];
  __shared__ int cmax_r, cmax_l, xmin_s;

  if (tid == 0) { n_out = 0; have_s[0] = 0; have_s[1] = 0; kmax_s = 0; cmax_r = 0; cmax_l = 0; xmin_s = 0x7fffffff; big_bucket = 0; }
  {
    uint4* z = reinterpret_cast<uint4*>(cnt);
    for (int i = tid; i < nb / 4; i += kThreadsB) z[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 3 * kBuckets; i += kThreadsB) bcnt[i] = 0u;
  }

  // ---- this thread's pixels of the left and right hash rows ---------------------------------------
  const uint4* row_l = reinterpret_cast<const uint4*>(args.hash + ((size_t)(2 * pair) * H + y) * W);
  const uint4* row_r = reinterpret_cast<const uint4*>(args.hash + ((size_t)(2 * pair + 1) * H + y) * W);
  const int nquads = W / 4;
  uint32_t v[2][4 * KQ];                       // [side][element]
  uint32_t any_l = 0, any_r = 0;
#pragma unroll
  for (int k = 0; k < KQ; k++) {
    const int q = tid + k * kThreadsB;
    uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
    if (q < nquads) { a = __ldg(row_l + q); b = __ldg(row_r + q); }
    v[0][4 * k] = a.x; v[0][4 * k + 1] = a.y; v[0][4 * k + 2] = a.z; v[0][4 * k + 3] = a.w;
    v[1][4 * k] = b.x; v[1][4 * k + 1] = b.y; v[1][4 * k + 2] = b.z; v[1][4 * k + 3] = b.w;
    any_l |= a.x | a.y | a.z | a.w;
    any_r |= b.x | b.y | b.z | b.w;
  }
  __syncthreads();
  if (any_l >> 31) have_s[0] = 1;                                              // benign same-value race
  if (any_r >> 31) have_s[1] = 1;
  __syncthreads();
  int m = 0;

  if (have_s[0] && have_s[1]) {
    const int rs = 32 - log2nb;                  // bucket = h >> rs
    const int es = log2nb - xb - 1;              // entry  = (h << log2nb) >> es | side << xb | x
    // ---- count: bucket counters; ranks packed two per register ------------------------------------
    uint32_t rank[2][2 * KQ];
#pragma unroll
    for (int sd = 0; sd < 2; sd++)
#pragma unroll
      for (int e = 0; e < 4 * KQ; e++) {
        uint32_t r = 0;
        if (v[sd][e] >> 31) r = atomic_inc_ret(&cnt[((v[sd][e] & 0x7fffffffu) * kHashMul) >> rs]);
        if (e & 1) rank[sd][e >> 1] |= r << 16; else rank[sd][e >> 1] = r;
      }
    __syncthreads();
    // ---- exclusive scan over the buckets in "layout order": thread t owns the quads of buckets
    // 4 * (k * 256 + t) .. + 3, k = 0, 1, ... (16-byte accesses, conflict free); the order of buckets
    // in the entry array is free, it only has to be a partition ----------------------------------------
    {
      uint4* c4 = reinterpret_cast<uint4*>(cnt);
      const int per = nb / (4 * kThreadsB);        // nb >= 1024 by construction
      uint32_t sum = 0;
      for (int k = 0; k < per; k++) { const uint4 q = c4[k * kThreadsB + tid]; sum += q.x + q.y + q.z + q.w; }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      if (lane == 31) warp_tot[wid] = incl;
      __syncthreads();
      uint32_t base = incl - sum;
#pragma unroll
      for (int w = 0; w < kThreadsB / 32; w++) if (w < wid) base += warp_tot[w];
      for (int k = 0; k < per; k++) {
        uint4 q = c4[k * kThreadsB + tid];
        uint4 o;
        o.x = base | (q.x << 16); base += q.x;
        o.y = base | (q.y << 16); base += q.y;
        o.z = base | (q.z << 16); base += q.z;
        o.w = base | (q.w << 16); base += q.w;
        c4[k * kThreadsB + tid] = o;
      }
    }
    __syncthreads();
    // ---- scatter ---------------------------------------------------------------------------------------
#pragma unroll
    for (int sd = 0; sd < 2; sd++)
#pragma unroll
      for (int e = 0; e < 4 * KQ; e++) {
        if (v[sd][e] >> 31) {
          const uint32_t h = (v[sd][e] & 0x7fffffffu) * kHashMul;
          const uint32_t r = (e & 1) ? (rank[sd][e >> 1] >> 16) : (rank[sd][e >> 1] & 0xffffu);
          const uint32_t x = 4u * (uint32_t)(tid + (e >> 2) * kThreadsB) + (uint32_t)(e & 3);
          entry[(cnt[h >> rs] & 0xffffu) + r] = ((h << log2nb) >> es) | ((uint32_t)sd << xb) | x;
        }
      }
    // ---- tail rules: only the globally last right key (largest row with right candidates) -----------
    const bool last_row = (args.lastrow[2 * pair + 1] == y);
    if (last_row) {
      uint32_t km = 0;
#pragma unroll
      for (int e = 0; e < 4 * KQ; e++) if (v[1][e] >> 31) km = max(km, v[1][e] & 0x7fffffffu);
      km = __reduce_max_sync(0xffffffffu, km);
      if (lane == 0 && atomicMax(&kmax_s, km) == 0xffffffffu) __trap();
      __syncthreads();
      km = kmax_s;
      int cr = 0, cl = 0, xm = 0x7fffffff;
#pragma unroll
      for (int e = 0; e < 4 * KQ; e++) {
        if ((v[1][e] >> 31) && (v[1][e] & 0x7fffffffu) == km) { cr++; xm = min(xm, 4 * (tid + (e >> 2) * kThreadsB) + (e & 3)); }
        if ((v[0][e] >> 31) && (v[0][e] & 0x7fffffffu) == km) cl++;
      }
      cr = __reduce_add_sync(0xffffffffu, cr);
      cl = __reduce_add_sync(0xffffffffu, cl);
      xm = __reduce_min_sync(0xffffffffu, xm);
      if (lane == 0) {
        if (atomicAdd(&cmax_r, cr) < 0 || atomicAdd(&cmax_l, cl) < 0 || atomicMin(&xmin_s, xm) < 0) __trap();
      }
    }
    __syncthreads();
    // ---- resolve + emit: a left state that is unique in its bucket on both sides is a match ---------
    const uint32_t kmax = kmax_s;
    const int cmr = cmax_r, cml = cmax_l, xmin = xmin_s;
    const uint32_t XL = 1u << xb;
    unsigned long long rec[4 * KQ];
    uint32_t okmask = 0;
#pragma unroll
    for (int e = 0; e < 4 * KQ; e++) {
      rec[e] = 0;
      if (v[0][e] >> 31) {
        const uint32_t key = v[0][e] & 0x7fffffffu;
        const uint32_t h = key * kHashMul;
        const uint32_t word = cnt[h >> rs];
        const uint32_t s0 = word & 0xffffu, n = word >> 16;
        const uint32_t mine = (h << log2nb) >> es;             // remainder field, side = 0, x = 0
        // t = entry ^ mine: same state, left  <=> t < XL (t = its x);  same state, right <=> XL <= t < 2 XL
        uint32_t nl = 0, nr = 0, xr = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t t = (j < (int)n) ? (entry[s0 + j] ^ mine) : 0xffffffffu;
          const uint32_t u = t - XL;
          nl += (t < XL) ? 1u : 0u;
          nr += (u < XL) ? 1u : 0u;
          xr = (u < XL) ? u : xr;
        }
        for (uint32_t j = 4; j < n && nl < 2u && nr < 2u; j++) { // rare: bucket longer than four entries
          const uint32_t t = entry[s0 + j] ^ mine;
          const uint32_t u = t - XL;
          nl += (t < XL) ? 1u : 0u;
          nr += (u < XL) ? 1u : 0u;
          xr = (u < XL) ? u : xr;
        }
        bool ok = (nl == 1u) && (nr == 1u);
        const int xl = 4 * (tid + (e >> 2) * kThreadsB) + (e & 3);
        if (last_row && key == kmax) {         // inference.hpp:243-249 on the tail of the sorted right keys
          ok = (cml == 1) && (cmr == 2);       // 1 right: the last element never matches; >=3: duplicates
          xr = (uint32_t)xmin;                 // 2: "first of the two" := smaller x (stable order)
        }
        const int dx = xl - (int)xr;
        ok = ok && (dx <= args.disp_high && -dx <= args.disp_high) && (0 <= args.vertical_tolerance);
        if (ok) {
          okmask |= 1u << e;
          rec[e] = ((unsigned long long)key << 32) | ((unsigned long long)xl << 16) | (unsigned long long)xr;
        }
      }
    }
    // one shared-memory reservation per warp for all of its matches
    {
      const int mine_n = __popc(okmask);
      int incl = mine_n;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      const int warp_n = __shfl_sync(0xffffffffu, incl, 31);
      int base = 0;
      if (lane == 31 && warp_n > 0) base = atomicAdd(&n_out, warp_n);
      base = __shfl_sync(0xffffffffu, base, 31) + incl - mine_n;
#pragma unroll
      for (int e = 0; e < 4 * KQ; e++)
        if ((okmask >> e) & 1u) out[base++] = rec[e];
    }
    __syncthreads();
    m = n_out;
  }

  order_and_stage<kThreadsB>(args, out, out2, bcnt, bstart, bfill, &big_bucket, m, pair, y);
  if (tid == 0) args.rowmatch[(size_t)pair * H + y] = m;
}

// Launch shapes: every thread owns 4 * KQ pixels per side.  Rows up to 1024 pixels run 256 threads,
// wider rows 512 or 1024 threads with KQ = 1 (wide rows need large tables, so few CTAs fit an SM and
// each has to bring more warps); beyond 4096 pixels KQ grows.
template <int KQ, int T>
static cudaError_t configure_one(int max_smem) {
  return cudaFuncSetAttribute(match_rows_kernel<KQ, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

cudaError_t configure_match_rows(int max_smem) {
  cudaError_t e = configure_one<1, 256>(max_smem);
  if (e == cudaSuccess) e = configure_one<1, 512>(max_smem);
  if (e == cudaSuccess) e = configure_one<1, 1024>(max_smem);
  if (e == cudaSuccess) e = configure_one<2, 1024>(max_smem);
  return e;
}

int match_rows_threads(int W) {
  const int quads = W / 4;
  return quads <= 256 ? 256 : quads <= 512 ? 512 : 1024;
}

cudaError_t launch_match_rows(const MatchArgs& args, int n_pairs, cudaStream_t stream) {
  int rows = args.H - 2 * kRadius;
  if (rows <= 0 || n_pairs <= 0) return cudaSuccess;
  const int quads = args.W / 4;
  dim3 grid(rows, n_pairs);
  size_t smem = match_smem_bytes(args.wcap, args.table_log2);
  if (quads <= 256) match_rows_kernel<1, 256><<<grid, 256, smem, stream>>>(args);
  else if (quads <= 512) match_rows_kernel<1, 512><<<grid, 512, smem, stream>>>(args);
  else if (quads <= 1024) match_rows_kernel<1, 1024><<<grid, 1024, smem, stream>>>(args);
  else if (quads <= 2048) match_rows_kernel<2, 1024><<<grid, 1024, smem, stream>>>(args);
  else return cudaErrorInvalidValue;
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

constexpr int kEmitRows = 8;       // rows per CTA, one warp each

__global__ void __launch_bounds__(32 * kEmitRows)
emit_supports_kernel(const EmitArgs a) {
  const int y = kRadius + blockIdx.x * kEmitRows + (threadIdx.x >> 5), pair = blockIdx.y, lane = threadIdx.x & 31;
  if (y >= a.H - kRadius) return;
  const int m = a.rowmatch[(size_t)pair * a.H + y];
  if (m == 0) return;
  const long long off = a.rowoff[(size_t)pair * a.H + y];
  long long base, limit;
  if (a.pair_base) { base = a.pair_base[pair]; limit = a.cap; }
  else { base = (long long)pair * a.cap; limit = base + a.cap; }
  const uint32_t* stage = a.stage + ((size_t)pair * a.H + y) * a.W;
  for (int k = lane; k < m; k += 32) {
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
  emit_supports_kernel<<<dim3((rows + kEmitRows - 1) / kEmitRows, n_pairs), 32 * kEmitRows, 0, stream>>>(a);
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

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
#include <cstdlib>

#include "gpc_device.cuh"

namespace gpc {

constexpr int kBuckets = 256;                 // ordering pass: counting sort on the top 8 state bits
constexpr int kBucketLimit = 64;              // above this the in-bucket rank pass falls back to a bitonic network

constexpr uint32_t kHashMul = 0x9E3779B1u;    // odd: state -> state * kHashMul is a bijection mod 2^32

__device__ __forceinline__ uint32_t atomic_inc_ret(uint32_t* p) { return atomicAdd(p, 1u); }

// Bounds checks of every computed shared-memory index, compiled in with -DGPC_DEBUG_CHECKS (the checked build the
// GPU tests and the fuzz campaign are also run with, see profiles/r02_checked_build.md; compute-sanitizer is not
// available on the GPU pool).  A failed check traps, which the host sees as a CUDA error.
#ifdef GPC_DEBUG_CHECKS
#define GPC_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define GPC_CHECK(cond) do { } while (0)
#endif

constexpr int kOvCap = 256;                   // fast matcher: capacity of each side's overflow list

// shared memory of the general matcher: out[pow2cap] u64 | cnt[nb] u32 | entry[2*wcap] u32 | bcnt, bstart
size_t match_smem_bytes(int wcap, int table_log2) {
  size_t nb = (size_t)1 << table_log2;
  size_t pow2 = 1; while ((int)pow2 < wcap) pow2 <<= 1;
  size_t body = ((size_t)2 * wcap * 4 + 15) / 16 * 16;
  if (body < pow2 * 8) body = pow2 * 8;                       // out2 (ordering pass) reuses the entry array
  return pow2 * 8 + nb * 4 + body + 2 * kBuckets * 4 + 64;
}

// shared memory of the fast matcher: nib[2^nib_log2] u32 | slot[2^slot_log2] u32 | live list 2 x [pow2cap] u32
size_t match_fast_smem_bytes(int nib_log2, int slot_log2, int pow2cap) {
  return ((size_t)4 << nib_log2) + ((size_t)4 << slot_log2) + (size_t)pow2cap * 8;
}
size_t order_rows_smem_bytes(int pow2cap) { return (size_t)pow2cap * 16 + 2 * kBuckets * 4; }
int match_ov_cap() { return kOvCap; }

// Orders the m matches of a row (records key << 32 | xl << 16 | xr in out[]) by state -- the keys are
// unique -- and writes xl << 16 | xr to the row's slice of `stage`.  Counting pass on the top 8 state bits,
// then an exact rank inside each (tiny) bucket; bitonic network for skewed states.  bcnt[] must be zero.
template <int kThreadsB>
__device__ __forceinline__ void order_and_stage(const MatchArgs& args, unsigned long long* out, unsigned long long* out2,
                                                uint32_t* bcnt, uint32_t* bstart, uint32_t* big_bucket,
                                                int m, int pair, int y) {
  const int tid = threadIdx.x, lane = tid & 31;
  uint32_t* stage = args.stage + ((size_t)pair * args.H + y) * args.W;
  if (m > 1) {
    const int shift = args.key_bits > 8 ? args.key_bits - 8 : 0;
    for (int i = tid; i < m; i += kThreadsB) atomicAdd(&bcnt[(uint32_t)(out[i] >> 32) >> shift], 1u);
    __syncthreads();
    if (tid < 32) {                                   // inclusive scan of 256 counters, 8 per lane: bstart = END of the bucket
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
      for (int i = tid; i < m; i += kThreadsB) {      // scatter into bucket segments (arbitrary order inside); the
        const unsigned long long rec = out[i];        // cursor bstart[b] ends at the bucket's end
        const uint32_t b = (uint32_t)(rec >> 32) >> shift;
        out2[atomic_inc_ret(&bstart[b])] = rec;
      }
      __syncthreads();
      for (int i = tid; i < m; i += kThreadsB) {      // rank inside the bucket, write to the final position
        const unsigned long long rec = out2[i];
        const uint32_t b = (uint32_t)(rec >> 32) >> shift;
        const uint32_t n = bcnt[b], s0 = bstart[b] - n;
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

// ------------------------------------------------------------------------------------------------
// General row matcher (any multiplicities, tail rules).  Every thread keeps its 4*KQ left and right
// pixels of the row in registers.  h = state * odd constant (a bijection on 32-bit words): the top
// log2(nb) bits pick the bucket, so inside a bucket two states are equal iff the remaining low bits
// of h are.  An entry therefore fits one word: remainder | side | x  (needs nb >= 2 * W, see launch).
//   count   : one atomicAdd per candidate on its bucket counter; the returned rank is kept
//   scan    : exclusive prefix of the counters -> word = start | count << 16 (bank-conflict free:
//             thread t owns buckets t, t+256, ...; the order of buckets in the entry array is free)
//   scatter : entry stored at start[bucket] + rank
//   resolve : every left candidate reads its bucket (<= 4 entries in one unrolled step, longer
//             buckets in a loop) counting equal states per side; unique on both sides = match
// Since round 2 it serves the rows the fast matcher hands over (the row holding the globally last right
// key, rows with many duplicated states) and GPC_MATCHER_ROWS_GENERAL.
// ------------------------------------------------------------------------------------------------
template <int KQ, int kThreadsB>
__device__ __forceinline__ void general_row(const MatchArgs& args, uint8_t* smem, const int pair, const int y) {
  const int W = args.W, H = args.H, log2nb = args.table_log2, xb = args.x_bits;
  const int nb = 1 << log2nb;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int pow2cap = args.pow2cap;

  unsigned long long* out = reinterpret_cast<unsigned long long*>(smem);           // [pow2cap] emitted matches
  uint32_t* cnt = reinterpret_cast<uint32_t*>(out + pow2cap);                      // [nb] counters -> start | count << 16
  uint32_t* entry = cnt + nb;                                                       // [2*wcap] remainder | side | x
  unsigned long long* out2 = reinterpret_cast<unsigned long long*>(entry);          // ordering pass, reuses entry
  size_t body = ((size_t)2 * args.wcap * 4 + 15) / 16 * 16;
  if (body < (size_t)pow2cap * 8) body = (size_t)pow2cap * 8;
  uint32_t* bcnt = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(entry) + body);   // [kBuckets] x2
  uint32_t* bstart = bcnt + kBuckets;
  __shared__ int n_out, have_s[2];
  __shared__ uint32_t kmax_s, big_bucket, warp_tot[kThreadsB / 32];
  __shared__ int cmax_r, cmax_l, xmin_s;

  if (tid == 0) { n_out = 0; have_s[0] = 0; have_s[1] = 0; kmax_s = 0; cmax_r = 0; cmax_l = 0; xmin_s = 0x7fffffff; big_bucket = 0; }
  {
    uint4* z = reinterpret_cast<uint4*>(cnt);
    for (int i = tid; i < nb / 4; i += kThreadsB) z[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kBuckets; i += kThreadsB) bcnt[i] = 0u;
  }

  // ---- this thread's pixels of the left and right hash rows ---------------------------------------
  const uint4* row_l = reinterpret_cast<const uint4*>(args.hash + ((size_t)(2 * pair) * H + y) * W);
  const uint4* row_r = reinterpret_cast<const uint4*>(args.hash + ((size_t)(2 * pair + 1) * H + y) * W);
  const int nquads = W / 4;
  const int rowcnt_l = __ldg(args.rowcnt + (size_t)(2 * pair) * H + y), rowcnt_r = __ldg(args.rowcnt + (size_t)(2 * pair + 1) * H + y);
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
          GPC_CHECK((cnt[h >> rs] & 0xffffu) + r < 2u * (uint32_t)args.wcap);
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
      if (lane == 0) atomicMax(&kmax_s, km);
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
      if (lane == 0) { atomicAdd(&cmax_r, cr); atomicAdd(&cmax_l, cl); atomicMin(&xmin_s, xm); }
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
        if ((okmask >> e) & 1u) { GPC_CHECK(base < pow2cap); out[base++] = rec[e]; }
    }
    __syncthreads();
    m = n_out;
  }

  order_and_stage<kThreadsB>(args, out, out2, bcnt, bstart, &big_bucket, m, pair, y);
  if (tid == 0) args.rowmatch[(size_t)pair * H + y] = m;
}

// hdr == nullptr: one CTA per (row, pair) of the grid.  Otherwise the CTAs walk the row list the fast matcher left
// (push_row below); the last CTA to finish clears the two counters for the next launch.
template <int KQ, int kThreadsB>
__global__ void __launch_bounds__(kThreadsB)
match_rows_general_kernel(const MatchArgs args, uint32_t* hdr, const uint32_t* ent) {
  extern __shared__ __align__(16) uint8_t smem[];
  if (hdr == nullptr) {
    general_row<KQ, kThreadsB>(args, smem, blockIdx.y, kRadius + blockIdx.x);
    return;
  }
  const uint32_t n = *reinterpret_cast<volatile uint32_t*>(hdr);
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    general_row<KQ, kThreadsB>(args, smem, (int)ent[2 * i], (int)ent[2 * i + 1]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&hdr[1], 1u) == gridDim.x - 1) { hdr[0] = 0u; hdr[1] = 0u; }
  }
}

// ------------------------------------------------------------------------------------------------
// Fast row matcher (round 2).  A match is a state that occurs exactly once in the left row and exactly
// once in the right row (inference.hpp:227-254 in epipolar mode); about nine candidates in ten have no
// partner at all, so the kernel finds that out as cheaply as possible and continues with short lists:
//   right    : 4 bits per bucket -- "a left", "a second left", "a right", "a second right" -- eight
//              buckets per word, ~32 buckets per candidate.  h = v * odd constant (a bijection; the
//              candidate flag stays in v, every candidate carries it); the top bits select the bucket.
//              Every right candidate ORs its bit in (a second atomicOr for the few that find it set).
//   left     : every left candidate ORs its bit in; the value the atomic returns already holds the final
//              right bits, so a candidate whose bucket has no right candidate is dropped on the spot
//              (~85 %); the others go to the "live" list (v, x).
//   pair     : a right candidate re-reads its nibble; in a bucket with exactly one left and one right it
//              claims slot[h >> ..] with atomicCAS (remainder | x in one word).  Buckets with both sides
//              and a duplicate go to the overflow lists; so does the loser of a slot (the slot table is
//              coarser than the buckets).
//   resolve  : one thread per live left entry: in a pair bucket it reads the slot -- same remainder = same
//              state = match; a stranger's entry = its partner lost the slot -> overflow list.
// Matches (unordered) and the overflow lists go to global memory; the warp-per-row tail kernel below resolves the
// overflow entries and orders the matches.  The row holding the globally last right key (tail rules) and rows whose
// overflow lists do not fit are appended to a row list for the general kernel.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kClassLut = (1u << 10) | (2u << 14) | (2u << 26) | (2u << 30);   // nibble -> 1 pair bucket, 2 overflow
#ifndef GPC_B_MINB
#define GPC_B_MINB 6                          // resident 256-thread CTAs per SM the register budget allows
#endif

__device__ __forceinline__ uint32_t atom_add_shared(uint32_t* p, uint32_t v) {
  // plain atom: ptxas wraps a uniform-address atomicAdd whose result is used into a 20-instruction warp scan
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
  return old;
}

// Row lists (rows handed to the general kernel / the block-wide ordering kernel): hdr[0] = number of rows,
// hdr[1] = CTAs of the consuming kernel that are done, ent[2i] = pair, ent[2i + 1] = row.  Headers and entries live in
// separate arrays: the slices of the resident buffers that different launches work on overlap in arbitrary ways.
__device__ __forceinline__ void push_row(uint32_t* hdr, uint32_t* ent, int pair, int y) {
  const uint32_t i = atomicAdd(&hdr[0], 1u);
  ent[2 * i] = (uint32_t)pair; ent[2 * i + 1] = (uint32_t)y;
}

template <int KQ, int kThreadsB>
__global__ void __launch_bounds__(kThreadsB, (KQ == 1 && kThreadsB <= 512) ? (GPC_B_MINB * 256) / kThreadsB : 1)
match_rows_fast_kernel(const MatchArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  static_assert(kOvCap % 4 == 0, "16-byte reads of the overflow lists");
  const int W = args.W, H = args.H, xb = args.x_bits, nwl = args.nib_log2, nsl = args.slot_log2;
  const int tid = threadIdx.x;
  const int y = kRadius + blockIdx.x, pair = blockIdx.y;

  uint32_t* nib = reinterpret_cast<uint32_t*>(smem);                                // [1 << nwl]
  uint32_t* slot = nib + ((size_t)1 << nwl);                                        // [1 << nsl]
  uint32_t* live_v = slot + ((size_t)1 << nsl);                                     // [pow2cap] live left candidates: v
  uint32_t* live_x = live_v + args.pow2cap;                                         // [pow2cap] x
  __shared__ uint32_t n_live, n_ovl, n_ovr, n_out;
  // row indices in 32 bits (gpc_create bounds max_batch * max_h): ONE widening multiply per base address instead of
  // 64-bit products assembled from several IMAD.WIDE / IMAD (a wide multiply holds the dispatch port for 3 - 4 cycles)
  const uint32_t grow = (uint32_t)pair * (uint32_t)H + (uint32_t)y;                 // this row's records in global memory:
  unsigned long long* mrec = args.mrec + (size_t)grow * (uint32_t)W;                //   matches, unordered (key << 32 | xl << 16 | xr)
  uint32_t* ovl_s = args.ovbuf + (size_t)grow * (uint32_t)(4 * kOvCap);             //   overflow lists: left v, x
  uint32_t* ovl_x = ovl_s + kOvCap;
  uint32_t* ovr_s = ovl_x + kOvCap;                                                 //   right v, x
  uint32_t* ovr_x = ovr_s + kOvCap;
  int32_t* hdr = args.rowhdr + (size_t)grow * 4u;                                   //   {matches, left overflow, right overflow, done}

  // ---- this thread's pixels of the left and right hash rows (issued first: the zero fill hides their latency)
  const uint32_t irow_l = grow + (uint32_t)pair * (uint32_t)H;                      // row y of image 2 * pair
  const uint4* row_l = reinterpret_cast<const uint4*>(args.hash + (size_t)irow_l * (uint32_t)W);
  const uint4* row_r = row_l + (uint32_t)H * (uint32_t)(W / 4);                     // image 2 * pair + 1
  const int nquads = W / 4;
  const int rowcnt_l = __ldg(args.rowcnt + irow_l), rowcnt_r = __ldg(args.rowcnt + irow_l + (uint32_t)H);
  uint32_t v[2][4 * KQ];                       // [side][element]
#pragma unroll
  for (int k = 0; k < KQ; k++) {
    const int q = tid + k * kThreadsB;
    uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
    if (q < nquads) { a = __ldg(row_l + q); b = __ldg(row_r + q); }
    v[0][4 * k] = a.x; v[0][4 * k + 1] = a.y; v[0][4 * k + 2] = a.z; v[0][4 * k + 3] = a.w;
    v[1][4 * k] = b.x; v[1][4 * k + 1] = b.y; v[1][4 * k + 2] = b.z; v[1][4 * k + 3] = b.w;
  }
  if (tid == 0) { n_live = 0; n_ovl = 0; n_ovr = 0; n_out = 0; }
  {
    uint4* z = reinterpret_cast<uint4*>(nib) + tid;
    const int rounds = ((1 << nwl) + (1 << nsl)) / (4 * kThreadsB);            // tables are multiples of 4 * kThreadsB words
    const uint4 zero = make_uint4(0, 0, 0, 0);
    if (rounds == 6) {                                                         // the usual sizing (16 + 8 bytes per candidate slot): straight-line stores
#pragma unroll
      for (int i = 0; i < 6; i++) z[i * kThreadsB] = zero;
    } else {
      for (int i = 0; i < rounds; i++, z += kThreadsB) *z = zero;
    }
  }
  // a side without candidates: kernel A1's per-row counts say so (one uniform load each instead of two block-wide
  // reductions over the candidate flags; the rows are issued with the hash loads above)
  const int have_l = rowcnt_l, have_r = rowcnt_r;
  __syncthreads();                                               // orders the zero fill before the atomics
  if (!(have_l && have_r)) {
    if (tid == 0) { args.rowmatch[grow] = 0; hdr[0] = 0; hdr[1] = 0; hdr[2] = 0; hdr[3] = 1; }
    return;
  }
  if (args.lastrow[2 * pair + 1] == y) {       // tail rules of inference.hpp:243-249: the general kernel's job
    if (tid == 0) {
      push_row(args.fb_hdr, args.fb_ent, pair, y);
      hdr[0] = 0; hdr[1] = 0; hdr[2] = 0; hdr[3] = 1;
    }
    return;
  }

  // Everything below is branch-free per pixel (a non-candidate ORs 0 into some word and is never live): the
  // pixels of a side keep independent chains in flight.
  const int bs = 29 - nwl;                     // bucket = h >> bs: word = bucket >> 3, nibble = bucket & 7
  const int s_addr = bs + 1, s_sh = bs - 2;    // byte offset of the word = (h >> s_addr) & ~3, bit offset of the nibble = (h >> s_sh) & 28
  uint8_t* const nib8 = reinterpret_cast<uint8_t*>(nib);
  // ---- right candidates: count ------------------------------------------------------------------------
  uint32_t pk[4 * KQ];                         // byte offset of the word | bit offset of the nibble << 16
  {
    uint32_t bit[4 * KQ], old[4 * KQ];
#pragma unroll
    for (int e = 0; e < 4 * KQ; e++) {
      const uint32_t h = v[1][e] * kHashMul;
      const uint32_t off = (h >> s_addr) & ~3u, sh = (h >> s_sh) & 28u;
      pk[e] = off | (sh << 16);
      bit[e] = (v[1][e] >> 31) << (sh | 2u);
      GPC_CHECK(off < (4u << nwl) && (sh | 2u) < 32u);
      old[e] = atomicOr(reinterpret_cast<uint32_t*>(nib8 + off), bit[e]);
    }
#pragma unroll
    for (int e = 0; e < 4 * KQ; e++)
      if (old[e] & bit[e]) atomicOr(reinterpret_cast<uint32_t*>(nib8 + (pk[e] & 0xffffu)), bit[e] << 1);
  }
  __syncthreads();
  // ---- left candidates: count; the returned word holds the final right bits -----------------------------
  {
    uint32_t bit[4 * KQ], old[4 * KQ], livem = 0;
#pragma unroll
    for (int e = 0; e < 4 * KQ; e++) {
      const uint32_t h = v[0][e] * kHashMul;
      const uint32_t off = (h >> s_addr) & ~3u, sh = (h >> s_sh) & 28u;
      bit[e] = (v[0][e] >> 31) << sh;
      old[e] = atomicOr(reinterpret_cast<uint32_t*>(nib8 + off), bit[e]);
      if (old[e] & bit[e]) atomicOr(reinterpret_cast<uint32_t*>(nib8 + off), bit[e] << 1);
    }
#pragma unroll
    for (int e = 0; e < 4 * KQ; e++) livem |= ((old[e] & (bit[e] << 2)) ? 1u : 0u) << e;      // a right candidate in my bucket
    if (livem) {
      uint32_t p = atom_add_shared(&n_live, (uint32_t)__popc(livem));
      do {
        const int e = __ffs((int)livem) - 1;
        livem &= livem - 1u;
        uint32_t val = v[0][0];
#pragma unroll
        for (int k = 1; k < 4 * KQ; k++) val = (e == k) ? v[0][k] : val;
        GPC_CHECK(p < (uint32_t)args.pow2cap);
        live_v[p] = val;
        live_x[p] = 4u * (uint32_t)(tid + (e >> 2) * kThreadsB) + (uint32_t)(e & 3);
        p++;
      } while (livem);
    }
  }
  __syncthreads();
  // ---- right candidates of pair buckets claim their slot -----------------------------------------------
  const int ss = 32 - nsl;                     // slot = h >> ss, entry = (h << nsl) >> (nsl - xb) | x
  const int s_slot = ss - 2;                   // byte offset of the slot = (h >> s_slot) & ~3
  uint8_t* const slot8 = reinterpret_cast<uint8_t*>(slot);
  {
    uint32_t ovm = 0;
#pragma unroll
    for (int e = 0; e < 4 * KQ; e++) {
      const uint32_t t = *reinterpret_cast<const uint32_t*>(nib8 + (pk[e] & 0xffffu)) >> (pk[e] >> 16);
      uint32_t c = (v[1][e] >> 31) ? ((kClassLut >> ((t & 15u) << 1)) & 3u) : 0u;
      const uint32_t h = v[1][e] * kHashMul;
      const uint32_t x = 4u * (uint32_t)(tid + (e >> 2) * kThreadsB) + (uint32_t)(e & 3);
      uint32_t prev = 0;
      GPC_CHECK(((h >> s_slot) & ~3u) < (4u << nsl) && (c != 1u || x != 0u));
      if (c == 1u) prev = atomicCAS(reinterpret_cast<uint32_t*>(slot8 + ((h >> s_slot) & ~3u)), 0u, ((h << nsl) >> (nsl - xb)) | x);
      c = prev ? 2u : c;                       // slot taken by another pair bucket: overflow
      ovm |= (c >> 1) << e;
    }
    if (ovm) {
      uint32_t p = atom_add_shared(&n_ovr, (uint32_t)__popc(ovm));
      do {
        const int e = __ffs((int)ovm) - 1;
        ovm &= ovm - 1u;
        uint32_t val = v[1][0];
#pragma unroll
        for (int k = 1; k < 4 * KQ; k++) val = (e == k) ? v[1][k] : val;
        if (p < (uint32_t)kOvCap) { ovr_s[p] = val; ovr_x[p] = 4u * (uint32_t)(tid + (e >> 2) * kThreadsB) + (uint32_t)(e & 3); }
        p++;
      } while (ovm);
    }
  }
  __syncthreads();
  // ---- live left candidates: pair buckets read their slot -------------------------------------------------
  const uint32_t XL = 1u << xb;
  const int same_bucket_shift = xb + 29 - nwl;       // entry bits above this: the bucket bits the slot index lacks
  const uint32_t nlive = n_live;
  for (uint32_t i = tid; i < nlive; i += kThreadsB) {
    const uint32_t lv = live_v[i], xl = live_x[i];
    const uint32_t h = lv * kHashMul;
    const uint32_t nibble = (*reinterpret_cast<const uint32_t*>(nib8 + ((h >> s_addr) & ~3u)) >> ((h >> s_sh) & 28u)) & 15u;
    const uint32_t t = *reinterpret_cast<const uint32_t*>(slot8 + ((h >> s_slot) & ~3u)) ^ ((h << nsl) >> (nsl - xb));
    const int dx = (int)xl - (int)t;
    // pair bucket, t < XL: same state, t = x of the right candidate.  Otherwise bits above same_bucket_shift tell
    // a stranger's entry (my partner lost the slot: overflow lists) from the bucket's own right candidate.
    const bool pairb = (nibble == 5u);
    if (pairb && t < XL && dx <= args.disp_high && -dx <= args.disp_high && 0 <= args.vertical_tolerance) {
      const uint32_t p = atom_add_shared(&n_out, 1u);
      GPC_CHECK(p < (uint32_t)W);
      mrec[p] = ((unsigned long long)(lv & 0x7fffffffu) << 32) | (unsigned long long)((xl << 16) | t);
    } else if (nibble != 5u || (t >> same_bucket_shift) != 0u) {     // duplicates in the bucket, or a stranger in the slot
      const uint32_t p = atom_add_shared(&n_ovl, 1u);
      if (p < (uint32_t)kOvCap) { ovl_s[p] = lv; ovl_x[p] = xl; }
    }
  }
  __syncthreads();
  // ---- hand over: the tail kernel resolves the overflow lists and orders the matches -------------------------
  if (tid == 0) {
    const uint32_t nl = n_ovl, nr = n_ovr;
    if (nl > (uint32_t)kOvCap || nr > (uint32_t)kOvCap) {        // too many duplicated states: general kernel
      push_row(args.fb_hdr, args.fb_ent, pair, y);
      hdr[0] = 0; hdr[1] = 0; hdr[2] = 0; hdr[3] = 1;
    } else {
      hdr[0] = (int32_t)n_out; hdr[1] = (int32_t)nl; hdr[2] = (int32_t)nr; hdr[3] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tail of the fast matcher: ONE WARP per row (no block barriers, small footprint, many rows in flight per SM),
// for the low-parallelism work the row CTA would otherwise hold its shared memory and seven idle warps for:
//   overflow : every left overflow entry counts equal states in both lists (tens of entries; complete per state,
//              because equal states share a bucket and a bucket is listed as a whole); unique on both sides and
//              inside the disparity range = one more match record
//   order    : the row's matches sorted by state (unique keys): counting pass on the top 8 state bits, exact rank
//              inside each (tiny) bucket, written as xl << 16 | xr to `stage`; rowmatch[row] = their number
// Rows with more than kTailCap matches, or a crowded bucket, go to a list for the block-wide ordering kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kTailWarps = 8;                 // rows per CTA
constexpr int kTailCap = 256;                 // matches a warp orders in its slice of shared memory


// per-warp shared memory: kGroups group counters + kGroups cursors (the ordering uses 32 of each), keys and payloads of
// up to kSlots overflow entries / kTailCap match records.  Two sizes: rows up to 1024 pixels rarely list more than a few
// dozen overflow entries, so the small slice (3 KB per warp, 8 CTAs = 64 warps per SM) serves them; wide rows get the full one.
#ifndef GPC_TAIL_MINB_SMALL
#define GPC_TAIL_MINB_SMALL 8
#endif
#ifndef GPC_TAIL_GROUPS_SMALL
#define GPC_TAIL_GROUPS_SMALL 128
#endif
constexpr int kTailSlotsBig = (2 * kOvCap > kTailCap) ? 2 * kOvCap : kTailCap, kTailGroupsBig = 256;
constexpr int kTailSlotsSmall = kTailCap, kTailGroupsSmall = GPC_TAIL_GROUPS_SMALL;
constexpr size_t tail_smem_bytes(int slots, int groups) { return (size_t)kTailWarps * (2 * groups + 2 * slots) * 4; }

template <int kSlots, int kGroups>
__global__ void __launch_bounds__(32 * kTailWarps, kSlots <= kTailSlotsSmall ? GPC_TAIL_MINB_SMALL : 4)
match_rows_tail_kernel(const MatchArgs args) {
  static_assert(kSlots >= kTailCap && kGroups >= 32 && kGroups % 128 == 0 && (kGroups & (kGroups - 1)) == 0, "slice layout");
  constexpr int kTailWords = 2 * kGroups + 2 * kSlots;
  constexpr int kGroupShift = 32 - (kGroups == 128 ? 7 : kGroups == 256 ? 8 : 9);
  static_assert(kGroups == 128 || kGroups == 256 || kGroups == 512, "group bits");
  extern __shared__ __align__(16) uint32_t tail_smem[];
  __shared__ uint32_t sm_m[kTailWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t* cnt = tail_smem + (size_t)wid * kTailWords;             // [kGroups]
  uint32_t* cur = cnt + kGroups;                                     // [kGroups]
  uint32_t* bk = cur + kGroups;                                      // [kSlots] keys
  uint32_t* bv = bk + kSlots;                                        // [kSlots] payloads
  // grid (row groups, pairs): no division, row indices in 32 bits
  const int pair = blockIdx.y, y = kRadius + (int)blockIdx.x * kTailWarps + wid;
  if (y >= args.H - kRadius) return;
  const uint32_t grow = (uint32_t)pair * (uint32_t)args.H + (uint32_t)y;
  const int4 hdr = *reinterpret_cast<const int4*>(args.rowhdr + (size_t)grow * 4u);
  if (hdr.w != 0) return;                                   // empty row, or one the general kernel owns
  unsigned long long* mrec = args.mrec + (size_t)grow * (uint32_t)args.W;
  uint32_t m = (uint32_t)hdr.x;
  const uint32_t nl = (uint32_t)hdr.y, nr = (uint32_t)hdr.z;
  // ---- overflow entries: a hash-partitioned join (linear in their number) ------------------------------------
  // group = 8 bits of a multiplicative hash of the state; both lists are counted, scanned and scattered into group
  // segments (left entries carry bit 31 clear, right entries set), then every left entry walks its own group
  if (nl + nr > (uint32_t)kSlots) {                         // more overflow entries than a warp's slice holds: general kernel
    if (lane == 0) push_row(args.fb_hdr, args.fb_ent, pair, y);
    return;
  }
  if (nl > 0u && nr > 0u) {
    const uint32_t* ovl_s = args.ovbuf + (size_t)grow * (uint32_t)(4 * kOvCap);
    const uint32_t* ovl_x = ovl_s + kOvCap;
    const uint32_t* ovr_s = ovl_x + kOvCap;
    const uint32_t* ovr_x = ovr_s + kOvCap;
    const uint32_t n = nl + nr;
#pragma unroll
    for (int k = 0; k < kGroups / 32 / 4; k++) reinterpret_cast<uint4*>(cnt)[lane + 32 * k] = make_uint4(0, 0, 0, 0);
    if (lane == 0) sm_m[wid] = m;
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) {
      const uint32_t v = (i < nl) ? ovl_s[i] : ovr_s[i - nl];
      atomicAdd(&cnt[((v & 0x7fffffffu) * kHashMul) >> kGroupShift], 1u);
    }
    __syncwarp();
    {                                                       // exclusive scan of the group counters, kGroups / 32 per lane
      constexpr int kPer4 = kGroups / 128;                  // uint4s per lane
      uint4 c4[kPer4];
      uint32_t sum = 0;
#pragma unroll
      for (int k = 0; k < kPer4; k++) {
        c4[k] = reinterpret_cast<const uint4*>(cnt)[kPer4 * lane + k];
        sum += c4[k].x + c4[k].y + c4[k].z + c4[k].w;
      }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      uint32_t run = incl - sum;
#pragma unroll
      for (int k = 0; k < kPer4; k++) {
        uint4 o;
        o.x = run; run += c4[k].x; o.y = run; run += c4[k].y; o.z = run; run += c4[k].z; o.w = run; run += c4[k].w;
        reinterpret_cast<uint4*>(cur)[kPer4 * lane + k] = o;
      }
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) {               // scatter; cur[g] ends at the group's end
      const bool left = i < nl;
      const uint32_t v = (left ? ovl_s[i] : ovr_s[i - nl]) & 0x7fffffffu;
      const uint32_t x = left ? ovl_x[i] : ovr_x[i - nl];
      const uint32_t p = atomicAdd(&cur[(v * kHashMul) >> kGroupShift], 1u);
      bk[p] = left ? v : (v | 0x80000000u);
      bv[p] = x;
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) {               // every left entry: equal states in its group, per side
      const uint32_t key = bk[i];
      if (key >> 31) continue;
      const uint32_t g = (key * kHashMul) >> kGroupShift;
      const uint32_t gn = cnt[g], g0 = cur[g] - gn;
      uint32_t cl = 0, cr = 0, x2 = 0;
      for (uint32_t j = 0; j < gn; j++) {
        const uint32_t t = bk[g0 + j] ^ key;                // 0: same state, left; 0x80000000: same state, right
        cl += (t == 0u) ? 1u : 0u;
        if (t == 0x80000000u) { cr++; x2 = bv[g0 + j]; }
      }
      if (cl == 1u && cr == 1u) {
        const uint32_t xl = bv[i];
        const int dx = (int)xl - (int)x2;
        if (dx <= args.disp_high && -dx <= args.disp_high && 0 <= args.vertical_tolerance) {
          const uint32_t p = atom_add_shared(&sm_m[wid], 1u);
          GPC_CHECK(p < (uint32_t)args.W);
          mrec[p] = ((unsigned long long)key << 32) | (unsigned long long)((xl << 16) | x2);
        }
      }
    }
    __syncwarp();
    m = sm_m[wid];
    __syncwarp();                                         // the arrays are reused below
  }
  if (lane == 0) args.rowmatch[grow] = (int32_t)m;
  if (m == 0u) return;
  // ---- order by state and stage -------------------------------------------------------------------------
  uint32_t* stage = args.stage + (size_t)grow * (uint32_t)args.W;
  if (m > (uint32_t)kTailCap) {                           // a crowd: the block-wide ordering kernel
    if (lane == 0) push_row(args.big_hdr, args.big_ent, pair, y);
    return;
  }
  const int shift = args.key_bits > 5 ? args.key_bits - 5 : 0;        // bucket = top 5 state bits = the lane that owns it
  // the row's match records (the ones this warp appended are visible after the __syncwarp above)
  unsigned long long rec[kTailCap / 32];
#pragma unroll
  for (int j = 0; j < kTailCap / 32; j++) {
    const uint32_t i = (uint32_t)lane + 32u * j;
    rec[j] = (i < m) ? mrec[i] : 0ull;
  }
  cnt[lane] = 0u;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < kTailCap / 32; j++) {
    const uint32_t i = (uint32_t)lane + 32u * j;
    if (i < m) atomicAdd(&cnt[(uint32_t)(rec[j] >> 32) >> shift], 1u);
  }
  __syncwarp();
  {                                                       // exclusive scan of the 32 counters, one per lane
    const uint32_t c = cnt[lane];
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    cur[lane] = incl - c;
    if (__any_sync(0xffffffffu, c > 32u)) {               // skewed states: the block-wide kernel has finer buckets and a sorting network
      if (lane == 0) push_row(args.big_hdr, args.big_ent, pair, y);
      return;
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < kTailCap / 32; j++) {               // scatter into bucket segments; cur[b] ends at the bucket's end
    const uint32_t i = (uint32_t)lane + 32u * j;
    if (i < m) {
      const uint32_t key = (uint32_t)(rec[j] >> 32);
      const uint32_t p = atomicAdd(&cur[key >> shift], 1u);
      bk[p] = key; bv[p] = (uint32_t)rec[j];
    }
  }
  __syncwarp();
  for (uint32_t i = lane; i < m; i += 32) {               // rank inside the bucket, write to the final position
    const uint32_t key = bk[i];
    const uint32_t b = key >> shift;
    const uint32_t n = cnt[b], s0 = cur[b] - n;
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; j++) rank += (bk[s0 + j] < key) ? 1u : 0u;
    stage[s0 + rank] = bv[i];
  }
}

// Block-wide ordering of the rows the tail kernel listed (more than kTailCap matches, or skewed states): 256 threads
// whatever the row width -- a row has at most a few hundred matches, wider CTAs only add idle warps to every barrier.
constexpr int kOrderThreads = 256;
template <int kThreadsB>
__global__ void __launch_bounds__(kThreadsB)
order_rows_kernel(const MatchArgs args, uint32_t* hdr, const uint32_t* ent) {
  extern __shared__ __align__(16) uint8_t smem[];
  unsigned long long* out = reinterpret_cast<unsigned long long*>(smem);           // [pow2cap] x2, then bcnt / bstart
  unsigned long long* out2 = out + args.pow2cap;
  uint32_t* bcnt = reinterpret_cast<uint32_t*>(out2 + args.pow2cap);
  uint32_t* bstart = bcnt + kBuckets;
  __shared__ uint32_t big_bucket;
  const int tid = threadIdx.x;
  const uint32_t n = *reinterpret_cast<volatile uint32_t*>(hdr);
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const int pair = (int)ent[2 * i], y = (int)ent[2 * i + 1];
    const size_t grow = (size_t)pair * args.H + y;
    const int m = args.rowmatch[grow];
    const unsigned long long* mrec = args.mrec + grow * args.W;
    for (int k = tid; k < m; k += kThreadsB) out[k] = mrec[k];
    for (int k = tid; k < kBuckets; k += kThreadsB) bcnt[k] = 0u;
    if (tid == 0) big_bucket = 0;
    __syncthreads();
    order_and_stage<kThreadsB>(args, out, out2, bcnt, bstart, &big_bucket, m, pair, y);
    __syncthreads();
  }
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&hdr[1], 1u) == gridDim.x - 1) { hdr[0] = 0u; hdr[1] = 0u; }
  }
}

// Launch shapes: every thread owns 4 * KQ pixels per side.  Rows up to 1024 pixels run 256 threads,
// wider rows 512 or 1024 threads with KQ = 1 (wide rows need large tables, so few CTAs fit an SM and
// each has to bring more warps); beyond 4096 pixels KQ grows.
template <int KQ, int T>
static cudaError_t configure_one(int max_smem) {
  cudaError_t e = cudaFuncSetAttribute(match_rows_general_kernel<KQ, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(match_rows_fast_kernel<KQ, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  if (e == cudaSuccess && KQ == 1 && T == kOrderThreads) e = cudaFuncSetAttribute(order_rows_kernel<kOrderThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  return e;
}

cudaError_t configure_match_rows(int max_smem) {
  cudaError_t e = cudaFuncSetAttribute(match_rows_tail_kernel<kTailSlotsBig, kTailGroupsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tail_smem_bytes(kTailSlotsBig, kTailGroupsBig));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(match_rows_tail_kernel<kTailSlotsSmall, kTailGroupsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)tail_smem_bytes(kTailSlotsSmall, kTailGroupsSmall));
  if (e == cudaSuccess) e = configure_one<1, 256>(max_smem);
  if (e == cudaSuccess) e = configure_one<1, 512>(max_smem);
  if (e == cudaSuccess) e = configure_one<1, 1024>(max_smem);
  if (e == cudaSuccess) e = configure_one<2, 1024>(max_smem);
  return e;
}

int match_rows_threads(int W) {
  const int quads = W / 4;
  return quads <= 256 ? 256 : quads <= 512 ? 512 : 1024;
}

// general != 0: every row through the general kernel (GPC_MATCHER_ROWS_GENERAL).  Otherwise the fast kernel
// over all rows, then the general kernel over the row list it left (args.fb_list, sized for every row).
cudaError_t launch_match_rows(const MatchArgs& args, int n_pairs, int general, int sm_count, cudaStream_t stream) {
  int rows = args.H - 2 * kRadius;
  if (rows <= 0 || n_pairs <= 0) return cudaSuccess;
  const int quads = args.W / 4;
  if (quads > 2048) return cudaErrorInvalidValue;
  dim3 grid(rows, n_pairs);
  const size_t smem_g = match_smem_bytes(args.wcap, args.table_log2);
  const size_t smem_f = match_fast_smem_bytes(args.nib_log2, args.slot_log2, args.pow2cap);
  const size_t smem_o = order_rows_smem_bytes(args.pow2cap);
  const long long all_rows = (long long)rows * n_pairs;
  const int list_grid = (int)(all_rows < 4ll * sm_count ? all_rows : 4ll * sm_count);
  const dim3 tail_grid((rows + kTailWarps - 1) / kTailWarps, n_pairs);
  const int order_grid = (int)(all_rows < 6ll * sm_count ? all_rows : 6ll * sm_count);
  static const int tail_env = std::getenv("GPC_B_TAIL") ? std::atoi(std::getenv("GPC_B_TAIL")) : 0;   // 1 small, 2 big slice
  const bool small_tail = tail_env ? tail_env == 1 : quads <= 256;
#define GPC_LAUNCH_ROWS(KQ, T)                                                                                  \
  do {                                                                                                          \
    if (general) match_rows_general_kernel<KQ, T><<<grid, T, smem_g, stream>>>(args, nullptr, nullptr);                  \
    else {                                                                                                      \
      match_rows_fast_kernel<KQ, T><<<grid, T, smem_f, stream>>>(args);                                         \
      if (small_tail) match_rows_tail_kernel<kTailSlotsSmall, kTailGroupsSmall><<<tail_grid, 32 * kTailWarps, tail_smem_bytes(kTailSlotsSmall, kTailGroupsSmall), stream>>>(args); \
      else match_rows_tail_kernel<kTailSlotsBig, kTailGroupsBig><<<tail_grid, 32 * kTailWarps, tail_smem_bytes(kTailSlotsBig, kTailGroupsBig), stream>>>(args); \
      order_rows_kernel<kOrderThreads><<<order_grid, kOrderThreads, smem_o, stream>>>(args, args.big_hdr, args.big_ent);                         \
      match_rows_general_kernel<KQ, T><<<list_grid, T, smem_g, stream>>>(args, args.fb_hdr, args.fb_ent);                 \
    }                                                                                                           \
  } while (0)
  if (quads <= 256) GPC_LAUNCH_ROWS(1, 256);
  else if (quads <= 512) GPC_LAUNCH_ROWS(1, 512);
  else if (quads <= 1024) GPC_LAUNCH_ROWS(1, 1024);
  else GPC_LAUNCH_ROWS(2, 1024);
#undef GPC_LAUNCH_ROWS
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

// Exclusive prefix of n int32 values into long long out[0 .. n] (out[n] = total): one CTA of 1024 threads walks the
// input in 1024-element steps (warp shuffles + one shared-memory hop), so batches of thousands of pairs and images
// of thousands of rows cost a few microseconds instead of a serial loop.
template <typename OutT>
__device__ __forceinline__ void block_exclusive_scan(const int32_t* __restrict__ in, int n, OutT* __restrict__ out) {
  __shared__ long long warp_sums[32];
  __shared__ long long carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const long long v = (i < n) ? (long long)in[i] : 0ll;
    long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const long long w = warp_sums[lane];
      long long wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    const long long base = carry + warp_sums[wid];
    if (i < n) out[i] = (OutT)(base + incl - v);
    __syncthreads();
    if (tid == 1023) carry = base + incl;
    __syncthreads();
  }
  if (tid == 0) out[n] = (OutT)carry;
}

// Exclusive prefix of the per-pair totals (packed output mode).
__global__ void __launch_bounds__(1024)
pair_scan_kernel(const int32_t* __restrict__ totals, int n_pairs, long long* __restrict__ pair_base) {
  block_exclusive_scan<long long>(totals, n_pairs, pair_base);
}

cudaError_t launch_pair_scan(const int32_t* totals, int n_pairs, long long* pair_base, cudaStream_t stream) {
  pair_scan_kernel<<<1, 1024, 0, stream>>>(totals, n_pairs, pair_base);
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

// One warp per row.  The row's length, its two offsets and its first 128 staged matches are requested TOGETHER (the
// kernel is a chain of dependent global loads per warp otherwise -- ncu: 39 warps stalled on long_scoreboard per issue
// cycle at 26 % issue slots); the stage words are the row's own (zero-filled at context creation, whatever lies beyond
// the row's length is ignored), so reading them before the length is known is harmless.
__device__ __forceinline__ uint32_t ld_nc_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));      // volatile: stays where it is written
  return v;
}
__device__ __forceinline__ long long ld_nc_s64(const void* p) {
  long long v;
  asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(32 * kEmitRows)
emit_supports_kernel(const EmitArgs a) {
  const int y = kRadius + blockIdx.x * kEmitRows + (threadIdx.x >> 5), pair = blockIdx.y, lane = threadIdx.x & 31;
  if (y >= a.H - kRadius) return;
  const size_t grow = (size_t)pair * a.H + y;
  const uint32_t* stage = a.stage + grow * a.W;
  constexpr int kAhead = 4;                               // 128 matches: more than a Sintel row holds on average
  const int m = (int)ld_nc_u32(a.rowmatch + grow);
  const long long off = (long long)(int)ld_nc_u32(a.rowoff + grow);
  long long base, limit;
  if (a.pair_base) { base = ld_nc_s64(a.pair_base + pair); limit = a.cap; }
  else { base = (long long)pair * a.cap; limit = base + a.cap; }
  uint32_t ahead[kAhead];
#pragma unroll
  for (int j = 0; j < kAhead; j++) ahead[j] = (lane + 32 * j < a.W) ? ld_nc_u32(stage + lane + 32 * j) : 0u;
  // no early exit for empty rows: ptxas would sink the loads it does not need for the decision below the branch
  float* const out = a.out + 3 * (base + off);
  const long long room = limit - (base + off);            // records of this row that fit the output
  auto put = [&](int k, uint32_t u) {
    const int xl = (int)(u >> 16), xr = (int)(u & 0xffffu);
    float* o = out + 3 * (long long)k;
    o[0] = __int_as_float(xl);
    o[1] = __int_as_float(y);
    o[2] = (float)(xl - xr);                 // Support::d = float(xL - xR), exact (inference.hpp:389-390)
  };
#pragma unroll
  for (int j = 0; j < kAhead; j++) {
    const int k = lane + 32 * j;
    if (k < m && k < room) put(k, ahead[j]);
  }
  // longer rows (dense images): four more loads in flight per round instead of a load -> store chain per match
  for (int k0 = lane + 32 * kAhead; k0 < m && k0 < room; k0 += 32 * kAhead) {
    uint32_t more[kAhead];
#pragma unroll
    for (int j = 0; j < kAhead; j++) more[j] = (k0 + 32 * j < m) ? ld_nc_u32(stage + k0 + 32 * j) : 0u;
#pragma unroll
    for (int j = 0; j < kAhead; j++) {
      const int k = k0 + 32 * j;
      if (k < m && k < room) put(k, more[j]);
    }
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
__global__ void __launch_bounds__(1024)
mask_scan_kernel(const int32_t* __restrict__ rowcnt, int H, int32_t* __restrict__ rowoff) {
  block_exclusive_scan<int32_t>(rowcnt, H, rowoff);
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
  mask_scan_kernel<<<1, 1024, 0, stream>>>(rowcnt, H, rowoff);
  mask_rows_kernel<<<H, 32, 0, stream>>>(hash, rowoff, W, H, mask, cap);
  return cudaGetLastError();
}

}  // namespace gpc

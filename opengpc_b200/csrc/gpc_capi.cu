// gpc_capi.cu -- the C ABI of include/gpc_b200.h: resident per-GPU context, forest upload,
// batched launch sequence, host<->device staging.  No CPU fallback: every compute entry point
// runs the CUDA kernels of smooth_sobel.cu / hash_tiles.cu / match_rows.cu / match_global.cu or fails with GPC_E_CUDA.
#include "../../include/gpc_b200.h"
#include "gpc_device.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

namespace gpc {
cudaError_t launch_smooth_sobel(const PreprocessArgs&, int n_img, bool debug_out, const void* raw_tmap, cudaStream_t);
void smooth_sobel_tma_box(int* box_w, int* box_h);
int make_u8_tensor_map(void* out_map, const uint8_t* base, int W, int H, int n_img, int box_w, int box_h);
cudaError_t launch_prep_from_smooth(const uint8_t* smooth, const uint8_t* flags, uint8_t* smooth_x, uint16_t* cand, int32_t* rowcnt,
                                    int32_t* lastrow, int W, int H, int naive, cudaStream_t);
cudaError_t configure_hash_tiles();
int make_smooth_tensor_map(void* out_map, const uint8_t* base, int W, int H, int n_img);
cudaError_t launch_hash_tiles(const void* tensor_map, const HashArgs&, const ForestDev&, int n_img, cudaStream_t);
size_t match_smem_bytes(int wcap, int table_log2);
size_t match_fast_smem_bytes(int nib_log2, int slot_log2, int pow2cap);
size_t order_rows_smem_bytes(int pow2cap);
int match_ov_cap();
cudaError_t configure_match_rows(int max_smem);
cudaError_t configure_match_global();
cudaError_t launch_match_rows(const MatchArgs&, int n_pairs, int general, int sm_count, cudaStream_t);
int match_rows_threads(int W);
cudaError_t launch_row_scan(const int32_t* rowmatch, const int32_t* rowcnt, int H, int n_pairs, int32_t* rowoff,
                            int32_t* totals, int32_t* n_cand, cudaStream_t);
cudaError_t launch_pair_scan(const int32_t* totals, int n_pairs, long long* pair_base, cudaStream_t);
cudaError_t launch_emit_supports(const uint32_t* stage, const int32_t* rowmatch, const int32_t* rowoff,
                                 const long long* pair_base, void* out, long long cap, int W, int H, int n_pairs,
                                 cudaStream_t);
cudaError_t launch_mask_list(const uint32_t* hash, const int32_t* rowcnt, int32_t* rowoff, int W, int H, int32_t* mask,
                             int cap, cudaStream_t);
size_t global_workspace_bytes(long long max_records, int n_pairs, int H, int W);
cudaError_t launch_match_global(const uint32_t* hash, const int32_t* rowcnt, int W, int H, int n_pairs, int epipolar, int key_bits,
                                int disp_high, int vertical_tolerance, int mode, void* ws, long long max_records, void* out,
                                long long out_stride, long long cap, int32_t* n_out, int32_t* n_cand, cudaStream_t stream,
                                int* launches, int hashtable, long long* pair_base, int first_chunk);
cudaError_t launch_hashmatch_keys(void* ws, long long max_records, const unsigned long long* d_keys64, int ns, int nt, int32_t* out_pairs,
                                  long long cap, int32_t* n_out, cudaStream_t stream, int* launches);
cudaError_t launch_match_keys(void* ws, long long max_records, int ns, int nt, int key_bits, int32_t* out_pairs, long long cap,
                              int32_t* n_out, cudaStream_t stream, int* launches);
void* global_key_buffer(void* ws, long long max_records);
size_t wide_workspace_bytes(long long max_records);
cudaError_t launch_wide_rank(void* ws, long long max_records, const uint32_t* hi, const uint32_t* lo, long long n_pix, int key_bits,
                             uint32_t* out, cudaStream_t stream, int* launches);
cudaError_t launch_downsample2x(const uint8_t* src, uint8_t* dst, int sw, int sh, int n_img, cudaStream_t stream);
struct JitKernel;
JitKernel* jit_build_hash_tiles(const ForestDev& f, std::string* why);
void jit_destroy(JitKernel* k);
cudaError_t jit_launch_hash_tiles(JitKernel* k, const void* tensor_map, const HashArgs& args, const ForestDev& forest, int n_img,
                                  cudaStream_t stream);
}  // namespace gpc

static thread_local std::string g_create_error;

// live contexts by serial: gpc_image_release gives an image's block back to its context if that still exists
static std::mutex g_registry_mu;
static std::map<long long, gpc_ctx*> g_registry;

struct gpc_ctx {
  long long id = 0;
  int device = 0;
  int max_w = 0, max_h = 0, max_batch = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  bool has_forest = false;
  long long forest_serial = 0;     // bumped whenever forest_dev is re-baked (resident images key their hash cache on it)
  gpc_forest forest_host{};
  int result_mode = GPC_RESULTS_SSE;   // gpc_set_result_mode
  gpc::ForestDev forest_dev{};
  gpc::JitKernel* jit = nullptr;   // kernel A2 specialised for forest_dev (NVRTC), or nullptr -> generic kernel
  std::string jit_note = "no forest set";
  // resident device buffers
  uint8_t* d_raw = nullptr;        // [2B][H][W]
  uint8_t* d_smooth = nullptr;     // [2B][H][W]  biased smoothed images (kernel A1 -> TMA -> kernel A2)
  uint16_t* d_cand = nullptr;      // [2B][H][W/16] candidate masks
  alignas(64) unsigned char tmap[128];   // CUtensorMap over d_smooth for images of tmap_w x tmap_h
  int tmap_w = 0, tmap_h = 0;
  // tensor maps over RAW images for kernel A1, keyed by (pointer, w, h, images): a few recent ones (the chunks of a
  // pipelined batch alternate between slices of d_raw, callers of the device-resident entry bring their own buffers)
  struct RawMap { alignas(64) unsigned char map[128]; const uint8_t* ptr = nullptr; int w = 0, h = 0, n = 0; };
  RawMap raw_maps[8];
  int raw_map_next = 0;
  bool a1_tma = true;              // GPC_A1_TMA=0: always the staging-loop kernel
  uint32_t* d_hash = nullptr;      // [2B][H][W]
  uint32_t* d_stage = nullptr;     // [B][H][W]
  int32_t* d_rows = nullptr;       // rowcnt [2B][H]   (cleared per launch)
  int32_t* d_lastrow = nullptr;    // [2B]             (cleared per launch)
  // gpc_match_batch pipelines chunks of pairs over these lanes (upload / kernels / download); every
  // chunk works on its own slice of the resident buffers
  static constexpr int kLanes = 3;
  cudaStream_t lane_stream[kLanes] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_chunk;
  int chunk_pairs = 0;             // GPC_CHUNK_PAIRS; 0 = chosen per call (chunk_for)
  int32_t* d_rowmatch = nullptr;   // [B][H]
  uint32_t* d_fb = nullptr;        // [B][2 * max_h + 2] row lists of the fast row matcher (one per slot, at p0 * stride)
  uint32_t* d_big = nullptr;       // same, rows for the block-wide ordering kernel
  unsigned long long* d_mrec = nullptr;   // [B][H][W] unordered match records (fast row matcher -> tail kernel)
  uint32_t* d_ovbuf = nullptr;     // [B][H][4][ov_cap] overflow lists of the rows
  int32_t* d_rowhdr = nullptr;     // [B][H][4]
  int sm_count = 148;
  int32_t* d_rowoff = nullptr;     // [B][H+1]
  int32_t* d_totals = nullptr;     // [B]
  int32_t* d_ncand = nullptr;      // [B][2]
  long long* d_pair_base = nullptr;  // [B+1]
  gpc_support* d_out = nullptr;    // [out_cap]
  uint8_t* d_dbg8 = nullptr;       // smooth | grad, debug entry points (lazily allocated)
  int32_t* d_mask = nullptr;       // candidate list, debug entry points (lazily allocated)
  long long out_cap = 0;
  // pinned host scratch for counts
  int32_t* h_counts = nullptr;     // totals [B] | ncand [2B]
  long long* h_pair_base = nullptr;  // [B+1]
  uint8_t* d_pyr = nullptr;        // pyramid levels 1.. of one pair (lazily allocated)
  void* d_gws = nullptr;           // radix-sort matcher workspace (lazily allocated, grown on demand)
  size_t gws_bytes = 0;
  // wide states (forests of more than 32 tests, gpc_set_wide_forest): one baked sub-forest per state word
  std::vector<gpc::ForestDev> wide_words;
  int wide_tests = 0;
  uint32_t* d_wide = nullptr;      // hash planes [2 * words][2][H][W] (lazily allocated)
  size_t wide_bytes = 0;
  int matcher = GPC_MATCHER_AUTO;
  // page-locked staging for the single-image entry points (uploads from / downloads to pageable caller memory are
  // several times slower than a memcpy through this buffer); ev_stage marks the last asynchronous use
  uint8_t* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  cudaEvent_t ev_stage = nullptr;
  int last_mask_n = 0;             // entries of the candidate list left in h_stage by gpc_image_preprocess (mask == NULL)
  // device blocks of released resident images, reused by gpc_image_upload (cudaMalloc / cudaFree cost more than the kernels)
  std::vector<std::pair<size_t, uint8_t*>> image_pool;
  int64_t launches = 0;
  int match_smem_max = 0;
  // optional per-kernel timing (CUDA events on the launching stream), see gpc_kernel_times
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  double k_ms[GPC_N_KERNELS] = {0, 0, 0, 0, 0};
  int64_t k_runs = 0;
  std::string err;
};

// A raw image resident on a context's device (Forest::PreprocessedImage's device side).
struct gpc_image {
  long long ctx_id = 0;       // serial of the owning context (a context pointer may be reused)
  gpc_ctx* owner = nullptr;   // only dereferenced after ctx_id was found in the registry of live contexts
  int device = 0;
  uint8_t* d_raw = nullptr;   // [h][w], tightly packed; start of the image's single device block
  size_t block_bytes = 0;
  int w = 0, h = 0;
  // outputs of gpc_image_preprocess kept for gpc_match_images / gpc_correspond_images, in the same block after the
  // raw image: biased smoothed image [h][w] | candidate masks u16 [h][w/16] | rowcnt i32 [h] + lastrow i32 | hash u32 [h][w]
  uint8_t* d_cache = nullptr;
  int cache_thr = -1, cache_mode = -1;     // gradient threshold / result mode the cache was computed with (-1: none)
  long long hash_serial = 0;               // forest serial the cached hash image belongs to (0: none)
  size_t off_cand() const { return ((size_t)w * h + 255) / 256 * 256; }
  size_t off_rows() const { return off_cand() + ((size_t)h * (w / 16) * 2 + 255) / 256 * 256; }
  size_t off_hash() const { return off_rows() + ((size_t)(h + 1) * 4 + 255) / 256 * 256; }
  size_t cache_bytes() const { return off_hash() + (size_t)w * h * 4; }
};

namespace {

int fail(gpc_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

#define GPC_CUDA(ctx, call)                                                                       \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail((ctx), GPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
  } while (0)

// Record a timing event on the context's stream (no-op unless gpc_enable_kernel_timing).
int mark(gpc_ctx* c) {
  if (!c->timing) return GPC_OK;
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e;
    GPC_CUDA(c, cudaEventCreate(&e));
    c->ev_pool.push_back(e);
  }
  GPC_CUDA(c, cudaEventRecord(c->ev_pool[c->ev_used++], c->stream));
  return GPC_OK;
}

// Bake the forest for the kernel's shared-memory layout (the reference bakes it for the image
// width instead, inference.hpp:427-428): see ForestDev.
void bake_forest(const gpc_forest& f, gpc::ForestDev* d, int result_mode) {
  std::memset(d, 0, sizeof(*d));
  d->n_tests = f.n_tests;
  d->type = f.type;
  d->naive = (result_mode == GPC_RESULTS_NAIVE) ? gpc::kResultsNaive : gpc::kResultsSse;
  auto imm = [](int dx, int dy) {
    const int o = dy * gpc::kPitch + dx;
    const int k = ((o % 4) + 4) % 4;
    return k * gpc::kCopyBytes + (o - k);
  };
  // test slots the forest does not fill compare a pixel with itself: never true, state bit 0 (kernel A evaluates
  // whole groups of tests without per-test guards); imm 0 = the quad's own word in copy 0
  // (bit placement of filter.hpp:574-584 -- test t < 8 -> bit t, test 8 OR-ed into bit 0, t >= 9 -> bit t - 1 -- is
  // compiled into the kernel's unrolled test loop)
  if (d->naive)
    for (int t = 0; t < gpc::kMaxTests; t++) d->mtau2[t] = 0x80008000u;      // 32768 - 0 in both lanes: "a + 0 > a" is false
  if (d->naive) {
    // gpcFilterNaive / gpcFilterTauNaive (filter.hpp:245-293): test t of T lands in bit T-1-t.  The kernel's slot s
    // feeds state bit s (s < 8) or s - 1 (s > 8); slot 8 (the SSE build's ninth test, OR-ed into bit 0) stays empty.
    const int T = f.n_tests;
    for (int t = 0; t < T; t++) {
      const int bit = T - 1 - t, slot = (bit < 8) ? bit : bit + 1;
      d->imm_a[slot] = imm(f.ix[t], f.iy[t]);
      d->imm_b[slot] = imm(f.jx[t], f.jy[t]);
      const int tau = std::max(-256, std::min(256, (int)f.tau[t]));          // beyond +-255 the comparison is constant
      const uint32_t k = (uint32_t)(32768 - (f.type == 1 ? tau : 0));
      d->mtau2[slot] = k | (k << 16);
    }
    d->n_tests = (T <= 8) ? T : T + 1;                       // highest slot in use + 1
    return;
  }
  for (int t = 0; t < f.n_tests; t++) {
    d->imm_a[t] = imm(f.ix[t], f.iy[t]);
    d->imm_b[t] = imm(f.jx[t], f.jy[t]);
    int tau8 = (int)(int8_t)f.tau[t];                        // _mm_set1_epi8(tau): low 8 bits, signed
    uint32_t m = (uint32_t)(uint16_t)(int16_t)(-tau8);
    d->mtau2[t] = (f.type == 1) ? (m | (m << 16)) : 0u;
  }
}

int check_dims(gpc_ctx* c, int w, int h, int n_pairs) {
  if (w <= 0 || h <= 0 || n_pairs <= 0) return fail(c, GPC_E_ARG, "non-positive dimension");
  if (w % 16 != 0) return fail(c, GPC_E_WIDTH16, "width must be multiple of 16!");   // filter.hpp:294
  if (w > 8192) return fail(c, GPC_E_DIMS, "width above 8192 is not supported by the row matcher");
  if (w > c->max_w || h > c->max_h || (long long)w * h > (long long)c->max_w * c->max_h || n_pairs > c->max_batch)
    return fail(c, GPC_E_DIMS, "image or batch exceeds the context's capacity");
  return GPC_OK;
}

int check_settings(gpc_ctx* c, const gpc_settings* s) {
  if (!s) return fail(c, GPC_E_ARG, "settings is NULL");
  if (s->gradient_threshold < 0 || s->gradient_threshold > 255)
    return fail(c, GPC_E_ARG, "gradientThreshold needs to be within 0...255");     // inference.hpp:303
  return GPC_OK;
}

int ceil_log2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

// log2 of the row matcher's bucket count: about half a candidate (left + right) per bucket, at
// least 4 per thread (scan layout) and at least 2 * W (so that remainder + side + x fit one 32-bit entry)
int table_log2_for(int w, int wcap) {
  int l = std::max(ceil_log2(4 * gpc::match_rows_threads(w)), ceil_log2(w) + 1);
  static const int mult = std::getenv("GPC_B_BUCKET_MULT") ? std::max(1, std::atoi(std::getenv("GPC_B_BUCKET_MULT"))) : 4;
  while ((1 << l) < mult * wcap) l++;
  return l;
}

// A slice of the context's resident buffers (pairs p0 .. p0 + n) and the stream that works on it.
struct Slot {
  int p0;
  cudaStream_t stream;
};

int mark_on(gpc_ctx* c, const Slot& sl) { return (sl.stream == c->stream) ? mark(c) : GPC_OK; }

// Kernels A1 + A2 over n_img resident images of the slot; clears and fills rowcnt / lastrow.
// d_flags != nullptr: d_images holds SMOOTHED images and d_flags the pixels to hash (gpc_hash_smooth).
// skip_a1: the slot's smooth / cand / rowcnt / lastrow buffers already hold kernel A1's output (resident images).
int run_preprocess(gpc_ctx* c, const Slot& sl, const uint8_t* d_images, int n_img, int w, int h, int thr,
                   const gpc::ForestDev& forest, uint8_t* d_smooth_out, uint8_t* d_grad_out, const uint8_t* d_flags = nullptr,
                   bool skip_a1 = false, uint32_t* hash_base = nullptr) {
  const size_t P = (size_t)w * h;
  const int img0 = 2 * sl.p0;
  int32_t* rowcnt = c->d_rows + (size_t)img0 * h;
  int32_t* lastrow = c->d_lastrow + img0;
  if (!skip_a1) {
    GPC_CUDA(c, cudaMemsetAsync(rowcnt, 0, (size_t)n_img * h * sizeof(int32_t), sl.stream));
    GPC_CUDA(c, cudaMemsetAsync(lastrow, 0xff, (size_t)n_img * sizeof(int32_t), sl.stream));   // -1
  }
  if (c->tmap_w != w || c->tmap_h != h) {
    const int n_cap = (int)std::min<size_t>(2 * (size_t)c->max_batch * ((size_t)c->max_w * c->max_h / P), 1u << 30);
    if (gpc::make_smooth_tensor_map(c->tmap, c->d_smooth, w, h, n_cap) != 0)
      return fail(c, GPC_E_CUDA, "cuTensorMapEncodeTiled failed");
    c->tmap_w = w; c->tmap_h = h;
  }
  uint8_t* smooth_x = c->d_smooth + (size_t)img0 * P;
  uint16_t* cand = c->d_cand + (size_t)img0 * h * (w / 16);
  int rc = mark_on(c, sl); if (rc) return rc;                                      // event 0
  if (skip_a1) {
  } else if (d_flags) {
    if (n_img != 1) return fail(c, GPC_E_ARG, "internal: the smooth seam handles one image");
    GPC_CUDA(c, gpc::launch_prep_from_smooth(d_images, d_flags, smooth_x, cand, rowcnt, lastrow, w, h, forest.naive, sl.stream));
  } else {
    gpc::PreprocessArgs a{};
    a.raw = d_images; a.smooth_x = smooth_x; a.cand = cand; a.rowcnt = rowcnt; a.lastrow = lastrow;
    a.smooth_out = d_smooth_out; a.grad_out = d_grad_out; a.W = w; a.H = h;
    a.naive = forest.naive;
    a.thr2 = forest.naive ? thr * thr : (int32_t)(int16_t)(thr * thr);             // filter.hpp:159 / :418
    const void* raw_tmap = nullptr;
    if (c->a1_tma && !forest.naive && (reinterpret_cast<uintptr_t>(d_images) & 15u) == 0) {
      gpc_ctx::RawMap* hit = nullptr;
      for (gpc_ctx::RawMap& m : c->raw_maps)
        if (m.ptr == d_images && m.w == w && m.h == h && m.n == n_img) { hit = &m; break; }
      if (!hit) {
        hit = &c->raw_maps[c->raw_map_next];
        c->raw_map_next = (c->raw_map_next + 1) % 8;
        int bw = 0, bh = 0;
        gpc::smooth_sobel_tma_box(&bw, &bh);
        if (gpc::make_u8_tensor_map(hit->map, d_images, w, h, n_img, bw, bh) == 0) { hit->ptr = d_images; hit->w = w; hit->h = h; hit->n = n_img; }
        else { hit->ptr = nullptr; hit = nullptr; }
      }
      if (hit) raw_tmap = hit->map;
    }
    GPC_CUDA(c, gpc::launch_smooth_sobel(a, n_img, d_smooth_out || d_grad_out, raw_tmap, sl.stream));
  }
  rc = mark_on(c, sl); if (rc) return rc;                                          // event 1
  gpc::HashArgs ha{};
  ha.cand = c->d_cand; ha.hash = hash_base ? hash_base : c->d_hash;    // hash_base: another [..][H][W] plane set (wide states)
  ha.W = w; ha.H = h; ha.img0 = img0;
  ha.hash_y_end = forest.naive ? h - gpc::kRadius : h - 15;                       // filter.hpp:601-604; the naive filters hash every candidate
  if (c->jit && &forest == &c->forest_dev)
    GPC_CUDA(c, gpc::jit_launch_hash_tiles(c->jit, c->tmap, ha, forest, n_img, sl.stream));
  else
    GPC_CUDA(c, gpc::launch_hash_tiles(c->tmap, ha, forest, n_img, sl.stream));
  c->launches += skip_a1 ? 1 : 2;
  return mark_on(c, sl);                                                           // event 2
}

int ensure_global_ws(gpc_ctx* c, size_t bytes) {
  if (bytes > c->gws_bytes) {
    if (c->d_gws) { GPC_CUDA(c, cudaDeviceSynchronize()); cudaFree(c->d_gws); c->d_gws = nullptr; c->gws_bytes = 0; }   // any lane may still use it
    GPC_CUDA(c, cudaMalloc(&c->d_gws, bytes));
    c->gws_bytes = bytes;
  }
  return GPC_OK;
}

// Radix-sort matcher (match_global.cu) for pairs p0 .. p0 + n of the resident hash images, a chunk of pairs
// per launch sequence, on `stream`.  mode 0: filtered supports, 1: unfiltered correspondences.  Strided output
// (pair_base == nullptr): pair p0 + i writes at d_out + i * out_stride records (capacity cap each).  Packed
// output: the pairs' records back to back from d_out (capacity cap in total), prefix in pair_base[0 .. n].
// Counts to d_n_out[i], candidate counts to d_n_cand[2i..].
size_t sort_workspace_bytes(int w, int h, int n, int* chunk_out) {
  const long long records = 2ll * std::max(w - 2 * gpc::kRadius, 0) * std::max(h - 2 * gpc::kRadius, 0) + 2;
  const size_t per_pair = gpc::global_workspace_bytes(records, 1, h, w);
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>(64, ((size_t)1 << 30) / per_pair));
  if (chunk_out) *chunk_out = chunk;
  return gpc::global_workspace_bytes(records, std::min(chunk, std::max(n, 1)), h, w);
}

int run_match_sort(gpc_ctx* c, const uint32_t* hash, int p0, int n, int w, int h, const gpc_settings* s, int mode, void* d_out,
                   size_t record_bytes, long long out_stride, long long cap, int32_t* d_n_out, int32_t* d_n_cand,
                   cudaStream_t stream, long long* pair_base) {
  const size_t P = (size_t)w * h;
  const long long records = 2ll * std::max(w - 2 * gpc::kRadius, 0) * std::max(h - 2 * gpc::kRadius, 0) + 2;
  int chunk = 1;
  sort_workspace_bytes(w, h, n, &chunk);
  for (int q0 = 0; q0 < n; q0 += chunk) {
    const int m = std::min(chunk, n - q0);
    int rc = ensure_global_ws(c, gpc::global_workspace_bytes(records, m, h, w)); if (rc) return rc;
    int launches = 0;
    GPC_CUDA(c, gpc::launch_match_global(hash + (size_t)(2 * (p0 + q0)) * P, c->d_rows + (size_t)(2 * (p0 + q0)) * h, w, h, m,
                                         s->epipolar_mode ? 1 : 0, 31, s->disp_high, s->vertical_tolerance, mode, c->d_gws, records,
                                         reinterpret_cast<uint8_t*>(d_out) + (size_t)q0 * (size_t)out_stride * record_bytes, out_stride,
                                         cap, d_n_out + q0, d_n_cand ? d_n_cand + 2 * q0 : nullptr, stream, &launches,
                                         s->use_hashtable ? 1 : 0, pair_base ? pair_base + q0 : nullptr, q0 == 0 ? 1 : 0));
    c->launches += launches;
  }
  return GPC_OK;
}

// the per-row matcher covers the sort path's semantics in epipolar mode only; global mode and the reference's
// hashtable matcher (useHashtable, inference.hpp:204-225) go through the device-wide sort
bool use_sort_matcher(const gpc_ctx* c, const gpc_settings* s) {
  return !s->epipolar_mode || s->use_hashtable || c->matcher == GPC_MATCHER_SORT;
}

// Kernels B, scan, C over the slot's hash images.  packed: supports of the slot's pairs back to back
// from d_out[0], prefix in d_pair_base + 2 * p0; else pair i at d_out + i * cap.
// foreign_hash: the hash images came from the caller (gpc_match_hash_images), so nothing is known about which state
// bits are in use.
int run_match(gpc_ctx* c, const Slot& sl, int n_pairs, int w, int h, const gpc_settings* s, gpc_support* d_out,
              long long cap, bool packed, int32_t* d_n_out, int32_t* d_n_cand, bool foreign_hash = false) {
  const size_t P = (size_t)w * h;
  const uint32_t* hash = c->d_hash + (size_t)(2 * sl.p0) * P;
  if (use_sort_matcher(c, s)) {
    // radix sort + segmented scan over the slot's pairs (match_global.cu); same output contract as the row path
    for (int k = 0; k < 2; k++) { int rc = mark_on(c, sl); if (rc) return rc; }   // events 3, 4 (no row kernels here)
    int rc = run_match_sort(c, c->d_hash, sl.p0, n_pairs, w, h, s, 0, d_out, sizeof(gpc_support), packed ? 0 : cap, cap, d_n_out,
                            d_n_cand, sl.stream, packed ? c->d_pair_base + 2 * sl.p0 : nullptr);
    if (rc) return rc;
    return mark_on(c, sl);                                                         // event 5
  }
  const int32_t* rowcnt = c->d_rows + (size_t)(2 * sl.p0) * h;
  int32_t* rowmatch = c->d_rowmatch + (size_t)sl.p0 * h;
  int32_t* rowoff = c->d_rowoff + (size_t)sl.p0 * h;
  uint32_t* stage = c->d_stage + (size_t)sl.p0 * P;
  gpc::MatchArgs m{};
  m.hash = hash; m.lastrow = c->d_lastrow + 2 * sl.p0; m.rowcnt = rowcnt; m.stage = stage; m.rowmatch = rowmatch;
  m.W = w; m.H = h; m.disp_high = s->disp_high; m.vertical_tolerance = s->vertical_tolerance;
  m.wcap = std::max(w - 2 * gpc::kRadius, 16);
  m.table_log2 = table_log2_for(w, m.wcap);
  m.x_bits = ceil_log2(w);
  m.pow2cap = 1 << ceil_log2(m.wcap);
  while ((int)gpc::match_smem_bytes(m.wcap, m.table_log2) > 72 * 1024 &&
         m.table_log2 > std::max(ceil_log2(4 * gpc::match_rows_threads(w)), m.x_bits + 1))
    m.table_log2--;                                  // wide rows: fewer, longer buckets keep >= 3 CTAs per SM
  // significant state bits (only the balance of the ordering buckets depends on it): the kernels place test t < 8 in
  // bit t and test t > 8 in bit t - 1 (test 8 is OR-ed into bit 0, filter.hpp:574-584); the naive mode uses bits 0 .. T-1
  {
    const int T = c->has_forest ? c->forest_host.n_tests : 0;
    const int bits = (c->result_mode == GPC_RESULTS_NAIVE) ? T : (T <= 8 ? T : T - 1);
    m.key_bits = (foreign_hash || !c->has_forest) ? 31 : std::max(1, std::min(31, bits));
  }
  // fast matcher: ~32 four-bit buckets and ~4 slots per candidate of a side.  Constraints: remainder + x of a slot
  // entry fit one word (slot_log2 >= x_bits), buckets refine slots (slot_log2 <= nib_log2 + 3), the byte offset of a
  // word fits 16 bits (nib_log2 <= 14), and the ordering pass (8 bytes per match) reuses both tables.
  static const int nib_env = std::getenv("GPC_B_NIB_LOG2") ? std::atoi(std::getenv("GPC_B_NIB_LOG2")) : 0;
  static const int slot_env = std::getenv("GPC_B_SLOT_LOG2") ? std::atoi(std::getenv("GPC_B_SLOT_LOG2")) : 0;
  const int tbl_min = std::max(ceil_log2(4 * gpc::match_rows_threads(w)), m.x_bits);     // the zero fill writes 16 bytes per thread and round
  m.nib_log2 = std::min(14, std::max(ceil_log2(4 * m.wcap), tbl_min));
  m.slot_log2 = std::max(m.nib_log2 - 1, tbl_min);       // measured: 2 slots per candidate beat 4 (one more CTA per SM)
  while ((int)gpc::match_fast_smem_bytes(m.nib_log2, m.slot_log2, m.pow2cap) > 80 * 1024 && (1 << (m.nib_log2 - 1)) >= m.pow2cap &&
         m.nib_log2 - 1 >= tbl_min) {
    m.nib_log2--; m.slot_log2 = std::max(m.slot_log2 - 1, tbl_min);   // wide rows: keep several CTAs per SM
  }
  if (nib_env > 0) m.nib_log2 = std::min(14, std::max(nib_env, tbl_min));
  if (slot_env > 0) m.slot_log2 = std::min(m.nib_log2 + 3, std::max(slot_env, tbl_min));
  while ((1 << m.nib_log2) + (1 << m.slot_log2) < 2 * m.pow2cap) { if (m.nib_log2 < 14) m.nib_log2++; else m.slot_log2++; }
  // row lists: 2 header words per pair slot at the front of each array, then 2 * max_h entry words per pair slot
  m.fb_hdr = c->d_fb + 2 * (size_t)sl.p0;
  m.fb_ent = c->d_fb + 2 * (size_t)c->max_batch + (size_t)sl.p0 * 2 * c->max_h;
  m.big_hdr = c->d_big + 2 * (size_t)sl.p0;
  m.big_ent = c->d_big + 2 * (size_t)c->max_batch + (size_t)sl.p0 * 2 * c->max_h;
  m.mrec = c->d_mrec + (size_t)sl.p0 * P;
  m.ovbuf = c->d_ovbuf + (size_t)sl.p0 * h * 4 * gpc::match_ov_cap();
  m.rowhdr = c->d_rowhdr + (size_t)sl.p0 * h * 4;
  const bool general = (c->matcher == GPC_MATCHER_ROWS_GENERAL);
  if ((int)gpc::match_smem_bytes(m.wcap, m.table_log2) > c->match_smem_max ||
      (int)gpc::match_fast_smem_bytes(m.nib_log2, m.slot_log2, m.pow2cap) > c->match_smem_max ||
      (int)gpc::order_rows_smem_bytes(m.pow2cap) > c->match_smem_max)
    return fail(c, GPC_E_DIMS, "image too wide for the row matcher's shared memory");
  if (h - 2 * gpc::kRadius <= 0) GPC_CUDA(c, cudaMemsetAsync(rowmatch, 0, (size_t)n_pairs * h * sizeof(int32_t), sl.stream));
  GPC_CUDA(c, gpc::launch_match_rows(m, n_pairs, general ? 1 : 0, c->sm_count, sl.stream));
  int rc = mark_on(c, sl); if (rc) return rc;                                      // event 3
  GPC_CUDA(c, gpc::launch_row_scan(rowmatch, rowcnt, h, n_pairs, rowoff, d_n_out, d_n_cand, sl.stream));
  c->launches += general ? 2 : 5;
  const long long* pair_base = nullptr;
  if (packed) {
    GPC_CUDA(c, gpc::launch_pair_scan(d_n_out, n_pairs, c->d_pair_base + 2 * sl.p0, sl.stream));
    c->launches += 1;
    pair_base = c->d_pair_base + 2 * sl.p0;
  }
  rc = mark_on(c, sl); if (rc) return rc;                                          // event 4
  GPC_CUDA(c, gpc::launch_emit_supports(stage, rowmatch, rowoff, pair_base, d_out, cap, w, h, n_pairs, sl.stream));
  if (h - 2 * gpc::kRadius > 0) c->launches += 1;
  return mark_on(c, sl);                                                           // event 5
}

// Pinned staging of at least `bytes`; waits for the previous asynchronous use of the buffer.
int ensure_host_stage(gpc_ctx* c, size_t bytes) {
  if (!c->ev_stage) GPC_CUDA(c, cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming));
  else GPC_CUDA(c, cudaEventSynchronize(c->ev_stage));
  if (bytes > c->h_stage_bytes) {
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr; c->h_stage_bytes = 0;
    const size_t want = std::max(bytes, (size_t)c->max_w * c->max_h * 4);
    GPC_CUDA(c, cudaMallocHost(&c->h_stage, want));
    c->h_stage_bytes = want;
  }
  return GPC_OK;
}

int ensure_debug_buffers(gpc_ctx* c) {
  size_t P = (size_t)c->max_w * c->max_h;
  if (!c->d_dbg8) GPC_CUDA(c, cudaMalloc(&c->d_dbg8, 2 * P));
  if (!c->d_mask) GPC_CUDA(c, cudaMalloc(&c->d_mask, P * sizeof(int32_t)));
  return GPC_OK;
}

}  // namespace

extern "C" {

const char* gpc_status_string(int status) {
  switch (status) {
    case GPC_OK: return "ok";
    case GPC_E_ARG: return "invalid argument";
    case GPC_E_WIDTH16: return "width must be multiple of 16";
    case GPC_E_DIMS: return "dimensions exceed context capacity";
    case GPC_E_CUDA: return "CUDA error";
    case GPC_E_CAPACITY: return "output capacity too small";
    case GPC_E_UNSUPPORTED: return "unsupported mode";
    case GPC_E_FOREST: return "invalid or missing forest";
    case GPC_E_IO: return "cannot open file";
    default: return "unknown status";
  }
}

const char* gpc_last_error(const gpc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gpc_create(gpc_ctx** out, int device, int max_w, int max_h, int max_batch) {
  if (!out) return fail(nullptr, GPC_E_ARG, "out is NULL");
  *out = nullptr;
  if (max_w <= 0 || max_h <= 0 || max_batch <= 0) return fail(nullptr, GPC_E_ARG, "non-positive capacity");
  if (max_w % 16 != 0) return fail(nullptr, GPC_E_WIDTH16, "max_w must be a multiple of 16");
  // the kernels index image rows of a batch in 32 bits (2 * max_batch * max_h rows); far beyond any buffer that fits HBM
  if ((long long)max_batch * max_h > (1ll << 29)) return fail(nullptr, GPC_E_DIMS, "max_batch * max_h exceeds 2^29 rows");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, GPC_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                         " (libgpc_b200 has no CPU fallback)");
  if (device < 0 || device >= n_dev) return fail(nullptr, GPC_E_ARG, "device index out of range");
  static std::atomic<long long> next_id{1};                // contexts may be created from several host threads
  gpc_ctx* c = new gpc_ctx();
  c->id = next_id++;
  c->device = device; c->max_w = max_w; c->max_h = max_h; c->max_batch = max_batch;
  auto bail = [&](const char* what, cudaError_t err) {
    std::string msg = std::string(what) + ": " + cudaGetErrorString(err);
    gpc_destroy(c);
    return fail(nullptr, GPC_E_CUDA, msg);
  };
#define TRY(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) return bail(#call, e2); } while (0)
  TRY(cudaSetDevice(device));
  TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  TRY(gpc::configure_hash_tiles());
  int smem_optin = 0;
  TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  c->match_smem_max = smem_optin - 1024;
  TRY(gpc::configure_match_rows(c->match_smem_max));
  TRY(gpc::configure_match_global());
  const size_t P = (size_t)max_w * max_h, B = (size_t)max_batch;
  TRY(cudaMalloc(&c->d_raw, 2 * B * P));
  TRY(cudaMalloc(&c->d_smooth, 2 * B * P));
  TRY(cudaMalloc(&c->d_cand, 2 * B * (size_t)max_h * (max_w / 16) * sizeof(uint16_t)));
  TRY(cudaMalloc(&c->d_hash, 2 * B * P * sizeof(uint32_t)));
  TRY(cudaMalloc(&c->d_stage, B * P * sizeof(uint32_t)));
  TRY(cudaMemsetAsync(c->d_stage, 0, B * P * sizeof(uint32_t), c->stream));   // kernel C reads a row's words ahead of its length
  TRY(cudaMalloc(&c->d_rows, 2 * B * max_h * sizeof(int32_t)));
  TRY(cudaMalloc(&c->d_lastrow, 2 * B * sizeof(int32_t)));
  for (int l = 0; l < gpc_ctx::kLanes; l++) TRY(cudaStreamCreateWithFlags(&c->lane_stream[l], cudaStreamNonBlocking));
  TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  if (const char* e = std::getenv("GPC_CHUNK_PAIRS")) c->chunk_pairs = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("GPC_A1_TMA")) c->a1_tma = std::atoi(e) != 0;
  TRY(cudaMalloc(&c->d_rowmatch, B * max_h * sizeof(int32_t)));
  TRY(cudaMalloc(&c->d_fb, B * (2 * (size_t)max_h + 2) * sizeof(uint32_t)));
  TRY(cudaMemsetAsync(c->d_fb, 0, B * (2 * (size_t)max_h + 2) * sizeof(uint32_t), c->stream));
  TRY(cudaMalloc(&c->d_big, B * (2 * (size_t)max_h + 2) * sizeof(uint32_t)));
  TRY(cudaMemsetAsync(c->d_big, 0, B * (2 * (size_t)max_h + 2) * sizeof(uint32_t), c->stream));
  TRY(cudaMalloc(&c->d_mrec, B * P * sizeof(unsigned long long)));
  TRY(cudaMalloc(&c->d_ovbuf, B * (size_t)max_h * 4 * gpc::match_ov_cap() * sizeof(uint32_t)));
  TRY(cudaMalloc(&c->d_rowhdr, B * (size_t)max_h * 4 * sizeof(int32_t)));
  TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  TRY(cudaMalloc(&c->d_rowoff, B * (max_h + 1) * sizeof(int32_t)));
  TRY(cudaMalloc(&c->d_totals, B * sizeof(int32_t)));
  TRY(cudaMalloc(&c->d_ncand, 2 * B * sizeof(int32_t)));
  TRY(cudaMalloc(&c->d_pair_base, (2 * B + 2) * sizeof(long long)));
  long long per_pair = (long long)std::max(max_w - 26, 0) * std::max(max_h - 26, 0);
  c->out_cap = std::max<long long>(per_pair * (long long)B, 1);
  TRY(cudaMalloc(&c->d_out, (size_t)c->out_cap * sizeof(gpc_support)));
  TRY(cudaMallocHost(&c->h_counts, 3 * B * sizeof(int32_t)));
  TRY(cudaMallocHost(&c->h_pair_base, (2 * B + 2) * sizeof(long long)));
  TRY(cudaMemsetAsync(c->d_rowmatch, 0, B * max_h * sizeof(int32_t), c->stream));
  TRY(cudaStreamSynchronize(c->stream));
#undef TRY
  { std::lock_guard<std::mutex> lk(g_registry_mu); g_registry[c->id] = c; }
  *out = c;
  return GPC_OK;
}

void gpc_destroy(gpc_ctx* c) {
  if (!c) return;
  { std::lock_guard<std::mutex> lk(g_registry_mu); g_registry.erase(c->id); }
  cudaSetDevice(c->device);
  for (auto& b : c->image_pool) cudaFree(b.second);
  cudaFree(c->d_raw); cudaFree(c->d_smooth); cudaFree(c->d_cand); cudaFree(c->d_hash); cudaFree(c->d_stage); cudaFree(c->d_rows); cudaFree(c->d_lastrow); cudaFree(c->d_rowmatch); cudaFree(c->d_fb); cudaFree(c->d_big); cudaFree(c->d_mrec); cudaFree(c->d_ovbuf); cudaFree(c->d_rowhdr);
  for (int l = 0; l < gpc_ctx::kLanes; l++) if (c->lane_stream[l]) cudaStreamDestroy(c->lane_stream[l]);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  for (cudaEvent_t e : c->ev_chunk) cudaEventDestroy(e);
  cudaFree(c->d_rowoff); cudaFree(c->d_totals); cudaFree(c->d_ncand); cudaFree(c->d_pair_base); cudaFree(c->d_out);
  gpc::jit_destroy(c->jit);
  cudaFree(c->d_dbg8); cudaFree(c->d_mask); cudaFree(c->d_gws); cudaFree(c->d_pyr); cudaFree(c->d_wide);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->ev_stage) cudaEventDestroy(c->ev_stage);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->h_pair_base) cudaFreeHost(c->h_pair_base);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int gpc_set_stream(gpc_ctx* c, void* cuda_stream) {
  if (!c) return GPC_E_ARG;
  c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return GPC_OK;
}

int gpc_synchronize(gpc_ctx* c) {
  if (!c) return GPC_E_ARG;
  GPC_CUDA(c, cudaSetDevice(c->device));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  return GPC_OK;
}

int64_t gpc_launch_count(const gpc_ctx* c) { return c ? c->launches : 0; }

int64_t gpc_context_id(const gpc_ctx* c) { return c ? (int64_t)c->id : 0; }

const char* gpc_jit_status(const gpc_ctx* c) { return c ? c->jit_note.c_str() : ""; }

int gpc_enable_kernel_timing(gpc_ctx* c, int on) {
  if (!c) return GPC_E_ARG;
  c->timing = (on != 0);
  c->ev_used = 0;
  for (int k = 0; k < GPC_N_KERNELS; k++) c->k_ms[k] = 0.0;
  c->k_runs = 0;
  return GPC_OK;
}

// Synchronises the stream, folds the recorded event groups into per-kernel sums and returns
// the accumulated milliseconds per kernel and the number of batch runs they cover.
int gpc_kernel_times(gpc_ctx* c, double* ms, int64_t* runs) {
  if (!c || !ms || !runs) return GPC_E_ARG;
  GPC_CUDA(c, cudaSetDevice(c->device));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i + (GPC_N_KERNELS + 1) <= c->ev_used; i += GPC_N_KERNELS + 1) {
    for (int k = 0; k < GPC_N_KERNELS; k++) {
      float t = 0.f;
      GPC_CUDA(c, cudaEventElapsedTime(&t, c->ev_pool[i + k], c->ev_pool[i + k + 1]));
      c->k_ms[k] += t;
    }
    c->k_runs++;
  }
  c->ev_used = 0;
  for (int k = 0; k < GPC_N_KERNELS; k++) ms[k] = c->k_ms[k];
  *runs = c->k_runs;
  return GPC_OK;
}

// Forest::readForest (inference.hpp:404-446): whitespace-separated text; scale tag ignored;
// at most 32 tests kept; type 1 iff ANY test of the file (kept or not) has tau != 0.
int gpc_read_forest(const char* path, gpc_forest* out) {
  if (!path || !out) return GPC_E_ARG;
  std::memset(out, 0, sizeof(*out));
  std::ifstream ff(path);
  if (ff.fail()) return GPC_E_IO;
  int num_ferns = 0, nonzero = 0;
  ff >> num_ferns;
  out->n_ferns = num_ferns;
  for (int i = 0; i < num_ferns && ff.good(); i++) {
    int id = 0, nt = 0;
    std::string scale;
    ff >> id >> scale >> nt;
    for (int j = 0; j < nt; j++) {
      int lvl = 0, ix = 0, iy = 0, jx = 0, jy = 0, tau = 0;
      ff >> lvl >> ix >> iy >> jx >> jy >> tau;
      if (ff.fail()) break;
      if (out->n_tests < GPC_MAX_TESTS) {
        int t = out->n_tests++;
        out->ix[t] = ix; out->iy[t] = iy; out->jx[t] = jx; out->jy[t] = jy; out->tau[t] = tau;
      } else {
        out->n_discarded++;
      }
      if (tau != 0) nonzero++;
    }
  }
  out->type = nonzero ? 1 : 0;
  return GPC_OK;
}

int gpc_set_forest(gpc_ctx* c, const gpc_forest* f) {
  if (!c || !f) return GPC_E_ARG;
  if (f->n_tests < 0 || f->n_tests > GPC_MAX_TESTS) return fail(c, GPC_E_FOREST, "a forest has at most 32 tests");
  for (int t = 0; t < f->n_tests; t++) {
    const int v[4] = {f->ix[t], f->iy[t], f->jx[t], f->jy[t]};
    for (int k = 0; k < 4; k++)
      if (v[k] < -GPC_PATCH_RADIUS || v[k] > GPC_PATCH_RADIUS)
        return fail(c, GPC_E_FOREST, "test offset outside the 27x27 patch (|offset| <= 13)");
  }
  if (c->has_forest && std::memcmp(&c->forest_host, f, sizeof(gpc_forest)) == 0) return GPC_OK;   // unchanged: keep the kernel
  GPC_CUDA(c, cudaSetDevice(c->device));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));     // earlier launches may still use the previous specialised kernel
  for (int l = 0; l < gpc_ctx::kLanes; l++) GPC_CUDA(c, cudaStreamSynchronize(c->lane_stream[l]));
  if (c->result_mode == GPC_RESULTS_NAIVE && f->n_tests > 31)
    return fail(c, GPC_E_UNSUPPORTED, "GPC_RESULTS_NAIVE supports forests of at most 31 tests (bit 31 of a hash word is the candidate flag)");
  gpc::jit_destroy(c->jit);
  c->jit = nullptr;
  c->forest_host = *f;
  bake_forest(*f, &c->forest_dev, c->result_mode);
  c->has_forest = true;
  c->forest_serial++;
  if (c->result_mode == GPC_RESULTS_NAIVE) { c->jit_note = "generic: naive result mode"; return GPC_OK; }
  std::string why;
  c->jit = gpc::jit_build_hash_tiles(c->forest_dev, &why);
  c->jit_note = c->jit ? "specialised" : ("generic: " + why);
  return GPC_OK;
}

// Which build of the reference the results reproduce: GPC_RESULTS_SSE (default; the reference built with
// -D_INTRINSICS_SSE, samples/CMakeLists.txt:13-17) or GPC_RESULTS_NAIVE (its SSE=OFF build: boxNaive, sobelNaive,
// gpcFilterNaive / gpcFilterTauNaive, filter.hpp:157-293).  A forest set earlier is re-baked for the new mode.
int gpc_set_result_mode(gpc_ctx* c, int mode) {
  if (!c) return GPC_E_ARG;
  if (mode != GPC_RESULTS_SSE && mode != GPC_RESULTS_NAIVE) return fail(c, GPC_E_ARG, "unknown result mode");
  if (mode == c->result_mode) return GPC_OK;
  if (mode == GPC_RESULTS_NAIVE && c->has_forest && c->forest_host.n_tests > 31)
    return fail(c, GPC_E_UNSUPPORTED, "GPC_RESULTS_NAIVE supports forests of at most 31 tests (bit 31 of a hash word is the candidate flag)");
  const int old_mode = c->result_mode;
  c->result_mode = mode;
  c->wide_words.clear(); c->wide_tests = 0;       // the word layout depends on the mode
  if (!c->has_forest) return GPC_OK;
  const gpc_forest f = c->forest_host;
  c->has_forest = false;                              // defeat the "unchanged forest" shortcut
  const int rc = gpc_set_forest(c, &f);
  if (rc != GPC_OK && !c->has_forest) {               // nothing was re-baked: the context keeps its mode and forest
    c->result_mode = old_mode;
    c->has_forest = true;
  }
  return rc;
}

int gpc_match_batch_device(gpc_ctx* c, const uint8_t* d_images, int n_pairs, int w, int h, const gpc_settings* s,
                           gpc_support* d_out, int cap_per_pair, int32_t* d_n_out, int32_t* d_n_cand) {
  if (!c || !d_images || !d_out || !d_n_out || cap_per_pair < 0) return fail(c, GPC_E_ARG, "null argument");
  int rc = check_dims(c, w, h, n_pairs); if (rc) return rc;
  rc = check_settings(c, s); if (rc) return rc;
  if (!c->has_forest) return fail(c, GPC_E_FOREST, "no forest set");
  GPC_CUDA(c, cudaSetDevice(c->device));
  // Large batches run as two halves on two streams: the second half's stencil / hashing kernels (ALU bound) overlap
  // the first half's matcher tail, scans and emission (latency bound, few resident warps).  GPC_DEVICE_SPLIT=1: off.
  static const int split_env = std::getenv("GPC_DEVICE_SPLIT") ? std::atoi(std::getenv("GPC_DEVICE_SPLIT")) : 2;
  const size_t P = (size_t)w * h;
  if (split_env >= 2 && n_pairs >= 32 * split_env && !c->timing && !use_sort_matcher(c, s)) {
    cudaStream_t st[2] = {c->stream, c->lane_stream[0]};
    GPC_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
    GPC_CUDA(c, cudaStreamWaitEvent(st[1], c->ev_fork, 0));
    for (int k = 0; k < split_env; k++) {                          // part k on stream k % 2, each on its own slice
      const int p0 = (int)((long long)n_pairs * k / split_env), p1 = (int)((long long)n_pairs * (k + 1) / split_env);
      const Slot sl{p0, st[k & 1]};
      rc = run_preprocess(c, sl, d_images + (size_t)(2 * p0) * P, 2 * (p1 - p0), w, h, s->gradient_threshold, c->forest_dev, nullptr, nullptr);
      if (rc) return rc;
      rc = run_match(c, sl, p1 - p0, w, h, s, d_out + (size_t)p0 * cap_per_pair, cap_per_pair, false, d_n_out + p0,
                     d_n_cand ? d_n_cand + 2 * p0 : nullptr);
      if (rc) return rc;
    }
    if (!c->ev_join) GPC_CUDA(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    GPC_CUDA(c, cudaEventRecord(c->ev_join, st[1]));
    GPC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    return GPC_OK;
  }
  rc = run_preprocess(c, Slot{0, c->stream}, d_images, 2 * n_pairs, w, h, s->gradient_threshold, c->forest_dev, nullptr, nullptr);
  if (rc) return rc;
  return run_match(c, Slot{0, c->stream}, n_pairs, w, h, s, d_out, cap_per_pair, false, d_n_out, d_n_cand);
}

// ---- chunked, pipelined batches ---------------------------------------------------------------------------------
// A batch is cut into chunks of pairs.  Chunk k's supports start where the supports of chunks 0 .. k-1 end, so a
// download can only be aimed once every earlier chunk has reported its count.  ChunkBoard is that bookkeeping; with one
// context it is trivial, with a pool (one context and host thread per GPU, chunks dealt round-robin) it is the only
// thing the workers share: a host-side decoupled look-back.  Nothing is copied twice -- every chunk's supports go from
// its device straight to their final place in the caller's buffer.
struct ChunkBoard {
  std::vector<int> first;                 // [n_chunks + 1] first pair of each chunk
  std::vector<long long> total;           // [n_chunks] supports of the chunk, valid once known[k]
  std::vector<char> known;
  std::mutex mu;
  std::condition_variable cv;
  bool failed = false;                    // a worker gave up: nobody waits any longer

  // chunk sizes ramp up at the start and down at the end (CH/4, CH/2, CH, ..., CH, CH/2, CH/4): the first upload and
  // the last kernels + download are the only parts of the pipeline nothing overlaps
  void plan(int n_pairs, int CH) {
    std::vector<int> sizes, ramp;
    for (int v = std::max(1, CH / 4); v < CH; v *= 2) ramp.push_back(v);
    int ramp_sum = 0;
    for (int v : ramp) ramp_sum += v;
    int rest = n_pairs;
    const bool ramped = n_pairs >= 2 * ramp_sum + CH;
    if (ramped) { sizes = ramp; rest -= 2 * ramp_sum; }
    for (; rest > 0; rest -= CH) sizes.push_back(std::min(CH, rest));
    if (ramped) sizes.insert(sizes.end(), ramp.rbegin(), ramp.rend());
    first.assign(sizes.size() + 1, 0);
    for (size_t k = 0; k < sizes.size(); k++) first[k + 1] = first[k] + sizes[k];
    total.assign(sizes.size(), 0);
    known.assign(sizes.size(), 0);
  }
  int n_chunks() const { return (int)total.size(); }
  void publish(int k, long long t) {
    { std::lock_guard<std::mutex> lk(mu); total[(size_t)k] = t; known[(size_t)k] = 1; }
    cv.notify_all();
  }
  void fail_all() {
    { std::lock_guard<std::mutex> lk(mu); failed = true; }
    cv.notify_all();
  }
  // supports of chunks 0 .. k-1; false if another worker failed
  bool base_of(int k, long long* base) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] {
      if (failed) return true;
      for (int j = 0; j < k; j++) if (!known[(size_t)j]) return false;
      return true;
    });
    if (failed) return false;
    long long b = 0;
    for (int j = 0; j < k; j++) b += total[(size_t)j];
    *base = b;
    return true;
  }
};

// The chunks `mine` (ascending) of the board on context c: one stream uploads them back to back, a second runs the
// kernels chunk after chunk, a third downloads each chunk's supports as soon as its place is known -- upload, kernels
// and download overlap (PCIe is full duplex).  The chunks occupy consecutive slices of the context's resident buffers.
// *need (optional) receives the supports of the whole batch seen by this worker's last chunk base + its own total.
static int run_chunks(gpc_ctx* c, ChunkBoard& board, const std::vector<int>& mine, const uint8_t* images, int w, int h,
                      const gpc_settings* s, gpc_support* out, int64_t cap, int64_t* offsets, int32_t* n_cand, bool* overflow_out) {
  const size_t P = (size_t)w * h;
  const long long per_pair = c->out_cap / c->max_batch;
  // whatever the exit path, no copy from `images` or into `out` may still be queued when the caller gets control back
  struct LaneGuard {
    gpc_ctx* c;
    ~LaneGuard() { for (int l = 0; l < gpc_ctx::kLanes; l++) cudaStreamSynchronize(c->lane_stream[l]); }
  } lane_guard{c};
  const int nm = (int)mine.size();
  std::vector<int> local0((size_t)nm + 1, 0);                       // first local pair slot of each of my chunks
  for (int i = 0; i < nm; i++) local0[(size_t)i + 1] = local0[(size_t)i] + (board.first[(size_t)mine[(size_t)i] + 1] - board.first[(size_t)mine[(size_t)i]]);
  if (local0[(size_t)nm] > c->max_batch) return fail(c, GPC_E_DIMS, "image or batch exceeds the context's capacity");
  while ((int)c->ev_chunk.size() < 2 * nm) {
    cudaEvent_t e;
    GPC_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_chunk.push_back(e);
  }
  // lane 0: every upload, back to back; lane 1: the kernels, chunk after chunk; lane 2: the downloads
  cudaStream_t s_up = c->lane_stream[0], s_run = c->lane_stream[1], s_down = c->lane_stream[2];
  GPC_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));               // order after earlier work of the context
  for (int l = 0; l < gpc_ctx::kLanes; l++) GPC_CUDA(c, cudaStreamWaitEvent(c->lane_stream[l], c->ev_fork, 0));
  for (int i = 0; i < nm; i++) {
    const int g0 = board.first[(size_t)mine[(size_t)i]], n = local0[(size_t)i + 1] - local0[(size_t)i], p0 = local0[(size_t)i];
    GPC_CUDA(c, cudaMemcpyAsync(c->d_raw + (size_t)(2 * p0) * P, images + (size_t)(2 * g0) * P, 2 * (size_t)n * P,
                                cudaMemcpyHostToDevice, s_up));
    GPC_CUDA(c, cudaEventRecord(c->ev_chunk[2 * (size_t)i], s_up));
  }
  for (int i = 0; i < nm; i++) {
    const int n = local0[(size_t)i + 1] - local0[(size_t)i], p0 = local0[(size_t)i];
    const Slot sl{p0, s_run};
    GPC_CUDA(c, cudaStreamWaitEvent(s_run, c->ev_chunk[2 * (size_t)i], 0));
    int rc = run_preprocess(c, sl, c->d_raw + (size_t)(2 * p0) * P, 2 * n, w, h, s->gradient_threshold, c->forest_dev, nullptr, nullptr);
    if (rc) return rc;
    rc = run_match(c, sl, n, w, h, s, c->d_out + (size_t)p0 * per_pair, (long long)n * per_pair, true, c->d_totals + p0,
                   c->d_ncand + 2 * p0);
    if (rc) return rc;
    GPC_CUDA(c, cudaMemcpyAsync(c->h_pair_base + 2 * p0, c->d_pair_base + 2 * p0, (size_t)(n + 1) * sizeof(long long),
                                cudaMemcpyDeviceToHost, s_run));
    if (n_cand)
      GPC_CUDA(c, cudaMemcpyAsync(c->h_counts + 2 * p0, c->d_ncand + 2 * p0, 2 * (size_t)n * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, s_run));
    GPC_CUDA(c, cudaEventRecord(c->ev_chunk[2 * (size_t)i + 1], s_run));
  }
  bool overflow = false;
  for (int i = 0; i < nm; i++) {                                       // downloads follow the kernels chunk by chunk
    const int k = mine[(size_t)i], g0 = board.first[(size_t)k], n = local0[(size_t)i + 1] - local0[(size_t)i], p0 = local0[(size_t)i];
    GPC_CUDA(c, cudaEventSynchronize(c->ev_chunk[2 * (size_t)i + 1]));
    const long long* pb = c->h_pair_base + 2 * p0;
    const long long m = pb[n];
    board.publish(k, m);
    long long base = 0;
    if (!board.base_of(k, &base)) return fail(c, GPC_E_CUDA, "another device of the pool failed");
    for (int j = 0; j < n; j++) offsets[g0 + j + 1] = base + pb[j + 1];
    if (n_cand) std::memcpy(n_cand + 2 * g0, c->h_counts + 2 * p0, 2 * (size_t)n * sizeof(int32_t));
    if (base + m > cap) overflow = true;
    else if (m > 0)
      GPC_CUDA(c, cudaMemcpyAsync(out + base, c->d_out + (size_t)p0 * per_pair, (size_t)m * sizeof(gpc_support),
                                  cudaMemcpyDeviceToHost, s_down));
  }
  for (int l = 0; l < gpc_ctx::kLanes; l++) GPC_CUDA(c, cudaStreamSynchronize(c->lane_stream[l]));
  if (overflow_out) *overflow_out = overflow;
  return GPC_OK;
}

// Pairs per pipeline chunk.  Every chunk costs the host ~16 API calls and the GPU a chain of small kernels, every
// batch pays one chunk of pipeline fill and drain: a quarter of the batch, between 16 and 64 pairs (measured at 256
// Sintel pairs: 8 / 16 / 32 / 64 pairs per chunk = 0.74 / 0.82 / 0.87 / 0.90 of the box's copy ceiling).
static int chunk_for(const gpc_ctx* c, int n_pairs) {
  if (c->chunk_pairs > 0) return c->chunk_pairs;
  return std::max(16, std::min(64, n_pairs / 4));
}

// Pipelined body of gpc_match_batch: every chunk on this one context.
static int match_batch_pipelined(gpc_ctx* c, const uint8_t* images, int n_pairs, int w, int h, const gpc_settings* s,
                                 gpc_support* out, int64_t cap, int64_t* offsets, int32_t* n_cand) {
  ChunkBoard board;
  board.plan(n_pairs, chunk_for(c, n_pairs));
  std::vector<int> mine((size_t)board.n_chunks());
  for (int k = 0; k < board.n_chunks(); k++) mine[(size_t)k] = k;
  offsets[0] = 0;
  bool overflow = false;
  const int rc = run_chunks(c, board, mine, images, w, h, s, out, cap, offsets, n_cand, &overflow);
  if (rc) return rc;
  if (overflow) return fail(c, GPC_E_CAPACITY, "support buffer too small: need " + std::to_string(offsets[n_pairs]));
  return GPC_OK;
}

int gpc_match_batch(gpc_ctx* c, const uint8_t* images, int n_pairs, int w, int h, const gpc_settings* s,
                    gpc_support* out, int64_t cap, int64_t* offsets, int32_t* n_cand) {
  if (!c || !images || !offsets || cap < 0 || (cap > 0 && !out)) return fail(c, GPC_E_ARG, "null argument");
  int rc = check_dims(c, w, h, n_pairs); if (rc) return rc;
  rc = check_settings(c, s); if (rc) return rc;
  if (!c->has_forest) return fail(c, GPC_E_FOREST, "no forest set");
  GPC_CUDA(c, cudaSetDevice(c->device));
  const size_t P = (size_t)w * h;
  if (!c->timing && n_pairs >= 2 * chunk_for(c, n_pairs) && h > 2 * gpc::kRadius) {
    if (use_sort_matcher(c, s)) {                                  // size the sort workspace before the lanes start
      rc = ensure_global_ws(c, sort_workspace_bytes(w, h, chunk_for(c, n_pairs), nullptr)); if (rc) return rc;
    }
    return match_batch_pipelined(c, images, n_pairs, w, h, s, out, cap, offsets, n_cand);
  }
  GPC_CUDA(c, cudaMemcpyAsync(c->d_raw, images, 2 * (size_t)n_pairs * P, cudaMemcpyHostToDevice, c->stream));
  rc = run_preprocess(c, Slot{0, c->stream}, c->d_raw, 2 * n_pairs, w, h, s->gradient_threshold, c->forest_dev, nullptr, nullptr);
  if (rc) return rc;
  rc = run_match(c, Slot{0, c->stream}, n_pairs, w, h, s, c->d_out, c->out_cap, true, c->d_totals, c->d_ncand);
  if (rc) return rc;
  GPC_CUDA(c, cudaMemcpyAsync(c->h_pair_base, c->d_pair_base, (size_t)(n_pairs + 1) * sizeof(long long),
                              cudaMemcpyDeviceToHost, c->stream));
  if (n_cand)
    GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_ncand, 2 * (size_t)n_pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int p = 0; p <= n_pairs; p++) offsets[p] = c->h_pair_base[p];
  if (n_cand) std::memcpy(n_cand, c->h_counts, 2 * (size_t)n_pairs * sizeof(int32_t));
  const long long total = c->h_pair_base[n_pairs];
  if (total > cap) return fail(c, GPC_E_CAPACITY, "support buffer too small: need " + std::to_string(total));
  if (total > 0) {
    GPC_CUDA(c, cudaMemcpyAsync(out, c->d_out, (size_t)total * sizeof(gpc_support), cudaMemcpyDeviceToHost, c->stream));
    GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return GPC_OK;
}

int gpc_match_pair(gpc_ctx* c, const uint8_t* left, const uint8_t* right, int w, int h, int stride, const gpc_settings* s,
                   gpc_support* out, int cap, int* n_out, int* n_cand_l, int* n_cand_r) {
  if (!c || !left || !right || !n_out || cap < 0 || (cap > 0 && !out)) return fail(c, GPC_E_ARG, "null argument");
  if (stride < w) return fail(c, GPC_E_ARG, "stride smaller than width");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  rc = check_settings(c, s); if (rc) return rc;
  if (!c->has_forest) return fail(c, GPC_E_FOREST, "no forest set");
  GPC_CUDA(c, cudaSetDevice(c->device));
  const size_t P = (size_t)w * h;
  GPC_CUDA(c, cudaMemcpy2DAsync(c->d_raw, w, left, stride, w, h, cudaMemcpyHostToDevice, c->stream));
  GPC_CUDA(c, cudaMemcpy2DAsync(c->d_raw + P, w, right, stride, w, h, cudaMemcpyHostToDevice, c->stream));
  rc = run_preprocess(c, Slot{0, c->stream}, c->d_raw, 2, w, h, s->gradient_threshold, c->forest_dev, nullptr, nullptr);
  if (rc) return rc;
  rc = run_match(c, Slot{0, c->stream}, 1, w, h, s, c->d_out, c->out_cap, true, c->d_totals, c->d_ncand);
  if (rc) return rc;
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_totals, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts + 1, c->d_ncand, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_out = c->h_counts[0];
  if (n_cand_l) *n_cand_l = c->h_counts[1];
  if (n_cand_r) *n_cand_r = c->h_counts[2];
  if (*n_out > cap) return fail(c, GPC_E_CAPACITY, "support buffer too small: need " + std::to_string(*n_out));
  if (*n_out > 0) {
    GPC_CUDA(c, cudaMemcpyAsync(out, c->d_out, (size_t)*n_out * sizeof(gpc_support), cudaMemcpyDeviceToHost, c->stream));
    GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return GPC_OK;
}

static int preprocess_device(gpc_ctx* c, const uint8_t* d_img, int w, int h, int thr, uint8_t* smooth, uint8_t* grad,
                             int32_t* mask, int mask_cap, int* n_mask, gpc_image* im);

int gpc_preprocess(gpc_ctx* c, const uint8_t* img, int w, int h, int thr, uint8_t* smooth, uint8_t* grad,
                   int32_t* mask, int mask_cap, int* n_mask) {
  if (!c || !img) return fail(c, GPC_E_ARG, "null argument");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  if (thr < 0 || thr > 255) return fail(c, GPC_E_ARG, "gradientThreshold needs to be within 0...255");
  GPC_CUDA(c, cudaSetDevice(c->device));
  GPC_CUDA(c, cudaMemcpyAsync(c->d_raw, img, (size_t)w * h, cudaMemcpyHostToDevice, c->stream));
  return preprocess_device(c, c->d_raw, w, h, thr, smooth, grad, mask, mask_cap, n_mask, nullptr);
}

int gpc_hash(gpc_ctx* c, const uint8_t* img, int w, int h, int thr, uint32_t* states, int32_t* mask, int cap, int* n,
             uint32_t* hash_image) {
  if (!c || !img || !n) return fail(c, GPC_E_ARG, "null argument");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  if (!c->has_forest) return fail(c, GPC_E_FOREST, "no forest set");
  if (thr < 0 || thr > 255) return fail(c, GPC_E_ARG, "gradientThreshold needs to be within 0...255");
  GPC_CUDA(c, cudaSetDevice(c->device));
  rc = ensure_debug_buffers(c); if (rc) return rc;
  const size_t P = (size_t)w * h;
  GPC_CUDA(c, cudaMemcpyAsync(c->d_raw, img, P, cudaMemcpyHostToDevice, c->stream));
  rc = run_preprocess(c, Slot{0, c->stream}, c->d_raw, 1, w, h, thr, c->forest_dev, nullptr, nullptr);
  if (rc) return rc;
  GPC_CUDA(c, gpc::launch_mask_list(c->d_hash, c->d_rows, c->d_rowoff, w, h, c->d_mask, (int)std::min<size_t>(P, 0x7fffffff), c->stream));
  c->launches += 2;
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_rowoff + h, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  std::vector<uint32_t> himg(P);
  GPC_CUDA(c, cudaMemcpyAsync(himg.data(), c->d_hash, P * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  *n = c->h_counts[0];
  if (hash_image) std::memcpy(hash_image, himg.data(), P * sizeof(uint32_t));
  if (*n > cap && (states || mask)) return fail(c, GPC_E_CAPACITY, "state buffer too small: need " + std::to_string(*n));
  std::vector<int32_t> hmask((size_t)std::max(*n, 1));
  GPC_CUDA(c, cudaMemcpy(hmask.data(), c->d_mask, (size_t)*n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int i = 0; i < *n; i++) {                   // gather in mask order (inference.hpp:282-290)
    if (mask) mask[i] = hmask[i];
    if (states) states[i] = himg[(size_t)hmask[i]] & 0x7fffffffu;
  }
  return GPC_OK;
}

int gpc_match_hash_images(gpc_ctx* c, const uint32_t* hash_l, const uint32_t* hash_r, int w, int h, const gpc_settings* s,
                          gpc_support* out, int cap, int* n_out) {
  if (!c || !hash_l || !hash_r || !n_out || cap < 0 || (cap > 0 && !out)) return fail(c, GPC_E_ARG, "null argument");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  rc = check_settings(c, s); if (rc) return rc;
  GPC_CUDA(c, cudaSetDevice(c->device));
  const size_t P = (size_t)w * h;
  // rowcnt / lastrow are derived on the host here (kernel A normally provides them).  A candidate is by definition an
  // interior pixel (border lambda, inference.hpp:318-330) and every matcher sizes its tables for the interior, so the
  // flag of a word outside 13 <= x < w-13, 13 <= y < h-13 is cleared on the staged copy.
  std::vector<int32_t> rows((size_t)2 * h + 2, 0);
  rows[(size_t)2 * h] = rows[(size_t)2 * h + 1] = -1;
  const uint32_t* src[2] = {hash_l, hash_r};
  std::vector<uint32_t> staged(2 * P);
  const int R = gpc::kRadius;
  for (int k = 0; k < 2; k++)
    for (int y = 0; y < h; y++) {
      int cnt = 0;
      const bool row_in = (y >= R && y < h - R);
      for (int x = 0; x < w; x++) {
        uint32_t v = src[k][(size_t)y * w + x];
        if (!row_in || x < R || x >= w - R) v &= 0x7fffffffu;
        staged[(size_t)k * P + (size_t)y * w + x] = v;
        cnt += (int)(v >> 31);
      }
      rows[(size_t)k * h + y] = cnt;
      if (cnt) rows[(size_t)2 * h + k] = y;
    }
  GPC_CUDA(c, cudaMemcpyAsync(c->d_hash, staged.data(), 2 * P * 4, cudaMemcpyHostToDevice, c->stream));
  GPC_CUDA(c, cudaMemcpyAsync(c->d_rows, rows.data(), (size_t)2 * h * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  GPC_CUDA(c, cudaMemcpyAsync(c->d_lastrow, rows.data() + (size_t)2 * h, 2 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));   // `rows` and `staged` are pageable and about to go out of scope
  rc = run_match(c, Slot{0, c->stream}, 1, w, h, s, c->d_out, c->out_cap, true, c->d_totals, c->d_ncand, true);
  if (rc) return rc;
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_totals, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_out = c->h_counts[0];
  if (*n_out > cap) return fail(c, GPC_E_CAPACITY, "support buffer too small: need " + std::to_string(*n_out));
  if (*n_out > 0) GPC_CUDA(c, cudaMemcpy(out, c->d_out, (size_t)*n_out * sizeof(gpc_support), cudaMemcpyDeviceToHost));
  return GPC_OK;
}

// ---- resident images (the device side of Forest::PreprocessedImage) ---------------------------
int gpc_image_upload(gpc_ctx* c, const uint8_t* img, int w, int h, int stride, gpc_image** out) {
  if (!c || !img || !out) return fail(c, GPC_E_ARG, "null argument");
  *out = nullptr;
  if (stride < w) return fail(c, GPC_E_ARG, "stride smaller than width");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  GPC_CUDA(c, cudaSetDevice(c->device));
  gpc_image* im = new gpc_image();
  im->ctx_id = c->id; im->owner = c; im->device = c->device; im->w = w; im->h = h;
  const size_t raw_bytes = ((size_t)w * h + 255) / 256 * 256;
  im->block_bytes = raw_bytes + im->cache_bytes();
  cudaError_t e = cudaSuccess;
  for (size_t i = 0; i < c->image_pool.size(); i++)
    if (c->image_pool[i].first == im->block_bytes) {
      im->d_raw = c->image_pool[i].second;
      c->image_pool.erase(c->image_pool.begin() + (long)i);
      break;
    }
  if (!im->d_raw) e = cudaMalloc(&im->d_raw, im->block_bytes);
  im->d_cache = im->d_raw ? im->d_raw + raw_bytes : nullptr;
  // the caller's buffer may be pageable and short-lived: copy it into pinned staging on the host (a memcpy), then one
  // asynchronous DMA; nothing waits here -- the next user of the staging buffer waits for ev_stage
  if (e == cudaSuccess && ensure_host_stage(c, (size_t)w * h) != GPC_OK) e = cudaErrorMemoryAllocation;
  if (e == cudaSuccess) {
    for (int y = 0; y < h; y++) std::memcpy(c->h_stage + (size_t)y * w, img + (size_t)y * stride, (size_t)w);
    e = cudaMemcpyAsync(im->d_raw, c->h_stage, (size_t)w * h, cudaMemcpyHostToDevice, c->stream);
  }
  if (e == cudaSuccess) e = cudaEventRecord(c->ev_stage, c->stream);
  if (e != cudaSuccess) {
    cudaFree(im->d_raw);
    delete im;
    return fail(c, GPC_E_CUDA, std::string("gpc_image_upload: ") + cudaGetErrorString(e));
  }
  *out = im;
  return GPC_OK;
}

void gpc_image_release(gpc_image* im) {
  if (!im) return;
  bool pooled = false;
  {
    std::lock_guard<std::mutex> lk(g_registry_mu);
    auto it = g_registry.find(im->ctx_id);
    if (it != g_registry.end() && it->second == im->owner && im->owner->image_pool.size() < 8) {
      // later work of the context is stream-ordered after anything that read this block
      im->owner->image_pool.emplace_back(im->block_bytes, im->d_raw);
      pooled = true;
    }
  }
  if (!pooled) { cudaSetDevice(im->device); cudaFree(im->d_raw); }
  delete im;
}

// Copies between the slot-0 buffers of the context and an image's cache (device to device, on the context's stream).
// slot_img: 0 = left, 1 = right position of the pair in slot 0.
static int cache_copy(gpc_ctx* c, gpc_image* im, int slot_img, bool to_cache, bool a1, bool hash) {
  const int w = im->w, h = im->h;
  const size_t P = (size_t)w * h, CW = (size_t)h * (w / 16);
  auto cp = [&](void* slot_ptr, size_t off, size_t bytes) {
    return to_cache ? cudaMemcpyAsync(im->d_cache + off, slot_ptr, bytes, cudaMemcpyDeviceToDevice, c->stream)
                    : cudaMemcpyAsync(slot_ptr, im->d_cache + off, bytes, cudaMemcpyDeviceToDevice, c->stream);
  };
  if (a1) {
    GPC_CUDA(c, cp(c->d_smooth + (size_t)slot_img * P, 0, P));
    GPC_CUDA(c, cp(c->d_cand + (size_t)slot_img * CW, im->off_cand(), CW * 2));
    GPC_CUDA(c, cp(c->d_rows + (size_t)slot_img * h, im->off_rows(), (size_t)h * 4));
    GPC_CUDA(c, cp(c->d_lastrow + slot_img, im->off_rows() + (size_t)h * 4, 4));
  }
  if (hash) GPC_CUDA(c, cp(c->d_hash + (size_t)slot_img * P, im->off_hash(), P * 4));
  return GPC_OK;
}

// im != nullptr: the image's cache receives kernel A1's output (and the hash image if the context has a forest), so
// that gpc_match_images / gpc_correspond_images need not run the kernels again.
static int preprocess_device(gpc_ctx* c, const uint8_t* d_img, int w, int h, int thr, uint8_t* smooth, uint8_t* grad,
                             int32_t* mask, int mask_cap, int* n_mask, gpc_image* im) {
  int rc = GPC_OK;
  const bool dbg = smooth || grad;
  if (dbg || mask || n_mask) { rc = ensure_debug_buffers(c); if (rc) return rc; }
  const size_t P = (size_t)w * h;
  gpc::ForestDev none{};                           // no tests: hash image carries the candidate flag only
  none.naive = (c->result_mode == GPC_RESULTS_NAIVE) ? gpc::kResultsNaive : gpc::kResultsSse;
  uint8_t* d_smooth = dbg ? c->d_dbg8 : nullptr;
  uint8_t* d_grad = dbg ? c->d_dbg8 + (size_t)c->max_w * c->max_h : nullptr;
  const bool with_forest = im && c->has_forest;
  rc = run_preprocess(c, Slot{0, c->stream}, d_img, 1, w, h, thr, with_forest ? c->forest_dev : none, d_smooth, d_grad);
  if (rc) return rc;
  if (im) {
    rc = cache_copy(c, im, 0, true, true, with_forest); if (rc) return rc;
    im->cache_thr = thr; im->cache_mode = c->result_mode;
    im->hash_serial = with_forest ? c->forest_serial : 0;
  }
  if (mask || n_mask) {
    GPC_CUDA(c, gpc::launch_mask_list(c->d_hash, c->d_rows, c->d_rowoff, w, h, c->d_mask, (int)std::min<size_t>(P, 0x7fffffff), c->stream));
    c->launches += 2;
    GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_rowoff + h, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    // the whole candidate capacity of the interior goes to pinned staging in the same stream (its length is not known
    // on the host yet, the interior bounds it); the caller's copy, or gpc_mask_view, reads the first n entries
    const size_t cap_bytes = (size_t)std::max(w - 2 * gpc::kRadius, 0) * (size_t)std::max(h - 2 * gpc::kRadius, 0) * sizeof(int32_t);
    rc = ensure_host_stage(c, std::max<size_t>(cap_bytes, 4)); if (rc) return rc;
    if (cap_bytes) GPC_CUDA(c, cudaMemcpyAsync(c->h_stage, c->d_mask, cap_bytes, cudaMemcpyDeviceToHost, c->stream));
  }
  if (smooth) GPC_CUDA(c, cudaMemcpyAsync(smooth, d_smooth, P, cudaMemcpyDeviceToHost, c->stream));
  if (grad) GPC_CUDA(c, cudaMemcpyAsync(grad, d_grad, P, cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (mask || n_mask) {
    const int n = c->h_counts[0];
    if (n_mask) *n_mask = n;
    c->last_mask_n = n;
    if (mask) {
      if (n > mask_cap) return fail(c, GPC_E_CAPACITY, "mask buffer too small: need " + std::to_string(n));
      std::memcpy(mask, c->h_stage, (size_t)n * sizeof(int32_t));
    }
  }
  return GPC_OK;
}

int gpc_image_preprocess(gpc_ctx* c, gpc_image* im, int thr, uint8_t* smooth, uint8_t* grad, int32_t* mask,
                         int mask_cap, int* n_mask) {
  if (!c || !im || im->ctx_id != c->id) return fail(c, GPC_E_ARG, "image does not belong to this context");
  int rc = check_dims(c, im->w, im->h, 1); if (rc) return rc;
  if (thr < 0 || thr > 255) return fail(c, GPC_E_ARG, "gradientThreshold needs to be within 0...255");
  GPC_CUDA(c, cudaSetDevice(c->device));
  return preprocess_device(c, im->d_raw, im->w, im->h, thr, smooth, grad, mask, mask_cap, n_mask, im);
}

// The smoothed and gradient images of preprocessImage for a resident image (the C++ API fills
// PreprocessedImage::smooth / grad from here on first access; the kernels run again with their debug outputs on).
int gpc_image_fetch(gpc_ctx* c, const gpc_image* im, int thr, uint8_t* smooth, uint8_t* grad) {
  if (!c || !im || im->ctx_id != c->id) return fail(c, GPC_E_ARG, "image does not belong to this context");
  if (!smooth && !grad) return GPC_OK;
  int rc = check_dims(c, im->w, im->h, 1); if (rc) return rc;
  if (thr < 0 || thr > 255) return fail(c, GPC_E_ARG, "gradientThreshold needs to be within 0...255");
  GPC_CUDA(c, cudaSetDevice(c->device));
  return preprocess_device(c, im->d_raw, im->w, im->h, thr, smooth, grad, nullptr, 0, nullptr, nullptr);
}

// Brings the hash images of two resident images into slot 0 of the context (left = image 0, right = image 1) with as
// little work as their caches allow: hash images computed with the current forest are copied; kernel A1's cached output
// needs kernel A2 only (the result is cached); anything else goes through both kernels from the raw image.
static int resident_pair_hashes(gpc_ctx* c, gpc_image* l, gpc_image* r, const gpc_settings* s) {
  const int w = l->w, h = l->h;
  const size_t P = (size_t)w * h;
  gpc_image* im[2] = {l, r};
  bool a1_ok = true, hash_ok = true;
  for (int k = 0; k < 2; k++) {
    const bool a1 = im[k]->cache_thr >= 0 && im[k]->cache_thr == s->gradient_threshold && im[k]->cache_mode == c->result_mode;
    a1_ok = a1_ok && a1;
    hash_ok = hash_ok && a1 && im[k]->hash_serial == c->forest_serial;
  }
  if (hash_ok) {
    for (int k = 0; k < 2; k++) {
      GPC_CUDA(c, cudaMemcpyAsync(c->d_rows + (size_t)k * h, im[k]->d_cache + im[k]->off_rows(), (size_t)h * 4, cudaMemcpyDeviceToDevice, c->stream));
      GPC_CUDA(c, cudaMemcpyAsync(c->d_lastrow + k, im[k]->d_cache + im[k]->off_rows() + (size_t)h * 4, 4, cudaMemcpyDeviceToDevice, c->stream));
      GPC_CUDA(c, cudaMemcpyAsync(c->d_hash + (size_t)k * P, im[k]->d_cache + im[k]->off_hash(), P * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    return GPC_OK;
  }
  if (a1_ok) {
    for (int k = 0; k < 2; k++) { int rc = cache_copy(c, im[k], k, false, true, false); if (rc) return rc; }
    int rc = run_preprocess(c, Slot{0, c->stream}, c->d_raw, 2, w, h, s->gradient_threshold, c->forest_dev, nullptr, nullptr, nullptr, true);
    if (rc) return rc;
    for (int k = 0; k < 2; k++) {
      rc = cache_copy(c, im[k], k, true, false, true); if (rc) return rc;
      im[k]->hash_serial = c->forest_serial;
    }
    return GPC_OK;
  }
  GPC_CUDA(c, cudaMemcpyAsync(c->d_raw, l->d_raw, P, cudaMemcpyDeviceToDevice, c->stream));
  GPC_CUDA(c, cudaMemcpyAsync(c->d_raw + P, r->d_raw, P, cudaMemcpyDeviceToDevice, c->stream));
  return run_preprocess(c, Slot{0, c->stream}, c->d_raw, 2, w, h, s->gradient_threshold, c->forest_dev, nullptr, nullptr);
}

int gpc_match_images(gpc_ctx* c, gpc_image* l, gpc_image* r, const gpc_settings* s, gpc_support* out, int cap,
                     int* n_out, int* n_cand_l, int* n_cand_r) {
  if (!c || !l || !r || !n_out || cap < 0 || (cap > 0 && !out)) return fail(c, GPC_E_ARG, "null argument");
  if (l->ctx_id != c->id || r->ctx_id != c->id) return fail(c, GPC_E_ARG, "image does not belong to this context");
  if (l->w != r->w || l->h != r->h) return fail(c, GPC_E_DIMS, "left and right image dimensions differ");   // inference.hpp:350-353
  const int w = l->w, h = l->h;
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  rc = check_settings(c, s); if (rc) return rc;
  if (!c->has_forest) return fail(c, GPC_E_FOREST, "no forest set");
  GPC_CUDA(c, cudaSetDevice(c->device));
  rc = resident_pair_hashes(c, l, r, s);
  if (rc) return rc;
  rc = run_match(c, Slot{0, c->stream}, 1, w, h, s, c->d_out, c->out_cap, true, c->d_totals, c->d_ncand);
  if (rc) return rc;
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_totals, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts + 1, c->d_ncand, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_out = c->h_counts[0];
  if (n_cand_l) *n_cand_l = c->h_counts[1];
  if (n_cand_r) *n_cand_r = c->h_counts[2];
  if (*n_out > cap) return fail(c, GPC_E_CAPACITY, "support buffer too small: need " + std::to_string(*n_out));
  if (*n_out > 0) {
    GPC_CUDA(c, cudaMemcpyAsync(out, c->d_out, (size_t)*n_out * sizeof(gpc_support), cudaMemcpyDeviceToHost, c->stream));
    GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return GPC_OK;
}

// Forest::evalFastMaskOnSubsetSSE (inference.hpp:266-292) on a caller-provided smoothed image:
// states[i] = state of pixel idx[i] (= y*w + x).  Entries outside the 13-pixel border, where
// the reference reads across row ends, and rows H-15.. (never hashed, filter.hpp:601-604) give 0.
int gpc_hash_smooth(gpc_ctx* c, const uint8_t* smooth, int w, int h, const int32_t* idx, int n, uint32_t* states) {
  if (!c || !smooth || n < 0 || (n > 0 && (!idx || !states))) return fail(c, GPC_E_ARG, "null argument");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  if (!c->has_forest) return fail(c, GPC_E_FOREST, "no forest set");
  GPC_CUDA(c, cudaSetDevice(c->device));
  rc = ensure_debug_buffers(c); if (rc) return rc;
  const size_t P = (size_t)w * h;
  std::vector<uint8_t> flags(P, 0);
  for (int i = 0; i < n; i++) {
    if (idx[i] < 0 || (size_t)idx[i] >= P) return fail(c, GPC_E_ARG, "candidate index outside the image");
    flags[(size_t)idx[i]] = 255;
  }
  GPC_CUDA(c, cudaMemcpyAsync(c->d_raw, smooth, P, cudaMemcpyHostToDevice, c->stream));
  GPC_CUDA(c, cudaMemcpyAsync(c->d_dbg8, flags.data(), P, cudaMemcpyHostToDevice, c->stream));
  rc = run_preprocess(c, Slot{0, c->stream}, c->d_raw, 1, w, h, 0, c->forest_dev, nullptr, nullptr, c->d_dbg8);
  if (rc) return rc;
  std::vector<uint32_t> himg(P);
  GPC_CUDA(c, cudaMemcpyAsync(himg.data(), c->d_hash, P * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; i++) states[i] = himg[(size_t)idx[i]] & 0x7fffffffu;
  return GPC_OK;
}

// The candidate list the most recent gpc_image_preprocess / gpc_preprocess call left in the context's pinned staging
// buffer (n entries as reported through n_mask); valid until the next call on the context.
const int32_t* gpc_mask_view(gpc_ctx* c, int* n) {
  if (!c || !c->h_stage) { if (n) *n = 0; return nullptr; }
  if (n) *n = c->last_mask_n;
  return reinterpret_cast<const int32_t*>(c->h_stage);
}

// The first n supports of the context's most recent single-pair result (gpc_match_pair / gpc_match_images), still on
// the device: lets a caller that cannot bound the count ask for it first (cap = 0 -> GPC_E_CAPACITY with *n_out set).
int gpc_fetch_supports(gpc_ctx* c, gpc_support* out, int n) {
  if (!c || n < 0 || (n > 0 && !out)) return fail(c, GPC_E_ARG, "null argument");
  if (n > c->out_cap) return fail(c, GPC_E_ARG, "more supports requested than the context holds");
  GPC_CUDA(c, cudaSetDevice(c->device));
  if (n > 0) {                                                  // through pinned staging: `out` is usually pageable
    int rc = ensure_host_stage(c, (size_t)n * sizeof(gpc_support)); if (rc) return rc;
    GPC_CUDA(c, cudaMemcpyAsync(c->h_stage, c->d_out, (size_t)n * sizeof(gpc_support), cudaMemcpyDeviceToHost, c->stream));
    GPC_CUDA(c, cudaStreamSynchronize(c->stream));
    std::memcpy(out, c->h_stage, (size_t)n * sizeof(gpc_support));
    c->last_mask_n = 0;
  }
  return GPC_OK;
}

int gpc_set_matcher(gpc_ctx* c, int matcher) {
  if (!c) return GPC_E_ARG;
  if (matcher != GPC_MATCHER_AUTO && matcher != GPC_MATCHER_SORT && matcher != GPC_MATCHER_ROWS_GENERAL)
    return fail(c, GPC_E_ARG, "unknown matcher");
  c->matcher = matcher;
  return GPC_OK;
}

// Forest::stereoMatch (inference.hpp:344-361): every unique-unique correspondence, before the
// rectifiedMatch filter, in ascending key order.  Always runs the radix-sort matcher.
int gpc_correspond_images(gpc_ctx* c, gpc_image* l, gpc_image* r, const gpc_settings* s,
                          gpc_correspondence* out, int cap, int* n_out) {
  if (!c || !l || !r || !n_out || cap < 0 || (cap > 0 && !out)) return fail(c, GPC_E_ARG, "null argument");
  if (l->ctx_id != c->id || r->ctx_id != c->id) return fail(c, GPC_E_ARG, "image does not belong to this context");
  if (l->w != r->w || l->h != r->h) return fail(c, GPC_E_DIMS, "left and right image dimensions differ");
  const int w = l->w, h = l->h;
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  rc = check_settings(c, s); if (rc) return rc;
  if (!c->has_forest) return fail(c, GPC_E_FOREST, "no forest set");
  GPC_CUDA(c, cudaSetDevice(c->device));
  rc = resident_pair_hashes(c, l, r, s);
  if (rc) return rc;
  // a correspondence is 16 bytes, a support 12: d_out holds out_cap * 12 / 16 correspondences
  const long long dcap = c->out_cap * 12 / 16;
  rc = run_match_sort(c, c->d_hash, 0, 1, w, h, s, 1, c->d_out, sizeof(gpc_correspondence), 0, dcap, c->d_totals, nullptr, c->stream, nullptr);
  if (rc) return rc;
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_totals, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_out = c->h_counts[0];
  if (*n_out > cap || *n_out > dcap) return fail(c, GPC_E_CAPACITY, "correspondence buffer too small: need " + std::to_string(*n_out));
  if (*n_out > 0) GPC_CUDA(c, cudaMemcpy(out, c->d_out, (size_t)*n_out * sizeof(gpc_correspondence), cudaMemcpyDeviceToHost));
  return GPC_OK;
}

// Forest::findCorrespondences (inference.hpp:227-254) on explicit 64-bit keys.  out_pairs[2i],
// out_pairs[2i+1] = indices into src / tar (original order) of match i, ascending src key.
int gpc_find_correspondences(gpc_ctx* c, const uint64_t* src_keys, int n_src, const uint64_t* tar_keys, int n_tar,
                             int32_t* out_pairs, int cap, int* n_out) {
  if (!c || !n_out || n_src < 0 || n_tar < 0 || cap < 0 || (cap > 0 && !out_pairs)) return fail(c, GPC_E_ARG, "null argument");
  *n_out = 0;
  if (n_src == 0 || n_tar == 0) return GPC_OK;                 // len(tar) == 0 is UB in the reference; defined as no matches
  if (!src_keys || !tar_keys) return fail(c, GPC_E_ARG, "null argument");
  if ((long long)n_src + n_tar >= 0x7fffffffll) return fail(c, GPC_E_DIMS, "too many descriptors");
  GPC_CUDA(c, cudaSetDevice(c->device));
  const long long n = (long long)n_src + n_tar;
  int rc = ensure_global_ws(c, gpc::global_workspace_bytes(n + 2, 1, 0, 0)); if (rc) return rc;
  uint64_t kmax = 0;
  for (int i = 0; i < n_src; i++) kmax = std::max(kmax, src_keys[i]);
  for (int i = 0; i < n_tar; i++) kmax = std::max(kmax, tar_keys[i]);
  int key_bits = 8;
  while (key_bits < 64 && (kmax >> key_bits) != 0) key_bits += 8;
  uint64_t* d_keys = reinterpret_cast<uint64_t*>(gpc::global_key_buffer(c->d_gws, n + 2));
  GPC_CUDA(c, cudaMemcpyAsync(d_keys, src_keys, (size_t)n_src * 8, cudaMemcpyHostToDevice, c->stream));
  GPC_CUDA(c, cudaMemcpyAsync(d_keys + n_src, tar_keys, (size_t)n_tar * 8, cudaMemcpyHostToDevice, c->stream));
  const long long need = std::min<long long>(n_src, n_tar);
  int32_t* d_pairs = nullptr;
  GPC_CUDA(c, cudaMalloc(&d_pairs, (size_t)std::max<long long>(need, 1) * 2 * sizeof(int32_t)));
  int launches = 0;
  cudaError_t e = gpc::launch_match_keys(c->d_gws, n + 2, n_src, n_tar, key_bits, d_pairs, need, c->d_totals, c->stream, &launches);
  c->launches += launches;
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_counts, c->d_totals, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) {
    *n_out = c->h_counts[0];
    if (*n_out <= cap && *n_out > 0) e = cudaMemcpy(out_pairs, d_pairs, (size_t)*n_out * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost);
  }
  cudaFree(d_pairs);
  if (e != cudaSuccess) return fail(c, GPC_E_CUDA, std::string("gpc_find_correspondences: ") + cudaGetErrorString(e));
  if (*n_out > cap) return fail(c, GPC_E_CAPACITY, "pair buffer too small: need " + std::to_string(*n_out));
  return GPC_OK;
}

// ndb::Hashmatch as depthPriorFast drives it with useHashtable (inference.hpp:204-225, hashmatch.hpp:48-272)
// on explicit 64-bit keys: all src keys inserted, then all tar keys; pairs in bucket order, then list order.
int gpc_hashmatch(gpc_ctx* c, const uint64_t* src_keys, int n_src, const uint64_t* tar_keys, int n_tar,
                  int32_t* out_pairs, int cap, int* n_out) {
  if (!c || !n_out || n_src < 0 || n_tar < 0 || cap < 0 || (cap > 0 && !out_pairs)) return fail(c, GPC_E_ARG, "null argument");
  *n_out = 0;
  if (n_src == 0 || n_tar == 0) return GPC_OK;                 // a match needs one element of each list
  if (!src_keys || !tar_keys) return fail(c, GPC_E_ARG, "null argument");
  if ((long long)n_src + n_tar >= 0x7fffffffll) return fail(c, GPC_E_DIMS, "too many descriptors");
  GPC_CUDA(c, cudaSetDevice(c->device));
  const long long n = (long long)n_src + n_tar;
  int rc = ensure_global_ws(c, gpc::global_workspace_bytes(n + 2, 1, 0, 0)); if (rc) return rc;
  const long long need = std::min<long long>(n_src, n_tar);
  uint8_t* d_tmp = nullptr;                                    // keys, then the output pairs
  GPC_CUDA(c, cudaMalloc(&d_tmp, (size_t)n * 8 + (size_t)std::max<long long>(need, 1) * 2 * sizeof(int32_t)));
  unsigned long long* d_keys = reinterpret_cast<unsigned long long*>(d_tmp);
  int32_t* d_pairs = reinterpret_cast<int32_t*>(d_tmp + (size_t)n * 8);
  cudaError_t e = cudaMemcpyAsync(d_keys, src_keys, (size_t)n_src * 8, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_keys + n_src, tar_keys, (size_t)n_tar * 8, cudaMemcpyHostToDevice, c->stream);
  int launches = 0;
  if (e == cudaSuccess) e = gpc::launch_hashmatch_keys(c->d_gws, n + 2, d_keys, n_src, n_tar, d_pairs, need, c->d_totals, c->stream, &launches);
  c->launches += launches;
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_counts, c->d_totals, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) {
    *n_out = c->h_counts[0];
    if (*n_out <= cap && *n_out > 0) e = cudaMemcpy(out_pairs, d_pairs, (size_t)*n_out * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost);
  }
  cudaFree(d_tmp);
  if (e != cudaSuccess) return fail(c, GPC_E_CUDA, std::string("gpc_hashmatch: ") + cudaGetErrorString(e));
  if (*n_out > cap) return fail(c, GPC_E_CAPACITY, "pair buffer too small: need " + std::to_string(*n_out));
  return GPC_OK;
}

// Multi-level matching (BASELINE.json configs[3]; definition in SURVEY.md 8d -- the reference has
// no pyramid): level l+1 = 2x2 floor-mean of the raw level-l images, built on the device; the whole
// single-level path runs per level with the same forest and disp_high halved per level.  Supports
// of all levels are written back to back, level l in out[level_offsets[l] .. level_offsets[l+1]);
// n_cand (optional) = [n_levels][2].  Every level width must stay a multiple of 16.
int gpc_match_pyramid(gpc_ctx* c, const uint8_t* left, const uint8_t* right, int w, int h, int stride, int n_levels,
                      const gpc_settings* s, gpc_support* out, int64_t cap, int64_t* level_offsets, int32_t* n_cand) {
  if (!c || !left || !right || !level_offsets || cap < 0 || (cap > 0 && !out)) return fail(c, GPC_E_ARG, "null argument");
  if (n_levels < 1 || n_levels > 16) return fail(c, GPC_E_ARG, "n_levels must be within 1...16");
  if (stride < w) return fail(c, GPC_E_ARG, "stride smaller than width");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  rc = check_settings(c, s); if (rc) return rc;
  if (!c->has_forest) return fail(c, GPC_E_FOREST, "no forest set");
  for (int l = 1; l < n_levels; l++)
    if (((w >> l) % 16) != 0 || (w >> l) <= 0 || (h >> l) <= 0)
      return fail(c, GPC_E_WIDTH16, "every pyramid level needs a positive width that is a multiple of 16");
  GPC_CUDA(c, cudaSetDevice(c->device));
  const size_t P0 = (size_t)w * h;
  if (n_levels > 1 && !c->d_pyr) GPC_CUDA(c, cudaMalloc(&c->d_pyr, (size_t)c->max_w * c->max_h));   // 2 * (P/4 + P/16 + ...) < P
  GPC_CUDA(c, cudaMemcpy2DAsync(c->d_raw, w, left, stride, w, h, cudaMemcpyHostToDevice, c->stream));
  GPC_CUDA(c, cudaMemcpy2DAsync(c->d_raw + P0, w, right, stride, w, h, cudaMemcpyHostToDevice, c->stream));
  const uint8_t* level = c->d_raw;
  uint8_t* next = c->d_pyr;
  long long total = 0;
  bool overflow = false;
  level_offsets[0] = 0;
  gpc_settings ls = *s;
  for (int l = 0; l < n_levels; l++) {
    const int lw = w >> l, lh = h >> l;
    rc = run_preprocess(c, Slot{0, c->stream}, level, 2, lw, lh, ls.gradient_threshold, c->forest_dev, nullptr, nullptr);
    if (rc) return rc;
    rc = run_match(c, Slot{0, c->stream}, 1, lw, lh, &ls, c->d_out, c->out_cap, true, c->d_totals, c->d_ncand);
    if (rc) return rc;
    if (l + 1 < n_levels) {                                     // next level from this level's raw images
      GPC_CUDA(c, gpc::launch_downsample2x(level, next, lw, lh, 2, c->stream));
      c->launches += 1;
    }
    GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_totals, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    GPC_CUDA(c, cudaMemcpyAsync(c->h_counts + 1, c->d_ncand, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    GPC_CUDA(c, cudaStreamSynchronize(c->stream));
    const int n = c->h_counts[0];
    if (n_cand) { n_cand[2 * l] = c->h_counts[1]; n_cand[2 * l + 1] = c->h_counts[2]; }
    if (total + n > cap) overflow = true;
    else if (n > 0) GPC_CUDA(c, cudaMemcpyAsync(out + total, c->d_out, (size_t)n * sizeof(gpc_support), cudaMemcpyDeviceToHost, c->stream));
    total += n;
    level_offsets[l + 1] = total;
    if (l + 1 < n_levels) {
      level = next;
      next += 2 * (size_t)(lw / 2) * (lh / 2);
      ls.disp_high = ls.disp_high / 2;
    }
  }
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (overflow) return fail(c, GPC_E_CAPACITY, "support buffer too small: need " + std::to_string(total));
  return GPC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Forests of more than 32 tests ("extended mode"; no reference semantics -- the reference keeps the
// first 32 tests, inference.hpp:426) and 32-test forests in GPC_RESULTS_NAIVE (32 state bits + the candidate flag do not
// fit one hash word).  The state of a candidate is a tuple of words:
//   SSE results   : word k = tests 32k .. 32k+31 evaluated as a forest of their own (each word with the bit placement
//                   of filter.hpp:574-584), so word 0 is exactly the reference's truncated 32-test state
//   naive results : the T-bit state of gpcFilter[Tau]Naive (test t in bit T-1-t, filter.hpp:245-293) cut into 31-bit words
// Kernel A2 runs once per word on kernel A1's output; the planes are folded pairwise into dense ranks (match_global.cu:
// launch_wide_rank -- exact and order preserving) and the ordinary matchers run on the final rank plane.  Output order:
// ascending (y,) then the state tuple compared from the LAST word down (for the naive mode: the numeric T-bit state).
// ------------------------------------------------------------------------------------------------------------------
static int run_wide_pair(gpc_ctx* c, int w, int h, const gpc_settings* s) {
  const size_t P = (size_t)w * h;
  const int K = (int)c->wide_words.size();
  const size_t plane = 2 * P;                                        // one word of both images
  const size_t need = (size_t)(2 * K + 1) * plane * sizeof(uint32_t);
  if (need > c->wide_bytes) {
    GPC_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_wide); c->d_wide = nullptr; c->wide_bytes = 0;
    GPC_CUDA(c, cudaMalloc(&c->d_wide, need));
    c->wide_bytes = need;
  }
  const long long records = 2ll * std::max(w - 2 * gpc::kRadius, 0) * std::max(h - 2 * gpc::kRadius, 0) + 2;
  int rc = ensure_global_ws(c, gpc::wide_workspace_bytes(records)); if (rc) return rc;
  // kernel A1 once, kernel A2 once per word
  for (int k = 0; k < K; k++) {
    rc = run_preprocess(c, Slot{0, c->stream}, c->d_raw, 2, w, h, s->gradient_threshold, c->wide_words[(size_t)k], nullptr, nullptr, nullptr,
                        k > 0, c->d_wide + (size_t)k * plane);
    if (rc) return rc;
  }
  // fold the planes pairwise until one is left: (lo, hi) -> dense rank of the pair
  std::vector<uint32_t*> cur((size_t)K);
  for (int k = 0; k < K; k++) cur[(size_t)k] = c->d_wide + (size_t)k * plane;
  int next_free = K;
  while (cur.size() > 1) {
    std::vector<uint32_t*> nxt;
    for (size_t j = 0; j + 1 < cur.size(); j += 2) {
      uint32_t* out = c->d_wide + (size_t)next_free * plane;
      next_free = (next_free + 1 < 2 * K + 1) ? next_free + 1 : K;   // K + (K - 1) folds at most: never wraps onto a live plane
      GPC_CUDA(c, cudaMemsetAsync(out, 0, plane * sizeof(uint32_t), c->stream));
      int launches = 0;
      GPC_CUDA(c, gpc::launch_wide_rank(c->d_gws, records, cur[j + 1], cur[j], (long long)plane, 64, out, c->stream, &launches));
      c->launches += launches;
      nxt.push_back(out);
    }
    if (cur.size() & 1) nxt.push_back(cur.back());
    cur.swap(nxt);
  }
  GPC_CUDA(c, cudaMemcpyAsync(c->d_hash, cur[0], plane * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
  return run_match(c, Slot{0, c->stream}, 1, w, h, s, c->d_out, c->out_cap, true, c->d_totals, c->d_ncand, true);
}

extern "C" {

// tests: n_tests rows of {ix, iy, jx, jy, tau} in file order.  The result mode in force decides the word layout; calling
// gpc_set_result_mode afterwards invalidates the wide forest (set it again).
int gpc_set_wide_forest(gpc_ctx* c, const int32_t* tests, int n_tests) {
  if (!c || !tests) return GPC_E_ARG;
  if (n_tests < 1 || n_tests > GPC_MAX_WIDE_TESTS) return fail(c, GPC_E_FOREST, "a wide forest has 1 .. 256 tests");
  int type = 0;
  for (int t = 0; t < n_tests; t++) {
    for (int k = 0; k < 4; k++)
      if (tests[5 * t + k] < -GPC_PATCH_RADIUS || tests[5 * t + k] > GPC_PATCH_RADIUS)
        return fail(c, GPC_E_FOREST, "test offset outside the 27x27 patch (|offset| <= 13)");
    if (tests[5 * t + 4] != 0) type = 1;
  }
  const bool naive = (c->result_mode == GPC_RESULTS_NAIVE);
  const int per = naive ? 31 : 32;
  const int K = (n_tests + per - 1) / per;
  c->wide_words.assign((size_t)K, gpc::ForestDev{});
  for (int k = 0; k < K; k++) {
    // SSE: tests 32k ..; naive: word k holds state bits 31k .. 31k+30 = tests T-31k-Tk .. T-31k-1, in file order
    const int tk = std::min(per, n_tests - per * k);
    const int t0 = naive ? n_tests - per * k - tk : per * k;
    gpc_forest f;
    std::memset(&f, 0, sizeof(f));
    f.n_tests = tk; f.type = type;
    for (int j = 0; j < tk; j++) {
      const int32_t* src = tests + 5 * (size_t)(t0 + j);
      f.ix[j] = src[0]; f.iy[j] = src[1]; f.jx[j] = src[2]; f.jy[j] = src[3]; f.tau[j] = src[4];
    }
    bake_forest(f, &c->wide_words[(size_t)k], c->result_mode);
  }
  c->wide_tests = n_tests;
  return GPC_OK;
}

// gpc_match_pair with the wide forest: same arguments and result layout.
int gpc_match_pair_wide(gpc_ctx* c, const uint8_t* left, const uint8_t* right, int w, int h, int stride, const gpc_settings* s,
                        gpc_support* out, int cap, int* n_out, int* n_cand_l, int* n_cand_r) {
  if (!c || !left || !right || !n_out || cap < 0 || (cap > 0 && !out)) return fail(c, GPC_E_ARG, "null argument");
  if (stride < w) return fail(c, GPC_E_ARG, "stride smaller than width");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  rc = check_settings(c, s); if (rc) return rc;
  if (c->wide_words.empty()) return fail(c, GPC_E_FOREST, "no wide forest set");
  if (s->use_hashtable) return fail(c, GPC_E_UNSUPPORTED, "the hashtable matcher is defined on single-word states only");
  GPC_CUDA(c, cudaSetDevice(c->device));
  const size_t P = (size_t)w * h;
  GPC_CUDA(c, cudaMemcpy2DAsync(c->d_raw, w, left, stride, w, h, cudaMemcpyHostToDevice, c->stream));
  GPC_CUDA(c, cudaMemcpy2DAsync(c->d_raw + P, w, right, stride, w, h, cudaMemcpyHostToDevice, c->stream));
  rc = run_wide_pair(c, w, h, s);
  if (rc) return rc;
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts, c->d_totals, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaMemcpyAsync(c->h_counts + 1, c->d_ncand, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_out = c->h_counts[0];
  if (n_cand_l) *n_cand_l = c->h_counts[1];
  if (n_cand_r) *n_cand_r = c->h_counts[2];
  if (*n_out > cap) return fail(c, GPC_E_CAPACITY, "support buffer too small: need " + std::to_string(*n_out));
  if (*n_out > 0) {
    GPC_CUDA(c, cudaMemcpyAsync(out, c->d_out, (size_t)*n_out * sizeof(gpc_support), cudaMemcpyDeviceToHost, c->stream));
    GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return GPC_OK;
}

// The hash words of one image under the wide forest (stage seam for parity tests): words[k * h * w + y * w + x] =
// candidate flag | word k of the pixel's state, 0 for non-candidates.
int gpc_hash_wide(gpc_ctx* c, const uint8_t* img, int w, int h, int thr, uint32_t* words, int n_words_cap, int* n_words) {
  if (!c || !img || !words || !n_words) return fail(c, GPC_E_ARG, "null argument");
  int rc = check_dims(c, w, h, 1); if (rc) return rc;
  if (c->wide_words.empty()) return fail(c, GPC_E_FOREST, "no wide forest set");
  if (thr < 0 || thr > 255) return fail(c, GPC_E_ARG, "gradientThreshold needs to be within 0...255");
  const int K = (int)c->wide_words.size();
  *n_words = K;
  if (n_words_cap < K) return fail(c, GPC_E_CAPACITY, "word buffer too small: need " + std::to_string(K));
  GPC_CUDA(c, cudaSetDevice(c->device));
  const size_t P = (size_t)w * h;
  GPC_CUDA(c, cudaMemcpyAsync(c->d_raw, img, P, cudaMemcpyHostToDevice, c->stream));
  for (int k = 0; k < K; k++) {
    rc = run_preprocess(c, Slot{0, c->stream}, c->d_raw, 1, w, h, thr, c->wide_words[(size_t)k], nullptr, nullptr, nullptr, k > 0);
    if (rc) return rc;
    GPC_CUDA(c, cudaMemcpyAsync(words + (size_t)k * P, c->d_hash, P * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  GPC_CUDA(c, cudaStreamSynchronize(c->stream));
  return GPC_OK;
}

// The whole test list of a forest file (gpc_read_forest's parser without the 32-test cap): tests = rows of
// {ix, iy, jx, jy, tau}; *n_tests receives the number of tests in the file even if it exceeds cap.
int gpc_read_forest_tests(const char* path, int32_t* tests, int cap, int* n_tests, int* n_ferns) {
  if (!path || !n_tests || cap < 0 || (cap > 0 && !tests)) return GPC_E_ARG;
  std::ifstream ff(path);
  if (ff.fail()) return GPC_E_IO;
  int num_ferns = 0, n = 0;
  ff >> num_ferns;
  if (n_ferns) *n_ferns = num_ferns;
  for (int i = 0; i < num_ferns && ff.good(); i++) {
    int id = 0, nt = 0;
    std::string scale;
    ff >> id >> scale >> nt;
    for (int j = 0; j < nt; j++) {
      int lvl = 0, v[5] = {0, 0, 0, 0, 0};
      ff >> lvl >> v[0] >> v[1] >> v[2] >> v[3] >> v[4];
      if (ff.fail()) break;
      if (n < cap) for (int k = 0; k < 5; k++) tests[5 * (size_t)n + k] = v[k];
      n++;
    }
  }
  *n_tests = n;
  return GPC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// gpc_pool: the library-level multi-GPU driver.  One resident context and one host thread
// per GPU; a batch is cut into chunks that are dealt round-robin to the devices, every device pipelines its chunks
// (upload / kernels / download) and the supports land, in pair order, directly in the caller's buffer (ChunkBoard).
// Pairs are independent units: there is no device-to-device traffic and no collective.
// ------------------------------------------------------------------------------------------------------------------
struct gpc_pool {
  std::vector<gpc_ctx*> ctx;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  uint64_t job_serial = 0;
  int pending = 0;
  bool quit = false;
  std::function<int(int)> job;          // worker index -> status
  std::vector<int> rc;
  std::string err;
  int max_w = 0, max_h = 0, max_batch = 0;

  void worker_loop(int g) {
    uint64_t seen = 0;
    for (;;) {
      std::function<int(int)> fn;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return quit || job_serial != seen; });
        if (quit) return;
        seen = job_serial;
        fn = job;
      }
      const int r = fn(g);
      {
        std::lock_guard<std::mutex> lk(mu);
        rc[(size_t)g] = r;
        pending--;
      }
      cv_done.notify_all();
    }
  }
  // runs fn(worker) on every worker thread and waits; returns the first non-zero status
  int run(std::function<int(int)> fn) {
    {
      std::lock_guard<std::mutex> lk(mu);
      job = std::move(fn);
      pending = (int)ctx.size();
      job_serial++;
    }
    cv_job.notify_all();
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return pending == 0; });
    for (size_t g = 0; g < ctx.size(); g++)
      if (rc[g] != GPC_OK) { err = "device " + std::to_string(ctx[g]->device) + ": " + ctx[g]->err; return rc[g]; }
    return GPC_OK;
  }
};

extern "C" {

int gpc_pool_create(gpc_pool** out, const int* devices, int n_devices, int max_w, int max_h, int max_batch_per_device) {
  if (!out) return fail(nullptr, GPC_E_ARG, "out is NULL");
  *out = nullptr;
  if (!devices || n_devices <= 0 || n_devices > 64) return fail(nullptr, GPC_E_ARG, "need 1..64 devices");
  gpc_pool* p = new gpc_pool();
  p->max_w = max_w; p->max_h = max_h; p->max_batch = max_batch_per_device;
  for (int g = 0; g < n_devices; g++) {
    gpc_ctx* c = nullptr;
    const int rc = gpc_create(&c, devices[g], max_w, max_h, max_batch_per_device);
    if (rc != GPC_OK) {
      for (gpc_ctx* d : p->ctx) gpc_destroy(d);
      delete p;
      return rc;                                       // gpc_last_error(NULL) has the text
    }
    p->ctx.push_back(c);
  }
  p->rc.assign((size_t)n_devices, GPC_OK);
  for (int g = 0; g < n_devices; g++) p->workers.emplace_back([p, g] { p->worker_loop(g); });
  *out = p;
  return GPC_OK;
}

void gpc_pool_destroy(gpc_pool* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->quit = true;
  }
  p->cv_job.notify_all();
  for (std::thread& t : p->workers) t.join();
  for (gpc_ctx* c : p->ctx) gpc_destroy(c);
  delete p;
}

int gpc_pool_size(const gpc_pool* p) { return p ? (int)p->ctx.size() : 0; }
gpc_ctx* gpc_pool_context(gpc_pool* p, int i) { return (p && i >= 0 && i < (int)p->ctx.size()) ? p->ctx[(size_t)i] : nullptr; }
const char* gpc_pool_last_error(const gpc_pool* p) { return p ? p->err.c_str() : g_create_error.c_str(); }

int64_t gpc_pool_launch_count(const gpc_pool* p) {
  int64_t n = 0;
  if (p) for (gpc_ctx* c : p->ctx) n += c->launches;
  return n;
}

// every device bakes (and, with NVRTC, specialises) the forest on its own thread
int gpc_pool_set_forest(gpc_pool* p, const gpc_forest* f) {
  if (!p || !f) return GPC_E_ARG;
  return p->run([p, f](int g) { return gpc_set_forest(p->ctx[(size_t)g], f); });
}

int gpc_pool_set_result_mode(gpc_pool* p, int mode) {
  if (!p) return GPC_E_ARG;
  return p->run([p, mode](int g) { return gpc_set_result_mode(p->ctx[(size_t)g], mode); });
}

// gpc_match_batch over all devices of the pool: same arguments, same result (pair order, packed supports).
int gpc_pool_match_batch(gpc_pool* p, const uint8_t* images, int n_pairs, int w, int h, const gpc_settings* s,
                         gpc_support* out, int64_t cap, int64_t* offsets, int32_t* n_cand) {
  if (!p) return GPC_E_ARG;
  auto bad = [p](int code, const char* msg) { p->err = msg; return code; };
  if (!images || !offsets || cap < 0 || (cap > 0 && !out) || n_pairs <= 0) return bad(GPC_E_ARG, "null argument");
  if (!s) return bad(GPC_E_ARG, "settings is NULL");
  const int G = (int)p->ctx.size();
  for (gpc_ctx* c : p->ctx) {
    int rc = check_dims(c, w, h, 1); if (rc) { p->err = c->err; return rc; }
    rc = check_settings(c, s); if (rc) { p->err = c->err; return rc; }
    if (!c->has_forest) return bad(GPC_E_FOREST, "no forest set");
  }
  if (h <= 2 * gpc::kRadius) {                         // nothing can match: no kernels needed
    for (int i = 0; i <= n_pairs; i++) offsets[i] = 0;
    if (n_cand) std::memset(n_cand, 0, 2 * (size_t)n_pairs * sizeof(int32_t));
    return GPC_OK;
  }
  ChunkBoard board;
  const int CH = std::max(1, std::min(chunk_for(p->ctx[0], n_pairs / G), std::max(2, n_pairs / (4 * G))));
  board.plan(n_pairs, CH);
  std::vector<std::vector<int>> mine((size_t)G);
  for (int k = 0; k < board.n_chunks(); k++) mine[(size_t)(k % G)].push_back(k);
  offsets[0] = 0;
  std::vector<char> overflow((size_t)G, 0);
  const int rc = p->run([&](int g) {
    gpc_ctx* c = p->ctx[(size_t)g];
    if (mine[(size_t)g].empty()) return (int)GPC_OK;
    cudaError_t e = cudaSetDevice(c->device);
    int r = GPC_OK;
    if (e != cudaSuccess) r = fail(c, GPC_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    if (r == GPC_OK && use_sort_matcher(c, s)) r = ensure_global_ws(c, sort_workspace_bytes(w, h, CH, nullptr));
    bool ov = false;
    if (r == GPC_OK) r = run_chunks(c, board, mine[(size_t)g], images, w, h, s, out, cap, offsets, n_cand, &ov);
    overflow[(size_t)g] = ov ? 1 : 0;
    if (r != GPC_OK) board.fail_all();
    return r;
  });
  if (rc) return rc;
  for (char o : overflow)
    if (o) return bad(GPC_E_CAPACITY, ("support buffer too small: need " + std::to_string(offsets[n_pairs])).c_str());
  return GPC_OK;
}

// Page-locked host memory every device of the machine can DMA from / to (cudaHostAllocPortable): what `images`
// and `out` of the batch calls should live in.  Plain malloc'ed memory works too, at a fraction of the speed.
void* gpc_host_alloc(size_t bytes) {
  void* p = nullptr;
  return cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess ? p : nullptr;
}
void gpc_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"

/*
 * gpc_b200.h -- C ABI of libgpc_b200.so: the B200 (sm_100a) implementation of openGPC's
 * Global Patch Collider inference path (the window timed by samples/sparsematch.cpp:45-52).
 *
 * The reference has no FFI of its own (header-only C++); this ABI is the seam a binding would
 * sit on.  Each entry point names the reference interface it replaces (paths relative to the
 * openGPC tree).  The drop-in C++ headers include/gpc/{inference,buffer}.hpp forward to it.
 *
 * Conventions
 *  - plain pointers and sizes only, POD structs, no exceptions or aborts across the ABI;
 *  - every function returns a gpc_status (0 = ok); gpc_last_error() gives the text;
 *  - images are uint8 row-major with width % 16 == 0 (reference: filter.hpp:294,405,549,621);
 *  - a gpc_ctx owns one CUDA device's resident buffers + one stream; it is not thread-safe,
 *    distinct contexts are independent (one per GPU / host thread);
 *  - outputs are caller-owned with an explicit capacity; GPC_E_CAPACITY reports the need;
 *  - there is no CPU fallback: without a CUDA device gpc_create fails with GPC_E_CUDA.
 */
#ifndef GPC_B200_H
#define GPC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPC_MAX_TESTS 32   /* inference.hpp:426 */
#define GPC_PATCH_RADIUS 13 /* inference.hpp:322 border; shipped forests stay within +-13 */

typedef enum {
  GPC_OK = 0,
  GPC_E_ARG = 1,          /* null pointer / negative size */
  GPC_E_WIDTH16 = 2,      /* width not a multiple of 16 (assert at filter.hpp:294) */
  GPC_E_DIMS = 3,         /* exceeds the context's max_w / max_h / max_batch */
  GPC_E_CUDA = 4,         /* CUDA runtime error or no device */
  GPC_E_CAPACITY = 5,     /* output buffer too small; counts still report the need */
  GPC_E_UNSUPPORTED = 6,  /* a combination this library does not implement (see the entry point's comment) */
  GPC_E_FOREST = 7,       /* no forest set / more than 32 tests / offset outside +-13 */
  GPC_E_IO = 8            /* forest file cannot be opened (inference.hpp:409-412) */
} gpc_status;

/* == ndb::Support (buffer.hpp:86-92): 12 bytes */
typedef struct { int32_t x, y; float d; } gpc_support;

/* mirrors gpc::inference::InferenceSettings (inference.hpp:71-131) */
typedef struct {
  int32_t gradient_threshold;   /* gradientThreshold_, uint8 range */
  int32_t disp_high;            /* dispHigh_ */
  int32_t vertical_tolerance;   /* verticalTolerance_ */
  int32_t epipolar_mode;        /* epipolarMode_ */
  int32_t use_hashtable;        /* useHashtable_: 1 = the reference's hashtable matcher (inference.hpp:204-225,
                                   hashmatch.hpp:48-272), reproduced bucket by bucket; a different, smaller set */
  int32_t num_threads;          /* numThreads_: accepted, ignored */
} gpc_settings;

/* result of Forest::readForest (inference.hpp:404-446) before the per-width offset baking */
typedef struct {
  int32_t n_tests;              /* min(#tests in file, 32) */
  int32_t type;                 /* 0 zero forest, 1 tau forest */
  int32_t n_discarded;          /* tests dropped by the 32-test cap (one "Note:" line each) */
  int32_t ix[GPC_MAX_TESTS], iy[GPC_MAX_TESTS], jx[GPC_MAX_TESTS], jy[GPC_MAX_TESTS];
  int32_t tau[GPC_MAX_TESTS];
  int32_t n_ferns;              /* first number of the file ("number of ferns:" line, inference.hpp:417) */
} gpc_forest;

typedef struct gpc_ctx gpc_ctx;

/* ---- lifetime --------------------------------------------------------------------------- */
/* One resident context per GPU.  Buffers are sized for max_batch pairs of max_w x max_h. */
int gpc_create(gpc_ctx** out, int device, int max_w, int max_h, int max_batch);
void gpc_destroy(gpc_ctx* ctx);
const char* gpc_last_error(const gpc_ctx* ctx);   /* ctx may be NULL: last create error */
const char* gpc_status_string(int status);
/* Run all work of this context on an existing CUDA stream (cudaStream_t); NULL restores the
 * context's own stream. */
int gpc_set_stream(gpc_ctx* ctx, void* cuda_stream);
int gpc_synchronize(gpc_ctx* ctx);

/* ---- forest (replaces Forest::readForest, inference.hpp:404-446) ------------------------ */
/* Host-side parser of the reference's text format; no device work. */
int gpc_read_forest(const char* path, gpc_forest* out);
int gpc_set_forest(gpc_ctx* ctx, const gpc_forest* forest);

/* ---- whole path, host buffers (replaces preprocessImage x2 + rectifiedMatch,
 *      inference.hpp:302, :375; the sparsematch.cpp:46-51 window) ------------------------- */
int gpc_match_pair(gpc_ctx* ctx, const uint8_t* left, const uint8_t* right, int w, int h, int stride,
                   const gpc_settings* s, gpc_support* out, int cap, int* n_out,
                   int* n_cand_l, int* n_cand_r);

/* Batch of independent pairs, host buffers.  images = [n_pairs][2][h][w] (left, right).
 * out receives the pairs' support lists back to back; offsets[i]..offsets[i+1] delimit pair i
 * (offsets has n_pairs+1 entries).  n_cand (optional) = [n_pairs][2]. */
int gpc_match_batch(gpc_ctx* ctx, const uint8_t* images, int n_pairs, int w, int h,
                    const gpc_settings* s, gpc_support* out, int64_t cap, int64_t* offsets,
                    int32_t* n_cand);

/* Multi-level matching (BASELINE.json configs[3]).  The reference has no pyramid; SURVEY.md 8d
 * defines it: level l+1 = 2x2 floor-mean (a+b+c+d)/4 of the raw level-l images (built on the
 * device), the single-level path of inference.hpp:302-393 per level with the same forest, and
 * disp_high halved per level.  Level l's supports are out[level_offsets[l]..level_offsets[l+1]);
 * level_offsets has n_levels+1 entries; n_cand (optional) = [n_levels][2]. */
int gpc_match_pyramid(gpc_ctx* ctx, const uint8_t* left, const uint8_t* right, int w, int h, int stride,
                      int n_levels, const gpc_settings* s, gpc_support* out, int64_t cap,
                      int64_t* level_offsets, int32_t* n_cand);

/* ---- whole path, device-resident, stream-ordered (no host synchronisation) ---------------
 * d_images = [n_pairs][2][h][w] uint8 in device memory; d_out = [n_pairs][cap_per_pair]
 * gpc_support in device memory; d_n_out = [n_pairs] int32 (true counts, may exceed the
 * capacity, in which case that pair's list is truncated); d_n_cand = [n_pairs][2] or NULL. */
int gpc_match_batch_device(gpc_ctx* ctx, const uint8_t* d_images, int n_pairs, int w, int h,
                           const gpc_settings* s, gpc_support* d_out, int cap_per_pair,
                           int32_t* d_n_out, int32_t* d_n_cand);

/* ---- stage-level seams for parity tests -------------------------------------------------- */
/* ndb::box + clearBoundary, ndb::sobel, ndb::arr2ind + border filter (filter.hpp:293, :404,
 * :60; inference.hpp:302-333).  smooth / grad = [h][w] uint8 or NULL; mask = candidate indices
 * y*w+x in raster order (capacity mask_cap) or NULL. */
int gpc_preprocess(gpc_ctx* ctx, const uint8_t* img, int w, int h, int gradient_threshold,
                   uint8_t* smooth, uint8_t* grad, int32_t* mask, int mask_cap, int* n_mask);
/* Forest::evalFastMaskOnSubsetSSE (inference.hpp:266-292): one state per candidate, in mask
 * order.  hash_image (optional) = the raw [h][w] uint32 image the kernel writes: bit 31 marks
 * a candidate, bits 0..30 are its state, non-candidates are 0. */
int gpc_hash(gpc_ctx* ctx, const uint8_t* img, int w, int h, int gradient_threshold,
             uint32_t* states, int32_t* mask, int cap, int* n, uint32_t* hash_image);
/* Forest::findCorrespondences + rectifiedMatch filter (inference.hpp:227-254, :384-391) on
 * caller-provided hash images (same encoding as gpc_hash's hash_image). */
int gpc_match_hash_images(gpc_ctx* ctx, const uint32_t* hash_l, const uint32_t* hash_r, int w, int h,
                          const gpc_settings* s, gpc_support* out, int cap, int* n_out);

/* Forest::evalFastMaskOnSubsetSSE (inference.hpp:266-292) / ndb::gpcFilter[Tau] (filter.hpp:547,
 * :619) on a caller-provided SMOOTHED image: states[i] = state of pixel idx[i] (= y*w + x).
 * Entries outside the 13-pixel border (where the reference reads across row ends) and in the
 * never-hashed rows >= h-15 (filter.hpp:601-604) give 0. */
int gpc_hash_smooth(gpc_ctx* ctx, const uint8_t* smooth, int w, int h, const int32_t* idx, int n,
                    uint32_t* states);

/* ---- resident images: the device side of Forest::PreprocessedImage (inference.hpp:157-166) --
 * gpc_image_upload copies a raw image to the context's device once; gpc_image_preprocess is
 * preprocessImage (inference.hpp:302-333) on it, gpc_match_images is rectifiedMatch
 * (inference.hpp:375-393) on two resident images without another host->device copy. */
typedef struct gpc_image gpc_image;
int gpc_image_upload(gpc_ctx* ctx, const uint8_t* img, int w, int h, int stride, gpc_image** out);
void gpc_image_release(gpc_image* image);
/* smooth, grad, mask and n_mask are optional (NULL: not computed / not copied).  The image keeps what the kernels
 * produced on the device -- the smoothed image and candidates for this threshold and, if the context has a forest,
 * the hash image -- so that gpc_match_images / gpc_correspond_images with the same threshold and forest run the
 * matcher only (a changed threshold, result mode or forest is detected and recomputed). */
int gpc_image_preprocess(gpc_ctx* ctx, gpc_image* image, int gradient_threshold, uint8_t* smooth,
                         uint8_t* grad, int32_t* mask, int mask_cap, int* n_mask);
/* the candidate list of the most recent gpc_image_preprocess / gpc_preprocess call of this context, in the context's
 * page-locked staging buffer: n entries (as reported through n_mask); valid until the next call on the context */
const int32_t* gpc_mask_view(gpc_ctx* ctx, int* n);
/* preprocessImage's smoothed / gradient images of a resident image, on demand (either may be NULL) */
int gpc_image_fetch(gpc_ctx* ctx, const gpc_image* image, int gradient_threshold, uint8_t* smooth, uint8_t* grad);
int gpc_match_images(gpc_ctx* ctx, gpc_image* left, gpc_image* right, const gpc_settings* s,
                     gpc_support* out, int cap, int* n_out, int* n_cand_l, int* n_cand_r);
/* The first n supports of the most recent gpc_match_pair / gpc_match_images result, which stays on the device: a
 * caller that cannot bound the count calls the matcher with cap = 0 (GPC_E_CAPACITY, *n_out = count) and then this. */
int gpc_fetch_supports(gpc_ctx* ctx, gpc_support* out, int n);

/* == ndb::Correspondence (buffer.hpp:94-97): source point, target point */
typedef struct { int32_t xs, ys, xt, yt; } gpc_correspondence;
/* Forest::stereoMatch / depthPriorFast (inference.hpp:344-361, :184-226): every unique-unique
 * correspondence before rectifiedMatch's filter, ascending key order (both matching modes). */
int gpc_correspond_images(gpc_ctx* ctx, gpc_image* left, gpc_image* right, const gpc_settings* s,
                          gpc_correspondence* out, int cap, int* n_out);
/* Forest::findCorrespondences (inference.hpp:227-254) on explicit 64-bit keys (Descriptor::state):
 * out_pairs[2i], out_pairs[2i+1] = indices into src_keys / tar_keys of match i, ascending src key;
 * among equal keys the input order is kept (stable).  n_tar == 0 (undefined behaviour in the
 * reference) yields no matches. */
int gpc_find_correspondences(gpc_ctx* ctx, const uint64_t* src_keys, int n_src, const uint64_t* tar_keys,
                             int n_tar, int32_t* out_pairs, int cap, int* n_out);
/* ndb::Hashmatch<Descriptor> (hashmatch.hpp:208-272) driven as depthPriorFast does with useHashtable
 * (inference.hpp:204-225) on explicit 64-bit keys: 214673 buckets (key % 214673), all src keys inserted
 * first, then all tar keys; a bucket keeps the first 10 elements offered to it in stable ascending key
 * order and is walked by getDuplicates' rules (hashmatch.hpp:162-198).  out_pairs as above, in the
 * reference's output order: bucket index, then list order. */
int gpc_hashmatch(gpc_ctx* ctx, const uint64_t* src_keys, int n_src, const uint64_t* tar_keys, int n_tar,
                  int32_t* out_pairs, int cap, int* n_out);

/* Which build of the reference the results reproduce.  GPC_RESULTS_SSE (default): the reference compiled with
 * -D_INTRINSICS_SSE, its default (samples/CMakeLists.txt:13-17) and the build SURVEY.md 8a describes.
 * GPC_RESULTS_NAIVE: its SSE=OFF build, whose box, sobel and gpcFilter[Tau] forward to boxNaive, sobelNaive and
 * gpcFilter[Tau]Naive (filter.hpp:295-296, :406-407, :550-551, :622-623; bodies at :157-293): S = sum3x3 / 9,
 * signed Sobel responses divided towards zero, `a > b - tau` in int, first test in the highest state bit, every
 * candidate row hashed.  Different smooth / grad images, candidates, states and supports; matching is shared.
 * Forests of 32 tests are refused in this mode (GPC_E_UNSUPPORTED: bit 31 of a hash word is the candidate flag).
 * A forest set earlier is re-baked.  The fern hashing kernel runs un-specialised (no NVRTC build) in this mode. */
#define GPC_RESULTS_SSE 0
#define GPC_RESULTS_NAIVE 1
int gpc_set_result_mode(gpc_ctx* ctx, int mode);

/* Matcher selection.  AUTO: epipolar mode uses the per-row shared-memory matchers (a fast kernel for
 * ordinary rows, a general one for the rows it hands over), global mode the device-wide radix sort +
 * segmented scan.  SORT forces the radix-sort matcher for both, ROWS_GENERAL sends every row of epipolar
 * mode through the general row kernel (same results; used to cross-check the implementations). */
#define GPC_MATCHER_AUTO 0
#define GPC_MATCHER_SORT 1
#define GPC_MATCHER_ROWS_GENERAL 2
int gpc_set_matcher(gpc_ctx* ctx, int matcher);

/* After gpc_set_forest the hashing kernel is rebuilt with the forest baked into its code (NVRTC,
 * loaded with dlopen; about 1 s).  Returns "specialised", or "generic: <reason>" when the
 * precompiled kernel is in use (NVRTC absent, build failure, GPC_JIT=0).  Results are identical. */
const char* gpc_jit_status(const gpc_ctx* ctx);

/* ---- forests of more than 32 tests ("extended mode") and 32-test forests in GPC_RESULTS_NAIVE --------------------
 * The reference keeps the first 32 tests of a forest file and drops the rest (inference.hpp:426-432); config 5 of
 * BASELINE.json asks for all tests of a 16 x 12 forest.  There is no reference result for that --
 * PARITY UNPINNED: validated against this repository's scalar restatement (tests/test_wide.py) only.  Definition: the
 * state of a candidate is the tuple of words (word k = tests 32k .. 32k+31 hashed as a forest of their own, so word 0
 * is the reference's truncated state); two candidates match when the whole tuples are equal and unique on both sides;
 * findCorrespondences' tail rules apply to the tuple order; supports come in ascending (y, tuple) order, the tuple
 * compared from its last word down.
 * In GPC_RESULTS_NAIVE the same machinery carries the T-bit state of gpcFilter[Tau]Naive (filter.hpp:245-293) in
 * 31-bit words, which is what lets a 32-test forest through in that mode; there the reference defines the result
 * (pinned against the SSE=OFF build, tests/golden/naive32.json).
 * tests: n_tests rows of {ix, iy, jx, jy, tau}, file order, 1 <= n_tests <= GPC_MAX_WIDE_TESTS.  Set the result mode
 * first: gpc_set_result_mode drops the wide forest. */
#define GPC_MAX_WIDE_TESTS 256
int gpc_read_forest_tests(const char* path, int32_t* tests, int cap, int* n_tests, int* n_ferns);
int gpc_set_wide_forest(gpc_ctx* ctx, const int32_t* tests, int n_tests);
int gpc_match_pair_wide(gpc_ctx* ctx, const uint8_t* left, const uint8_t* right, int w, int h, int stride,
                        const gpc_settings* s, gpc_support* out, int cap, int* n_out, int* n_cand_l, int* n_cand_r);
/* stage seam: words[k][y][x] = candidate flag | word k of the pixel's state (0 for non-candidates) */
int gpc_hash_wide(gpc_ctx* ctx, const uint8_t* img, int w, int h, int gradient_threshold, uint32_t* words,
                  int n_words_cap, int* n_words);

/* ---- gpc_pool: one resident context + one host thread per GPU (multi-GPU driver) ----------
 * The reference has no multi-device path; stereo pairs are independent units, so a batch is cut into chunks that are
 * dealt round-robin to the devices.  Every device pipelines its chunks (upload / kernels / download) and the supports
 * land in pair order directly in the caller's buffer -- no device-to-device traffic, no collective, no second host copy.
 * gpc_pool_match_batch has gpc_match_batch's arguments and result.  `devices` may name a device more than once (several
 * contexts on one GPU).  max_batch_per_device bounds the pairs one device receives per call (ceil(n_pairs / n) + a
 * chunk).  A pool is driven from one host thread at a time. */
typedef struct gpc_pool gpc_pool;
int gpc_pool_create(gpc_pool** out, const int* devices, int n_devices, int max_w, int max_h, int max_batch_per_device);
void gpc_pool_destroy(gpc_pool* pool);
int gpc_pool_size(const gpc_pool* pool);
gpc_ctx* gpc_pool_context(gpc_pool* pool, int i);            /* the i-th device's context (settings such as gpc_set_matcher) */
const char* gpc_pool_last_error(const gpc_pool* pool);
int64_t gpc_pool_launch_count(const gpc_pool* pool);
int gpc_pool_set_forest(gpc_pool* pool, const gpc_forest* forest);
int gpc_pool_set_result_mode(gpc_pool* pool, int mode);
int gpc_pool_match_batch(gpc_pool* pool, const uint8_t* images, int n_pairs, int w, int h, const gpc_settings* s,
                         gpc_support* out, int64_t cap, int64_t* offsets, int32_t* n_cand);
/* page-locked host memory usable by every device (cudaHostAllocPortable) for `images` / `out`; NULL on failure */
void* gpc_host_alloc(size_t bytes);
void gpc_host_free(void* p);

/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t gpc_launch_count(const gpc_ctx* ctx);
/* process-unique serial of a context (never reused, unlike its address) */
int64_t gpc_context_id(const gpc_ctx* ctx);

/* Per-kernel device times, measured with CUDA events on the launching stream around each
 * kernel of the whole-path entry points (bench.py's roofline figures).  Slots:
 *   0 smooth_sobel (kernel A1)     1 hash_tiles (kernel A2)      2 match_rows (kernel B)
 *   3 row/pair scans               4 emit_supports (kernel C)
 * gpc_kernel_times synchronises the stream and returns accumulated milliseconds per slot and
 * the number of batch runs they cover. */
#define GPC_N_KERNELS 5
int gpc_enable_kernel_timing(gpc_ctx* ctx, int on);
int gpc_kernel_times(gpc_ctx* ctx, double* ms, int64_t* runs);

#ifdef __cplusplus
}
#endif
#endif

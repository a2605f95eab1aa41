/*
 * gpc_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, intrinsic-free restatement of the openGPC `sparsematch` inference path
 * (reference: lib/gpc/{filter,inference,buffer}.hpp, SSE build).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (opengpc_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  Every stage is checked bit-for-bit against the unmodified
 * reference headers compiled into oracle/_ref/libgpc_ref.so (oracle/Makefile,
 * oracle/ref_harness.cpp) and against the golden digests of SURVEY.md section 8c,
 * see tests/test_oracle_vs_reference.py and tests/golden/.
 *
 * Canonical conventions where the reference reads uninitialised memory:
 *   - smooth rows the SSE box never writes and clearBoundary never clears (row H-3 for
 *     even H) are 0;  grad rows 0 and H-3..H-1 are 0;
 *   - the one-byte reads at column -1 / W (previous / next row in linear memory) are 0.
 *     They only reach smooth columns 0, W-1 (cleared anyway) and grad columns 0, 1
 *     (outside the 13-pixel candidate border).
 */
#ifndef GPC_ORACLE_H
#define GPC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t x, y; float d; } gpco_support;          /* == ndb::Support, buffer.hpp:86-92 */
typedef struct { int32_t xs, ys, xt, yt; } gpco_correspondence;  /* == ndb::Correspondence, buffer.hpp:94-97 */

typedef struct {
  int32_t n_tests;          /* <= 32 (inference.hpp:426) */
  int32_t type;             /* 0 zero forest, 1 tau forest (inference.hpp:436-445) */
  int32_t n_discarded;      /* tests beyond the 32-test cap */
  int32_t ix[32], iy[32], jx[32], jy[32];
  int32_t tau[32];
} gpco_forest;

typedef struct {
  int32_t gradient_threshold;   /* uint8 in the reference (inference.hpp:74) */
  int32_t disp_high;            /* :76 */
  int32_t vertical_tolerance;   /* :78 */
  int32_t epipolar_mode;        /* :80 */
} gpco_settings;

/* SURVEY.md appendix C synthetic stereo pair (std::mt19937 raw outputs only). */
void gpco_synth(uint8_t* L, uint8_t* R, int w, int h, uint32_t seed);
/* digest of an ordered support list, SURVEY.md 8c */
uint64_t gpco_digest(const gpco_support* s, int n);

/* filter.hpp:293-392 + buffer.hpp:630-654 */
void gpco_box(const uint8_t* in, uint8_t* smooth, int w, int h);
/* filter.hpp:404-519 (with the unpacklo lane duplication of :504-507) */
void gpco_sobel(const uint8_t* in, uint8_t* grad, int w, int h, int thr);
/* filter.hpp:60-87 + inference.hpp:318-330; returns the number of candidates */
int gpco_candidates(const uint8_t* grad, int w, int h, int32_t* mask);
/* filter.hpp:547-606 / :619-683 + inference.hpp:266-292; one state per mask entry */
void gpco_hash(const uint8_t* smooth, int w, int h, const gpco_forest* f,
               const int32_t* mask, int n, uint32_t* states);
/* inference.hpp:227-254 on bare keys; out = (index into src, index into tar) pairs,
 * indices refer to the ORIGINAL (unsorted) order.  Equal keys keep input order
 * (stable), which is what std::sort yields for the <=16-element KATs and is the
 * documented convention for the tail-dup-2 case.  Returns the number of pairs. */
int gpco_find_correspondences(const uint64_t* src, int ns, const uint64_t* tar, int nt,
                              int32_t* out_pairs);
/* inference.hpp:184-202 + :227-254 + :375-393 */
int gpco_match(const int32_t* mask_l, const uint32_t* st_l, int nl,
               const int32_t* mask_r, const uint32_t* st_r, int nr,
               int w, const gpco_settings* s,
               gpco_correspondence* corr, int* n_corr,   /* may be NULL */
               gpco_support* supp);                       /* capacity >= min(nl,nr) */
/* useHashtable(true): ndb::Hashmatch (hashmatch.hpp:48-272) driven as in inference.hpp:204-225, on bare keys;
 * out = (index into src, index into tar) pairs in the reference's output order (bucket index, then list order) */
int gpco_hashmatch(const uint64_t* src, int ns, const uint64_t* tar, int nt, int32_t* out_pairs);
/* gpco_match with the hashtable matcher in place of findCorrespondences */
int gpco_match_hashtable(const int32_t* mask_l, const uint32_t* st_l, int nl,
                         const int32_t* mask_r, const uint32_t* st_r, int nr,
                         int w, const gpco_settings* s,
                         gpco_correspondence* corr, int* n_corr, gpco_support* supp);
/* inference.hpp:404-446; returns 0 ok, -1 cannot open */
int gpco_read_forest(const char* path, gpco_forest* f);

/* whole pair: preprocess x2 + rectifiedMatch (sparsematch.cpp:46-51).
 * Any of the optional outputs may be NULL.  Returns number of supports. */
int gpco_pair(const uint8_t* L, const uint8_t* R, int w, int h, const gpco_forest* f,
              const gpco_settings* s, gpco_support* supp, int* n_cand_l, int* n_cand_r);

/* The reference's SSE=OFF build: boxNaive (filter.hpp:207-231) + clearBoundary, sobelNaive (:157-187),
 * gpcFilterNaive / gpcFilterTauNaive (:245-262, :275-293); candidates and matching are shared with the SSE build. */
void gpco_box_naive(const uint8_t* in, uint8_t* smooth, int w, int h);
void gpco_sobel_naive(const uint8_t* in, uint8_t* grad, int w, int h, int thr);
void gpco_hash_naive(const uint8_t* smooth, int w, int h, const gpco_forest* f,
                     const int32_t* mask, int n, uint32_t* states);
int gpco_pair_naive(const uint8_t* L, const uint8_t* R, int w, int h, const gpco_forest* f,
                    const gpco_settings* s, int use_hashtable, gpco_support* supp, int* n_cand_l, int* n_cand_r);
/* gpco_pair with InferenceSettings::useHashtable(true) */
int gpco_pair_hashtable(const uint8_t* L, const uint8_t* R, int w, int h, const gpco_forest* f,
                        const gpco_settings* s, gpco_support* supp, int* n_cand_l, int* n_cand_r);

#ifdef __cplusplus
}
#endif
#endif

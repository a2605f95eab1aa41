/*
 * gpc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see gpc_oracle.h).
 *
 * Scalar restatement of the reference's SSE inference path.  Each function cites the
 * reference lines it follows (paths relative to the openGPC tree).  It deliberately
 * follows the SSE code, not the *Naive fall-backs (filter.hpp:157-282), which compute
 * a different function.
 */
#include "gpc_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * std::mt19937 (32-bit Mersenne twister, init_genrand seeding) -- input generator only.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int idx; } mt_t;

static void mt_seed(mt_t* g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; i++)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

static uint32_t mt_next(mt_t* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      g->mt[i] = v;
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* SURVEY.md appendix C */
void gpco_synth(uint8_t* L, uint8_t* R, int w, int h, uint32_t seed) {
  mt_t g;
  mt_seed(&g, seed);
  size_t cw = (size_t)(2 * w / 4 + 2), ch = (size_t)(h / 4 + 2);
  uint8_t* coarse = (uint8_t*)malloc(cw * ch);
  uint8_t* tex = (uint8_t*)malloc((size_t)w * h * 2);
  for (size_t i = 0; i < cw * ch; i++) coarse[i] = (uint8_t)(mt_next(&g) & 255u);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < 2 * w; x++) {
      int v = coarse[(size_t)(y / 4) * cw + (size_t)(x / 4)];
      tex[(size_t)y * 2 * w + x] = (uint8_t)((v * 3 + (int)(mt_next(&g) & 255u)) / 4);
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int d = 5 + (y * 40) / h;
      L[(size_t)y * w + x] = tex[(size_t)y * 2 * w + x + 200];
      int n = (int)(mt_next(&g) % 3u) - 1;
      int v = tex[(size_t)y * 2 * w + x + d + 200] + n;
      R[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  free(coarse);
  free(tex);
}

uint64_t gpco_digest(const gpco_support* s, int n) {
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < n; i++) {
    int32_t v[3] = { s[i].x, s[i].y, (int32_t)s[i].d };
    for (int k = 0; k < 3; k++) { h ^= (uint32_t)v[k]; h *= 1099511628211ull; }
  }
  return h;
}

/* ------------------------------------------------------------------------------------------
 * Row B: ndb::box (filter.hpp:293-392) + Buffer::clearBoundary (buffer.hpp:630-654)
 * ---------------------------------------------------------------------------------------- */
static inline int px(const uint8_t* in, int w, int h, int y, int x) {
  /* column -1 / w is the neighbouring row's byte in the reference's linear memory
   * (filter.hpp:325-326); canonical value 0, see header. */
  if (x < 0 || x >= w || y < 0 || y >= h) return 0;
  return in[(size_t)y * w + x];
}

static inline int third(int s) { return (s * 21846) >> 16; }  /* _mm_mulhi_epi16(.,21846), :304,:332 */
static inline int ninth(int s) { return (s * 7282) >> 16; }   /* _mm_mulhi_epi16(.,7282),  :416,:466 */

void gpco_box(const uint8_t* in, uint8_t* smooth, int w, int h) {
  memset(smooth, 0, (size_t)w * h);
  /* boxFilterSegment(1, height-3), two rows per iteration (:307, :388) */
  for (int y0 = 1; y0 < h - 3; y0 += 2)
    for (int y = y0; y <= y0 + 1; y++)
      for (int x = 0; x < w; x++) {
        int hs[3];
        for (int r = -1; r <= 1; r++)
          hs[r + 1] = third(px(in, w, h, y + r, x - 1) + px(in, w, h, y + r, x) + px(in, w, h, y + r, x + 1));
        smooth[(size_t)y * w + x] = (uint8_t)third(hs[0] + hs[1] + hs[2]);   /* :355-358 */
      }
  /* clearBoundary: columns 0,1 and w-1, row 0, rows h-2,h-1 (buffer.hpp:638-652) */
  for (int y = 0; y < h; y++) {
    smooth[(size_t)y * w + 0] = 0;
    if (w > 1) smooth[(size_t)y * w + 1] = 0;
    smooth[(size_t)y * w + w - 1] = 0;
  }
  for (int x = 0; x < w; x++) {
    smooth[x] = 0;
    if (h >= 2) smooth[(size_t)(h - 2) * w + x] = 0;
    if (h >= 1) smooth[(size_t)(h - 1) * w + x] = 0;
  }
}

/* ------------------------------------------------------------------------------------------
 * Row S: ndb::sobel (filter.hpp:404-519)
 * ---------------------------------------------------------------------------------------- */
static int sobel_true(const uint8_t* in, int w, int h, int y, int x, int16_t thr2) {
  int p00 = px(in, w, h, y - 1, x - 1), p01 = px(in, w, h, y - 1, x), p02 = px(in, w, h, y - 1, x + 1);
  int p10 = px(in, w, h, y, x - 1), p12 = px(in, w, h, y, x + 1);
  int p20 = px(in, w, h, y + 1, x - 1), p21 = px(in, w, h, y + 1, x), p22 = px(in, w, h, y + 1, x + 1);
  int a = ninth(p00 + p20 + p10 + p10);   /* :466 */
  int b = ninth(p02 + p22 + p12 + p12);   /* :470 */
  int c = ninth(p00 + p02 + p01 + p01);   /* :482 */
  int d = ninth(p20 + p22 + p21 + p21);   /* :486 */
  int16_t sx = (int16_t)((a - b) * (a - b));   /* _mm_mullo_epi16 :477 */
  int16_t sy = (int16_t)((c - d) * (c - d));   /* :494 */
  int sum = (int)sx + (int)sy;                 /* _mm_adds_epi16 :505 */
  if (sum > 32767) sum = 32767;
  if (sum < -32768) sum = -32768;
  return sum > (int)thr2;                      /* _mm_cmpgt_epi16 vs set1_epi16(thr*thr) :418 */
}

void gpco_sobel(const uint8_t* in, uint8_t* grad, int w, int h, int thr) {
  memset(grad, 0, (size_t)w * h);
  int16_t thr2 = (int16_t)((thr & 255) * (thr & 255));   /* wraps negative for thr >= 182 */
  for (int y = 1; y < h - 3; y++)                          /* sobelSSESegment(1, height-3) :517 */
    for (int s = 0; s < w; s += 16)
      for (int j = 0; j < 16; j++) {
        /* unpacklo_epi8 of a 16-bit compare mask duplicates lanes 0..3 (:504-507) */
        int src = (j < 8) ? s + j / 2 : s + 8 + (j - 8) / 2;
        grad[(size_t)y * w + s + j] = sobel_true(in, w, h, y, src, thr2) ? 255 : 0;
      }
}

/* ------------------------------------------------------------------------------------------
 * Row C: ndb::arr2ind (filter.hpp:60-87) + border lambda (inference.hpp:318-330)
 * ---------------------------------------------------------------------------------------- */
int gpco_candidates(const uint8_t* grad, int w, int h, int32_t* mask) {
  int n = 0;
  for (int y = 13; y < h - 13; y++)
    for (int x = 13; x < w - 13; x++)
      if (grad[(size_t)y * w + x]) mask[n++] = y * w + x;
  return n;
}

/* ------------------------------------------------------------------------------------------
 * Row H: ndb::gpcFilter / gpcFilterTau (filter.hpp:547-606, :619-683)
 * ---------------------------------------------------------------------------------------- */
static inline int sat8(int v) { return v < -128 ? -128 : (v > 127 ? 127 : v); }

void gpco_hash(const uint8_t* smooth, int w, int h, const gpco_forest* f,
               const int32_t* mask, int n, uint32_t* states) {
  for (int i = 0; i < n; i++) {
    int k = mask[i], x = k % w, y = k / w;
    uint32_t st = 0;
    /* only rows 13 <= y < h-15 are hashed (:601-604); others keep the zero fill of
     * inference.hpp:274 */
    if (y >= 13 && y < h - 15) {
      for (int t = 0; t < f->n_tests && t < 32; t++) {
        int a = smooth[(size_t)(y + f->iy[t]) * w + x + f->ix[t]];
        int b = smooth[(size_t)(y + f->jy[t]) * w + x + f->jx[t]];
        if (f->type == 1)   /* _mm_subs_epi8(b, set1_epi8(tau)) :649-651, then unsigned compare */
          b = sat8((int)(int8_t)b - (int)(int8_t)f->tau[t]) & 255;
        uint32_t bit = (uint32_t)(a > b);
        /* bit placement of the bitMask/k bookkeeping at :574-584 */
        if (t < 8) st |= bit << t;
        else if (t == 8) { if ((x & 7) != 0) st |= bit; }   /* 2x int64 add drops byte lanes 0 and 8 */
        else st |= bit << (t - 1);
      }
    }
    states[i] = st;
  }
}

/* ------------------------------------------------------------------------------------------
 * The reference's SSE=OFF build (samples/CMakeLists.txt:13-17 option SSE): box, sobel, gpcFilter and
 * gpcFilterTau forward to their *Naive versions (filter.hpp:295-296, :406-407, :550-551, :622-623),
 * which produce different images, candidates and states than the SSE code.  All three walk the image
 * as one linear array: "left" and "right" neighbours of the first / last column are the bytes of the
 * adjacent rows, and the last two outputs read two bytes past the image (taken as 0 here; both land
 * in rows / columns that are cleared or never become candidates).
 * ---------------------------------------------------------------------------------------- */
static inline int lin(const uint8_t* in, long long n, long long i) { return (i < 0 || i >= n) ? 0 : in[i]; }

/* boxNaive (filter.hpp:207-231) + clearBoundary: plain sum of the 3x3 neighbourhood / 9 */
void gpco_box_naive(const uint8_t* in, uint8_t* smooth, int w, int h) {
  long long n = (long long)w * h;
  memset(smooth, 0, (size_t)n);
  for (long long o = w + 1; o <= (long long)(h - 1) * w; o++) {          /* (height-2)*width outputs from blurred+width+1 */
    int sum = 0;
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) sum += lin(in, n, o + (long long)dy * w + dx);
    smooth[o] = (uint8_t)(sum / 9);
  }
  for (int y = 0; y < h; y++) {                                            /* clearBoundary, buffer.hpp:638-652 */
    smooth[(size_t)y * w + 0] = 0;
    if (w > 1) smooth[(size_t)y * w + 1] = 0;
    smooth[(size_t)y * w + w - 1] = 0;
  }
  for (int x = 0; x < w; x++) {
    smooth[x] = 0;
    if (h >= 2) smooth[(size_t)(h - 2) * w + x] = 0;
    if (h >= 1) smooth[(size_t)(h - 1) * w + x] = 0;
  }
}

/* sobelNaive (filter.hpp:157-187): signed Sobel responses, C division (towards zero), int compare.
 * Positions the reference never writes (row 0, (1,0), row h-1 from column 1) are 0 here. */
void gpco_sobel_naive(const uint8_t* in, uint8_t* grad, int w, int h, int thr) {
  long long n = (long long)w * h;
  int thr2 = (thr & 255) * (thr & 255);
  memset(grad, 0, (size_t)n);
  for (long long o = w + 1; o <= (long long)(h - 1) * w; o++) {
    int p11 = lin(in, n, o - w - 1), p12 = lin(in, n, o - w), p13 = lin(in, n, o - w + 1);
    int p21 = lin(in, n, o - 1), p23 = lin(in, n, o + 1);
    int p31 = lin(in, n, o + w - 1), p32 = lin(in, n, o + w), p33 = lin(in, n, o + w + 1);
    int sx = (p11 + p31 + 2 * p21 - p13 - 2 * p23 - p33) / 9;            /* :176 */
    int sy = (p11 + p13 + 2 * p12 - p31 - 2 * p32 - p33) / 9;            /* :177 */
    grad[o] = (sx * sx + sy * sy > thr2) ? 255 : 0;                        /* :179-181 */
  }
}

/* gpcFilterNaive / gpcFilterTauNaive (filter.hpp:245-262, :275-293): every candidate is hashed (no row
 * range), the first test ends up in the highest bit, the tau comparison is plain int arithmetic */
void gpco_hash_naive(const uint8_t* smooth, int w, int h, const gpco_forest* f,
                     const int32_t* mask, int n, uint32_t* states) {
  (void)h;
  for (int i = 0; i < n; i++) {
    int k = mask[i];
    uint32_t st = 0;
    for (int t = 0; t < f->n_tests && t < 32; t++) {
      int a = smooth[k + f->ix[t] + f->iy[t] * w];
      int b = smooth[k + f->jx[t] + f->jy[t] * w];
      st <<= 1;
      if (f->type == 1 ? (a > b - f->tau[t]) : (a > b)) st++;
    }
    states[i] = st;
  }
}

/* ------------------------------------------------------------------------------------------
 * Row M: Forest::findCorrespondences (inference.hpp:227-254)
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t key; int32_t idx; } kv_t;

/* stable LSD radix sort on the 64-bit key */
static void sort_kv(kv_t* a, int n) {
  if (n < 2) return;
  kv_t* b = (kv_t*)malloc((size_t)n * sizeof(kv_t));
  uint64_t all_or = 0, all_and = ~0ull;
  for (int i = 0; i < n; i++) { all_or |= a[i].key; all_and &= a[i].key; }
  for (int pass = 0; pass < 8; pass++) {
    int sh = pass * 8;
    if ((((all_or ^ all_and) >> sh) & 255u) == 0) continue;   /* digit constant over the input */
    size_t cnt[257] = {0};
    for (int i = 0; i < n; i++) cnt[((a[i].key >> sh) & 255u) + 1]++;
    for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
    for (int i = 0; i < n; i++) b[cnt[(a[i].key >> sh) & 255u]++] = a[i];
    memcpy(a, b, (size_t)n * sizeof(kv_t));
  }
  free(b);
}

static int scan_sorted(const kv_t* S, int ns, const kv_t* T, int nt, int32_t* out_pairs) {
  int m = 0;
  if (nt == 0) return 0;   /* reference: unsigned size()-1 underflow, UB; defined as no matches */
  uint32_t j = 0;
  uint32_t last = (uint32_t)nt - 1;
  for (uint32_t i = 0; i < (uint32_t)ns; ++i) {
    int unique = 1;
    while (i + 1 < (uint32_t)ns && S[i].key == S[i + 1].key) { ++i; unique = 0; }   /* :238-239 */
    if (unique) {
      for (; j < last; ++j)                                                          /* :243-246 */
        if (!(T[j].key < S[i].key)) break;
      if (j != last && T[j].key == S[i].key &&
          ((j + 1) == last || !(T[j].key == T[j + 1].key))) {                        /* :248-249 */
        out_pairs[2 * m] = S[i].idx;
        out_pairs[2 * m + 1] = T[j].idx;
        m++;
      }
    }
  }
  return m;
}

int gpco_find_correspondences(const uint64_t* src, int ns, const uint64_t* tar, int nt,
                              int32_t* out_pairs) {
  kv_t* S = (kv_t*)malloc((size_t)(ns > 0 ? ns : 1) * sizeof(kv_t));
  kv_t* T = (kv_t*)malloc((size_t)(nt > 0 ? nt : 1) * sizeof(kv_t));
  for (int i = 0; i < ns; i++) { S[i].key = src[i]; S[i].idx = i; }
  for (int i = 0; i < nt; i++) { T[i].key = tar[i]; T[i].idx = i; }
  sort_kv(S, ns);
  sort_kv(T, nt);
  int m = scan_sorted(S, ns, T, nt, out_pairs);
  free(S);
  free(T);
  return m;
}

/* Rows K, M, F: depthPriorFast (inference.hpp:184-202) + rectifiedMatch (:375-393) */
int gpco_match(const int32_t* mask_l, const uint32_t* st_l, int nl,
               const int32_t* mask_r, const uint32_t* st_r, int nr,
               int w, const gpco_settings* s,
               gpco_correspondence* corr, int* n_corr, gpco_support* supp) {
  uint64_t* kl = (uint64_t*)malloc((size_t)(nl > 0 ? nl : 1) * 8);
  uint64_t* kr = (uint64_t*)malloc((size_t)(nr > 0 ? nr : 1) * 8);
  for (int i = 0; i < nl; i++) {
    kl[i] = st_l[i];
    if (s->epipolar_mode) kl[i] |= (uint64_t)(uint32_t)(mask_l[i] / w) << 32;   /* :192-197 */
  }
  for (int i = 0; i < nr; i++) {
    kr[i] = st_r[i];
    if (s->epipolar_mode) kr[i] |= (uint64_t)(uint32_t)(mask_r[i] / w) << 32;
  }
  int cap = nl < nr ? nl : nr;
  int32_t* pairs = (int32_t*)malloc((size_t)(cap > 0 ? cap : 1) * 8);
  int m = gpco_find_correspondences(kl, nl, kr, nr, pairs);
  int ns = 0;
  for (int i = 0; i < m; i++) {
    int ks = mask_l[pairs[2 * i]], kt = mask_r[pairs[2 * i + 1]];
    int xs = ks % w, ys = ks / w, xt = kt % w, yt = kt / w;
    if (corr) { corr[i].xs = xs; corr[i].ys = ys; corr[i].xt = xt; corr[i].yt = yt; }
    if (abs(ys - yt) <= s->vertical_tolerance && abs(xs - xt) <= s->disp_high) {   /* :386-388 */
      supp[ns].x = xs; supp[ns].y = ys; supp[ns].d = (float)(xs - xt);              /* :389-390 */
      ns++;
    }
  }
  if (n_corr) *n_corr = m;
  free(kl); free(kr); free(pairs);
  return ns;
}

/* ------------------------------------------------------------------------------------------
 * useHashtable(true): ndb::Hashmatch (hashmatch.hpp:48-272) as driven by depthPriorFast
 * (inference.hpp:204-225).  214673 buckets, bucket = key % 214673 (buffer.hpp:84-86); all src
 * descriptors are inserted first, then all tar descriptors, each in list order.
 *
 * A bucket is a linked list kept ascending by key.  insert() (hashmatch.hpp:93-134) drops the
 * element when the bucket already holds 10, otherwise places it behind every element whose key
 * is <= its own: the bucket is the stable ascending order of the first 10 elements offered to it.
 * getDuplicates() (:162-198) walks the list once; restated on an array below.
 * ---------------------------------------------------------------------------------------- */
#define GPCO_HT_BUCKETS 214673   /* inference.hpp:212 */
#define GPCO_HT_DEPTH 10         /* hashmatch.hpp:95 terminateAfter */

typedef struct { uint64_t key; int32_t idx; int32_t is_src; } ht_item;

/* walk of one bucket (hashmatch.hpp:162-198); L = list in order, m = its length */
static int ht_walk(const ht_item* L, int m, int32_t* out_pairs, int n_out) {
  int j = 0;
  while (j < m) {
    int p = j;                         /* prev = next; next = next->next  (:167-168) */
    j = j + 1;
    if (j < m && L[p].key == L[j].key) {                       /* :171 */
      if (L[p].is_src != L[j].is_src) {                        /* :172 diffImgs */
        if (j + 1 < m) {                                       /* :174 a third element exists */
          if (L[j + 1].key != L[j].key) {                      /* :176 */
            out_pairs[2 * n_out] = L[p].idx; out_pairs[2 * n_out + 1] = L[j].idx; n_out++;
          }
          if (j + 2 >= m) return n_out;                        /* :179 the last triplet was just checked */
        } else {                                               /* :181 */
          out_pairs[2 * n_out] = L[p].idx; out_pairs[2 * n_out + 1] = L[j].idx; n_out++;
        }
      } else if (j + 1 < m && L[j].is_src != L[j + 1].is_src) { /* :189 skip over the false pair */
        j = j + 1;
      }
    }
  }
  return n_out;
}

int gpco_hashmatch(const uint64_t* src, int ns, const uint64_t* tar, int nt, int32_t* out_pairs) {
  int n = ns + nt;
  /* bucket contents in insertion order: counting sort of the insertion sequence by bucket */
  int32_t* start = (int32_t*)calloc((size_t)GPCO_HT_BUCKETS + 1, sizeof(int32_t));
  int32_t* order = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  for (int i = 0; i < n; i++) start[(i < ns ? src[i] : tar[i - ns]) % GPCO_HT_BUCKETS + 1]++;
  for (int b = 0; b < GPCO_HT_BUCKETS; b++) start[b + 1] += start[b];
  int32_t* fill = (int32_t*)malloc((size_t)GPCO_HT_BUCKETS * sizeof(int32_t));
  memcpy(fill, start, (size_t)GPCO_HT_BUCKETS * sizeof(int32_t));
  for (int i = 0; i < n; i++) order[fill[(i < ns ? src[i] : tar[i - ns]) % GPCO_HT_BUCKETS]++] = i;
  int n_out = 0;
  for (int b = 0; b < GPCO_HT_BUCKETS; b++) {                  /* hashmatch.hpp:254-259: buckets in index order */
    ht_item L[GPCO_HT_DEPTH];
    int m = 0;
    for (int q = start[b]; q < start[b + 1] && m < GPCO_HT_DEPTH; q++) {   /* :101 a full bucket drops the rest */
      int i = order[q];
      ht_item e;
      e.is_src = i < ns; e.idx = e.is_src ? i : i - ns; e.key = e.is_src ? src[i] : tar[i - ns];
      int pos = m;                                             /* :112-116 behind every element <= e */
      while (pos > 0 && L[pos - 1].key > e.key) { L[pos] = L[pos - 1]; pos--; }
      L[pos] = e;
      m++;
    }
    n_out = ht_walk(L, m, out_pairs, n_out);
  }
  free(start); free(order); free(fill);
  return n_out;
}

/* depthPriorFast with useHashtable_ (inference.hpp:204-225) + rectifiedMatch (:375-393).  A pair is
 * (first, second) in list order; src elements precede equal tar elements, so first is always src. */
int gpco_match_hashtable(const int32_t* mask_l, const uint32_t* st_l, int nl,
                         const int32_t* mask_r, const uint32_t* st_r, int nr,
                         int w, const gpco_settings* s,
                         gpco_correspondence* corr, int* n_corr, gpco_support* supp) {
  uint64_t* kl = (uint64_t*)malloc((size_t)(nl > 0 ? nl : 1) * 8);
  uint64_t* kr = (uint64_t*)malloc((size_t)(nr > 0 ? nr : 1) * 8);
  for (int i = 0; i < nl; i++) {
    kl[i] = st_l[i];
    if (s->epipolar_mode) kl[i] |= (uint64_t)(uint32_t)(mask_l[i] / w) << 32;
  }
  for (int i = 0; i < nr; i++) {
    kr[i] = st_r[i];
    if (s->epipolar_mode) kr[i] |= (uint64_t)(uint32_t)(mask_r[i] / w) << 32;
  }
  int cap = nl < nr ? nl : nr;
  int32_t* pairs = (int32_t*)malloc((size_t)(cap > 0 ? cap : 1) * 8);
  int m = gpco_hashmatch(kl, nl, kr, nr, pairs);
  int ns = 0;
  for (int i = 0; i < m; i++) {
    int ks = mask_l[pairs[2 * i]], kt = mask_r[pairs[2 * i + 1]];
    int xs = ks % w, ys = ks / w, xt = kt % w, yt = kt / w;
    if (corr) { corr[i].xs = xs; corr[i].ys = ys; corr[i].xt = xt; corr[i].yt = yt; }
    if (abs(ys - yt) <= s->vertical_tolerance && abs(xs - xt) <= s->disp_high) {
      supp[ns].x = xs; supp[ns].y = ys; supp[ns].d = (float)(xs - xt);
      ns++;
    }
  }
  if (n_corr) *n_corr = m;
  free(kl); free(kr); free(pairs);
  return ns;
}

/* ------------------------------------------------------------------------------------------
 * Row R: Forest::readForest (inference.hpp:404-446)
 * ---------------------------------------------------------------------------------------- */
int gpco_read_forest(const char* path, gpco_forest* f) {
  memset(f, 0, sizeof(*f));
  FILE* fp = fopen(path, "r");
  if (!fp) return -1;   /* :409-412: empty mask, type 0 */
  int num_ferns = 0, nonzero = 0;
  if (fscanf(fp, "%d", &num_ferns) != 1) num_ferns = 0;
  for (int i = 0; i < num_ferns; i++) {
    int id, nt;
    char scale[64];
    if (fscanf(fp, "%d %63s %d", &id, scale, &nt) != 3) break;
    for (int j = 0; j < nt; j++) {
      int lvl, ix, iy, jx, jy, tau;
      if (fscanf(fp, "%d %d %d %d %d %d", &lvl, &ix, &iy, &jx, &jy, &tau) != 6) { nt = 0; break; }
      if (f->n_tests < 32) {                    /* :426 */
        int t = f->n_tests++;
        f->ix[t] = ix; f->iy[t] = iy; f->jx[t] = jx; f->jy[t] = jy; f->tau[t] = tau;
      } else {
        f->n_discarded++;
      }
      if (tau != 0) nonzero++;                  /* :433, counted for discarded tests too */
    }
  }
  fclose(fp);
  f->type = nonzero ? 1 : 0;
  return 0;
}

/* sparsematch.cpp:46-51 */
static int pair_impl(const uint8_t* L, const uint8_t* R, int w, int h, const gpco_forest* f,
                     const gpco_settings* s, gpco_support* supp, int* n_cand_l, int* n_cand_r, int use_hashtable, int naive) {
  size_t P = (size_t)w * h;
  uint8_t* sm[2]; uint8_t* gr[2]; int32_t* mk[2]; uint32_t* st[2]; int n[2];
  const uint8_t* img[2] = { L, R };
  for (int k = 0; k < 2; k++) {
    sm[k] = (uint8_t*)malloc(P); gr[k] = (uint8_t*)malloc(P);
    mk[k] = (int32_t*)malloc(P * 4 + 4);
    if (naive) { gpco_box_naive(img[k], sm[k], w, h); gpco_sobel_naive(img[k], gr[k], w, h, s->gradient_threshold); }
    else { gpco_box(img[k], sm[k], w, h); gpco_sobel(img[k], gr[k], w, h, s->gradient_threshold); }
    n[k] = gpco_candidates(gr[k], w, h, mk[k]);
    st[k] = (uint32_t*)malloc((size_t)(n[k] > 0 ? n[k] : 1) * 4);
    if (naive) gpco_hash_naive(sm[k], w, h, f, mk[k], n[k], st[k]);
    else gpco_hash(sm[k], w, h, f, mk[k], n[k], st[k]);
  }
  int ns = use_hashtable ? gpco_match_hashtable(mk[0], st[0], n[0], mk[1], st[1], n[1], w, s, NULL, NULL, supp)
                         : gpco_match(mk[0], st[0], n[0], mk[1], st[1], n[1], w, s, NULL, NULL, supp);
  if (n_cand_l) *n_cand_l = n[0];
  if (n_cand_r) *n_cand_r = n[1];
  for (int k = 0; k < 2; k++) { free(sm[k]); free(gr[k]); free(mk[k]); free(st[k]); }
  return ns;
}

int gpco_pair(const uint8_t* L, const uint8_t* R, int w, int h, const gpco_forest* f,
              const gpco_settings* s, gpco_support* supp, int* n_cand_l, int* n_cand_r) {
  return pair_impl(L, R, w, h, f, s, supp, n_cand_l, n_cand_r, 0, 0);
}

int gpco_pair_hashtable(const uint8_t* L, const uint8_t* R, int w, int h, const gpco_forest* f,
                        const gpco_settings* s, gpco_support* supp, int* n_cand_l, int* n_cand_r) {
  return pair_impl(L, R, w, h, f, s, supp, n_cand_l, n_cand_r, 1, 0);
}

/* the whole pair in the reference's SSE=OFF result mode (use_hashtable as in InferenceSettings) */
int gpco_pair_naive(const uint8_t* L, const uint8_t* R, int w, int h, const gpco_forest* f,
                    const gpco_settings* s, int use_hashtable, gpco_support* supp, int* n_cand_l, int* n_cand_r) {
  return pair_impl(L, R, w, h, f, s, supp, n_cand_l, n_cand_r, use_hashtable, 1);
}

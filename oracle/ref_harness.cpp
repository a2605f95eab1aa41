// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C-callable wrapper around the UNMODIFIED reference headers, compiled where they lie
// (-I/root/reference/lib) into oracle/_ref/libgpc_ref.so by oracle/Makefile with the flags
// of samples/CMakeLists.txt:15-18.  Nothing from the reference is copied into this repo;
// Eigen/libpng are replaced by the declaration-level shims in oracle/ref_shim/.
//
// Used (a) to pin the C restatement in gpc_oracle.c, (b) as the CPU baseline
// ("kind": "reference") of bench.py.  The only intervention is the determinism fix of
// the rows of `smooth` that the reference never writes are zero (the shim
// zero-fills, and row H-3 is re-zeroed explicitly below).
#include <sys/stat.h>
#include <functional>
#include <iostream>
#include <sstream>
#include <thread>
#include <vector>
#include <cstring>
#include <cstdint>
#include <chrono>
#include "gpc/inference.hpp"

typedef gpc::inference::Forest Forest;
typedef gpc::inference::InferenceSettings Settings;

namespace {

struct Quiet {   // readForest prints to cout (inference.hpp:417,431)
  std::streambuf* old;
  std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

ndb::Buffer<uint8_t> to_buffer(const uint8_t* img, int w, int h) {
  ndb::Buffer<uint8_t> b(h, w);
  std::memcpy(b.data(), img, (size_t)w * h);
  return b;
}

Settings make_settings(int thr, int disp_high, int vt, int epipolar, int num_threads) {
  return Settings().builder().gradientThreshold((uint8_t)thr).verticalTolerance(vt).dispHigh(disp_high)
      .epipolarMode(epipolar != 0).useHashtable(false).numThreads(num_threads);
}

void zero_unwritten_row(Forest::PreprocessedImage& p) {
#ifdef _INTRINSICS_SSE
  int h = (int)p.smooth.rows(), w = (int)p.smooth.cols();
  if (h >= 3 && (h % 2) == 0) std::memset(p.smooth.data() + (size_t)(h - 3) * w, 0, (size_t)w);
#else
  (void)p;   // the SSE=OFF build (boxNaive, filter.hpp:207-231) writes every row that clearBoundary leaves
#endif
}

double ms_between(gpc::inference::time_point a, gpc::inference::time_point b) {
  return 1000.0 * std::chrono::duration_cast<std::chrono::duration<double>>(b - a).count();
}

}  // namespace

extern "C" {

struct ref_support { int32_t x, y; float d; };

// 1: compiled with -D_INTRINSICS_SSE (the reference's default build), 0: the SSE=OFF build (*Naive functions)
int ref_is_sse_build() {
#ifdef _INTRINSICS_SSE
  return 1;
#else
  return 0;
#endif
}

int ref_sizeof_descriptor() { return (int)sizeof(ndb::Descriptor); }
int ref_sizeof_support() { return (int)sizeof(ndb::Support); }
int ref_sizeof_correspondence() { return (int)sizeof(ndb::Correspondence); }

// Forest::readForest (inference.hpp:404).  offsets: up to 64 ints, tau: up to 32 ints.
int ref_read_forest(const char* path, int w, int h, int32_t* offsets, int32_t* tau, int* type, int* n_tau) {
  Quiet q;
  Forest forest;
  Forest::FilterMask fm = forest.readForest(path, w, h);
  for (size_t i = 0; i < fm.mask.size(); i++) offsets[i] = fm.mask[i];
  for (size_t i = 0; i < fm.tau.size(); i++) tau[i] = fm.tau[i];
  *type = fm.type;
  *n_tau = (int)fm.tau.size();
  return (int)fm.mask.size() / 2;
}

// Forest::preprocessImage (inference.hpp:302).  Returns the candidate count.
int ref_preprocess(const uint8_t* img, int w, int h, int thr, uint8_t* smooth, uint8_t* grad, int32_t* mask) {
  Forest forest;
  ndb::Buffer<uint8_t> b = to_buffer(img, w, h);
  Forest::PreprocessedImage p = forest.preprocessImage(b, make_settings(thr, 128, 0, 1, 1));
  zero_unwritten_row(p);
  if (smooth) std::memcpy(smooth, p.smooth.data(), (size_t)w * h);
  if (grad) std::memcpy(grad, p.grad.data(), (size_t)w * h);
  if (mask) std::memcpy(mask, p.mask.data(), p.mask.size() * sizeof(int));
  return (int)p.mask.size();
}

// evalFastMaskOnSubsetSSE (inference.hpp:266) on an already preprocessed image.
int ref_hash(const uint8_t* img, int w, int h, int thr, const char* forest_path, int num_threads, uint32_t* states) {
  Quiet q;
  Forest forest;
  Settings s = make_settings(thr, 128, 0, 1, num_threads);
  ndb::Buffer<uint8_t> b = to_buffer(img, w, h);
  Forest::FilterMask fm = forest.readForest(forest_path, w, h);
  Forest::PreprocessedImage p = forest.preprocessImage(b, s);
  zero_unwritten_row(p);
  std::vector<ndb::Descriptor> d = forest.evalFastMaskOnSubsetSSE(p.smooth, p.grad, p.mask, fm, s);
  for (size_t i = 0; i < d.size(); i++) states[i] = (uint32_t)d[i].state;
  return (int)d.size();
}

// Forest::findCorrespondences (inference.hpp:227) on bare keys; point.x carries the input index.
int ref_find_correspondences(const uint64_t* src, int ns, const uint64_t* tar, int nt, int32_t* out_pairs) {
  Forest forest;
  std::vector<ndb::Descriptor> S(ns), T(nt);
  for (int i = 0; i < ns; i++) S[i] = ndb::Descriptor(ndb::Point(i, 0), src[i]);
  for (int i = 0; i < nt; i++) T[i] = ndb::Descriptor(ndb::Point(i, 0), tar[i]);
  if (nt == 0) return 0;   // reference underflows size()-1 here (UB)
  std::vector<ndb::Correspondence> c = forest.findCorrespondences(S, T);
  for (size_t i = 0; i < c.size(); i++) { out_pairs[2 * i] = c[i].srcPt.x; out_pairs[2 * i + 1] = c[i].tarPt.x; }
  return (int)c.size();
}

// ndb::Hashmatch driven exactly as depthPriorFast drives it (inference.hpp:204-225) on bare keys;
// point.x carries the input index.
int ref_hashmatch(const uint64_t* src, int ns, const uint64_t* tar, int nt, int32_t* out_pairs) {
  ndb::Hashmatch<ndb::Descriptor> hm(214673, ns + nt);
  for (int i = 0; i < ns; i++) { ndb::Descriptor d(ndb::Point(i, 0), src[i]); d.srcDescr = true; hm.insert(d); }
  for (int i = 0; i < nt; i++) { ndb::Descriptor d(ndb::Point(i, 0), tar[i]); d.srcDescr = false; hm.insert(d); }
  std::vector<std::pair<ndb::Descriptor, ndb::Descriptor>> corr;
  hm.getDuplicates(corr);
  for (size_t i = 0; i < corr.size(); i++) {
    out_pairs[2 * i] = corr[i].first.point.x;
    out_pairs[2 * i + 1] = corr[i].second.point.x;
  }
  return (int)corr.size();
}

// ref_pair with InferenceSettings::useHashtable(true) (inference.hpp:204-225).
int ref_pair_hashtable(const uint8_t* L, const uint8_t* R, int w, int h, const char* forest_path,
                       int thr, int disp_high, int vt, int epipolar, ref_support* supp, int cap) {
  Quiet q;
  Forest forest;
  Settings s = make_settings(thr, disp_high, vt, epipolar, 1).useHashtable(true);
  ndb::Buffer<uint8_t> simg = to_buffer(L, w, h), timg = to_buffer(R, w, h);
  Forest::FilterMask fm = forest.readForest(forest_path, (int)simg.cols(), (int)simg.rows());
  Forest::PreprocessedImage sp = forest.preprocessImage(simg, s);
  Forest::PreprocessedImage tp = forest.preprocessImage(timg, s);
  zero_unwritten_row(sp);
  zero_unwritten_row(tp);
  std::vector<ndb::Support> out = forest.rectifiedMatch(sp, tp, fm, s);
  if ((int)out.size() > cap) return -(int)out.size();
  for (size_t i = 0; i < out.size(); i++) { supp[i].x = out[i].x; supp[i].y = out[i].y; supp[i].d = out[i].d; }
  return (int)out.size();
}

// The sparsematch.cpp:42-51 window.  Returns the support count (or -required if cap is too small).
int ref_pair(const uint8_t* L, const uint8_t* R, int w, int h, const char* forest_path,
             int thr, int disp_high, int vt, int epipolar, int num_threads,
             ref_support* supp, int cap, int* n_cand_l, int* n_cand_r,
             double* ms_preprocess, double* ms_match) {
  Quiet q;
  Forest forest;
  Settings s = make_settings(thr, disp_high, vt, epipolar, num_threads);
  ndb::Buffer<uint8_t> simg = to_buffer(L, w, h), timg = to_buffer(R, w, h);
  Forest::FilterMask fm = forest.readForest(forest_path, (int)simg.cols(), (int)simg.rows());
  gpc::inference::time_point t0 = gpc::inference::sysTick();
  Forest::PreprocessedImage sp = forest.preprocessImage(simg, s);
  Forest::PreprocessedImage tp = forest.preprocessImage(timg, s);
  zero_unwritten_row(sp);
  zero_unwritten_row(tp);
  gpc::inference::time_point t1 = gpc::inference::sysTick();
  std::vector<ndb::Support> out = forest.rectifiedMatch(sp, tp, fm, s);
  gpc::inference::time_point t2 = gpc::inference::sysTick();
  if (n_cand_l) *n_cand_l = (int)sp.mask.size();
  if (n_cand_r) *n_cand_r = (int)tp.mask.size();
  if (ms_preprocess) *ms_preprocess = ms_between(t0, t1);
  if (ms_match) *ms_match = ms_between(t1, t2);
  if ((int)out.size() > cap) return -(int)out.size();
  for (size_t i = 0; i < out.size(); i++) { supp[i].x = out[i].x; supp[i].y = out[i].y; supp[i].d = out[i].d; }
  return (int)out.size();
}

// CPU throughput: `threads` host threads, each running `iters` times over its own pair
// (pair k = images[k % n_pairs]); times only the t0..t2 window of sparsematch.cpp:45-52.
// Returns wall seconds for the whole job; *total_supports sums the outputs (keeps the work live).
double ref_time_pairs(const uint8_t* images /* n_pairs x 2 x h x w */, int n_pairs, int w, int h,
                      const char* forest_path, int thr, int disp_high, int vt, int epipolar,
                      int threads, int iters, long long* total_supports) {
  Quiet q;
  Forest forest0;
  Forest::FilterMask fm = forest0.readForest(forest_path, w, h);
  std::vector<long long> sums((size_t)threads, 0);
  std::vector<std::thread> pool;
  size_t P = (size_t)w * h;
  auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t]() {
      Forest forest;
      Forest::FilterMask fmt = fm;
      Settings s = make_settings(thr, disp_high, vt, epipolar, 1);
      for (int it = 0; it < iters; it++) {
        int k = (t * iters + it) % n_pairs;
        ndb::Buffer<uint8_t> simg = to_buffer(images + (size_t)(2 * k) * P, w, h);
        ndb::Buffer<uint8_t> timg = to_buffer(images + (size_t)(2 * k + 1) * P, w, h);
        Forest::PreprocessedImage sp = forest.preprocessImage(simg, s);
        Forest::PreprocessedImage tp = forest.preprocessImage(timg, s);
        zero_unwritten_row(sp);
        zero_unwritten_row(tp);
        std::vector<ndb::Support> out = forest.rectifiedMatch(sp, tp, fmt, s);
        sums[(size_t)t] += (long long)out.size();
      }
    });
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  long long tot = 0;
  for (long long v : sums) tot += v;
  if (total_supports) *total_supports = tot;
  return std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0).count();
}

}  // extern "C"

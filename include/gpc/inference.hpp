// gpc/inference.hpp -- source-compatible stand-in for the reference's lib/gpc/inference.hpp whose
// bodies forward to the B200 implementation through the C ABI of include/gpc_b200.h.
// samples/sparsematch.cpp compiles unchanged against this header (link with -lgpc_b200 -lz).
//
// Mirrors (same names, fields, argument meaning, printed messages):
//   gpc::inference::sysTick / tickToMs / time_point                      inference.hpp:61-70
//   gpc::inference::InferenceSettings + fluent builder                   inference.hpp:71-131
//   Forest::FilterMask / PreprocessedImage                               inference.hpp:137-166
//   Forest::readForest                                                   inference.hpp:404-446
//   Forest::preprocessImage                                              inference.hpp:302-333
//   Forest::rectifiedMatch / stereoMatch / depthPriorFast                inference.hpp:375-393, :344-361, :184-226
//   Forest::evalFastMaskOnSubsetSSE                                      inference.hpp:266-292
//   Forest::findCorrespondences                                          inference.hpp:227-254
//
// What differs from the reference, by design:
//   * There is no CPU path.  All pixel work runs in the CUDA kernels of libgpc_b200.so; if no
//     device is available the first call throws gpc::inference::GpcError.
//   * PreprocessedImage additionally carries a handle to the raw image kept resident on the GPU.
//     rectifiedMatch / stereoMatch on two such images run the fused preprocess+hash kernel and the
//     matcher on the device without another upload.  The host-visible smooth / grad / mask fields
//     are filled for API parity (sparsematch prints mask.size()); editing them does not change
//     what the resident path computes.  PreprocessedImage objects built by hand (no handle) go
//     through evalFastMaskOnSubsetSSE + findCorrespondences on the caller's smooth / mask data.
//   * useHashtable(true) reproduces the reference's hashtable matcher (hashmatch.hpp:48-272: 214673 buckets of
//     at most 10 elements, a different and smaller match set than the sort path).
//   * Results are those of the reference's default build (-D_INTRINSICS_SSE) whether or not that macro is
//     defined here.  Define GPC_B200_NAIVE_RESULTS before including this header to get the results of the
//     reference's SSE=OFF build instead (boxNaive, sobelNaive, gpcFilter[Tau]Naive; forests of at most 31 tests).
//   * numThreads is accepted and ignored.
#ifndef GPC_B200_INFERENCE_HPP
#define GPC_B200_INFERENCE_HPP

#include <sys/stat.h>

#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../gpc_b200.h"
#include "buffer.hpp"

namespace gpc {
namespace inference {

typedef std::chrono::high_resolution_clock::time_point time_point;
inline time_point sysTick() { return std::chrono::high_resolution_clock::now(); }
inline float tickToMs(time_point t0, time_point t1) {
  return (float)std::abs(1000. * std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0).count());
}

// Error raised when the C ABI reports a failure the reference API has no channel for.
struct GpcError : std::runtime_error {
  int status;
  GpcError(int status, const std::string& what) : std::runtime_error(what), status(status) {}
};

struct InferenceSettings {
  uint8_t gradientThreshold_ = 10;
  int dispHigh_ = 128;
  int verticalTolerance_ = 1;
  bool epipolarMode_ = false;
  bool useHashtable_ = false;
  int numThreads_ = 1;

  InferenceSettings(uint8_t gradientThreshold, int dispHigh, int verticalTolerance, bool epipolarMode, bool useHashtable,
                    int numThreads)
      : gradientThreshold_(gradientThreshold), dispHigh_(dispHigh), verticalTolerance_(verticalTolerance),
        epipolarMode_(epipolarMode), useHashtable_(useHashtable), numThreads_(numThreads) {}
  InferenceSettings() {}
  InferenceSettings& builder(void) { return *this; }
  InferenceSettings& gradientThreshold(uint8_t v) { gradientThreshold_ = v; return *this; }
  InferenceSettings& dispHigh(int v) { dispHigh_ = v; return *this; }
  InferenceSettings& verticalTolerance(int v) { verticalTolerance_ = v; return *this; }
  InferenceSettings& epipolarMode(bool v) { epipolarMode_ = v; return *this; }
  InferenceSettings& useHashtable(bool v) { useHashtable_ = v; return *this; }
  InferenceSettings& numThreads(int v) {
    const int hc = (int)std::thread::hardware_concurrency();
    numThreads_ = (v > hc) ? hc : v;
    return *this;
  }
};

namespace detail {

inline gpc_settings to_c(const InferenceSettings& s) {
  gpc_settings c;
  c.gradient_threshold = s.gradientThreshold_;
  c.disp_high = s.dispHigh_;
  c.vertical_tolerance = s.verticalTolerance_;
  c.epipolar_mode = s.epipolarMode_ ? 1 : 0;
  c.use_hashtable = s.useHashtable_ ? 1 : 0;
  c.num_threads = s.numThreads_;
  return c;
}

// A resident context and its lock.  The reference's Forest methods are stateless and may be called from several
// threads; here they share one device context per process, so every call into it holds `mu`.
struct Context {
  gpc_ctx* ctx = nullptr;
  int max_w = 0, max_h = 0, device = 0;
  std::recursive_mutex mu;      // recursive: a lazily fetched smooth / grad buffer may be touched inside a locked call
  ~Context() { if (ctx) gpc_destroy(ctx); }
};

// One current context per process; when a larger image arrives a NEW context is created and becomes current.  The old
// one lives on for as long as PreprocessedImages made with it do (they hold a reference), so their resident images stay
// valid and rectifiedMatch on two images of the same context keeps working.
struct Runtime {
  std::mutex mu;
  std::shared_ptr<Context> cur;
  std::shared_ptr<Context> get(int w, int h) {
    std::lock_guard<std::mutex> lk(mu);
    if (cur && w <= cur->max_w && h <= cur->max_h) return cur;
    auto nc = std::make_shared<Context>();
    const char* dev = std::getenv("GPC_DEVICE");
    nc->device = dev ? std::atoi(dev) : 0;
    nc->max_w = std::max(w, cur ? cur->max_w : 0); nc->max_h = std::max(h, cur ? cur->max_h : 0);
    const int rc = gpc_create(&nc->ctx, nc->device, nc->max_w, nc->max_h, 1);
    if (rc != GPC_OK) throw GpcError(rc, std::string("gpc_create: ") + gpc_last_error(nullptr));
#ifdef GPC_B200_NAIVE_RESULTS
    gpc_set_result_mode(nc->ctx, GPC_RESULTS_NAIVE);   // the reference's SSE=OFF build (filter.hpp *Naive functions)
#endif
    cur = nc;
    return cur;
  }
};

inline Runtime& runtime() {
  static Runtime rt;
  return rt;
}

inline void check(gpc_ctx* c, int rc, const char* where) {
  if (rc != GPC_OK) throw GpcError(rc, std::string(where) + ": " + gpc_last_error(c));
}

// Raw image resident on the device; keeps its context alive.
struct ResidentImage {
  std::shared_ptr<Context> cx;
  gpc_image* image = nullptr;
  int thr = 0;                // gradient threshold of the preprocessImage call that made it
  ~ResidentImage() {
    if (image) { std::lock_guard<std::recursive_mutex> lk(cx->mu); gpc_image_release(image); }
  }
};

}  // namespace detail

class Forest {
 public:
  // Result of readForest.  `mask` holds, per test, the two offsets baked for the image width
  // (ix + iy*W, jx + jy*W) exactly as in the reference (inference.hpp:427-428).
  struct FilterMask {
    std::vector<int32_t> mask;
    std::vector<int> tau;
    int width;
    int height;
    int type;
    FilterMask(std::vector<int32_t> mask, int width, int height, int type)
        : mask(mask), width(width), height(height), type(type) {}
    FilterMask(std::vector<int32_t> mask, std::vector<int> tau, int width, int height, int type)
        : mask(mask), tau(tau), width(width), height(height), type(type) {}
  };

  struct PreprocessedImage {
    ndb::Buffer<uint8_t> smooth;
    ndb::Buffer<uint8_t> grad;
    std::vector<int> mask;
    std::shared_ptr<detail::ResidentImage> resident;   // B200 addition: raw image kept on the device
    PreprocessedImage(ndb::Buffer<uint8_t>& smooth, ndb::Buffer<uint8_t>& grad, std::vector<int>& mask)
        : smooth(smooth), grad(grad), mask(mask) {}
    PreprocessedImage(ndb::Buffer<uint8_t>&& smooth, ndb::Buffer<uint8_t>&& grad, std::vector<int>&& mask)   // B200 addition
        : smooth(std::move(smooth)), grad(std::move(grad)), mask(std::move(mask)) {}
  };

  enum CorrMethod { sorting = 's', hashtable = 'h' };

  // inference.hpp:404-446.  Same text format, same 32-test cap with one "Note:" line per
  // discarded test, same message and empty mask when the file cannot be opened.
  FilterMask readForest(std::string forestPath, int width, int height) {
    std::vector<int32_t> mask;
    std::vector<int> tau;
    gpc_forest f;
    const int rc = gpc_read_forest(forestPath.c_str(), &f);
    if (rc != GPC_OK) {
      std::cout << "Error opening forest file" << std::endl;
      return FilterMask(mask, width, height, 0);
    }
    std::cout << "number of ferns:" << f.n_ferns << std::endl;
    for (int t = 0; t < f.n_tests; t++) {
      mask.push_back(f.ix[t] + f.iy[t] * width);
      mask.push_back(f.jx[t] + f.jy[t] * width);
      tau.push_back(f.tau[t]);
    }
    for (int i = 0; i < f.n_discarded; i++)
      std::cout << "Note: A maximum of 32 fern features are allowed, discarding remainder of forest." << std::endl;
    if (f.type == 0) return FilterMask(mask, width, height, 0);
    return FilterMask(mask, tau, width, height, 1);
  }

  // inference.hpp:302-333: box blur, Sobel mask, candidate indices -- here one fused CUDA kernel;
  // the raw image stays resident on the device for rectifiedMatch.
  // One pass: upload, kernel A1 (and A2 when a forest is already set on the context); the outputs stay cached on the
  // device for rectifiedMatch.  `mask` is filled eagerly (sparsematch.cpp:54-55 reads mask.size()); `smooth` and `grad`
  // are fetched from the device on first access (Buffer::setLazyFill).
  PreprocessedImage preprocessImage(ndb::Buffer<uint8_t>& img, InferenceSettings settings) {
    const int w = img.cols(), h = img.rows();
    assert(w % 16 == 0 && "width must be multiple of 16!");
    auto cx = detail::runtime().get(w, h);
    gpc_ctx* c = cx->ctx;
    auto res = std::make_shared<detail::ResidentImage>();
    res->cx = cx; res->thr = settings.gradientThreshold_;
    std::vector<int> mask;
    {
      std::lock_guard<std::recursive_mutex> lk(cx->mu);
      int n = 0;
      detail::check(c, gpc_image_upload(c, img.data(), w, h, w, &res->image), "gpc_image_upload");
      detail::check(c, gpc_image_preprocess(c, res->image, settings.gradientThreshold_, nullptr, nullptr, nullptr, 0, &n),
                    "gpc_image_preprocess");
      const int32_t* view = gpc_mask_view(c, &n);            // the list sits in the context's pinned staging buffer:
      static_assert(sizeof(int) == sizeof(int32_t), "mask element type");
      if (view && n > 0) mask.assign(view, view + n);         // one copy, no zero fill
    }
    // both images come from one kernel run: whichever is touched first fetches the pair into a shared block
    struct Pair { std::vector<uint8_t> s, g; bool done = false; std::mutex mu; };
    auto pr = std::make_shared<Pair>();
    const size_t P = (size_t)w * h;
    auto fetch = [res, pr, P](uint8_t* dst, bool want_smooth) {
      std::lock_guard<std::recursive_mutex> lk2(res->cx->mu);      // context lock first, always in this order
      std::lock_guard<std::mutex> lk(pr->mu);
      if (!pr->done) {
        pr->s.resize(P); pr->g.resize(P);
        detail::check(res->cx->ctx, gpc_image_fetch(res->cx->ctx, res->image, res->thr, pr->s.data(), pr->g.data()), "gpc_image_fetch");
        pr->done = true;
      }
      std::memcpy(dst, want_smooth ? pr->s.data() : pr->g.data(), P);
    };
    PreprocessedImage out(ndb::Buffer<uint8_t>::lazy(h, w, [fetch](uint8_t* d) { fetch(d, true); }),
                          ndb::Buffer<uint8_t>::lazy(h, w, [fetch](uint8_t* d) { fetch(d, false); }), std::move(mask));
    out.resident = res;
    return out;
  }

  // inference.hpp:375-393
  std::vector<ndb::Support> rectifiedMatch(PreprocessedImage& simg, PreprocessedImage& timg, FilterMask& forestmask,
                                           InferenceSettings settings) {
    std::vector<ndb::Support> supp;
    if (resident_pair(simg, timg)) {
      check_dims(simg, timg, forestmask);
      std::lock_guard<std::recursive_mutex> lk(simg.resident->cx->mu);
      gpc_ctx* c = simg.resident->cx->ctx;
      upload_forest(c, forestmask);
      gpc_settings cs = detail::to_c(settings);
      cs.gradient_threshold = simg.resident->thr;          // the candidates are preprocessImage's (inference.hpp:375-393 never re-thresholds)
      int n = 0;
      static_assert(sizeof(ndb::Support) == sizeof(gpc_support), "Support layout");
      // count first, then fetch exactly that many records (a worst-case vector would cost more than the matcher)
      const int rc = gpc_match_images(c, simg.resident->image, timg.resident->image, &cs, nullptr, 0, &n, nullptr, nullptr);
      if (rc != GPC_E_CAPACITY) detail::check(c, rc, "gpc_match_images");
      supp.resize((size_t)n);
      detail::check(c, gpc_fetch_supports(c, reinterpret_cast<gpc_support*>(supp.data()), n), "gpc_fetch_supports");
      return supp;
    }
    std::vector<ndb::Correspondence> corr = stereoMatch(simg, timg, forestmask, settings);
    for (auto& e : corr)
      if (std::abs(e.srcPt.y - e.tarPt.y) <= settings.verticalTolerance_ && std::abs(e.srcPt.x - e.tarPt.x) <= settings.dispHigh_)
        supp.push_back(ndb::Support(e.srcPt.x, e.srcPt.y, (float)(e.srcPt.x - e.tarPt.x)));
    return supp;
  }

  // inference.hpp:344-361: all unique-unique correspondences, before the disparity filter.
  std::vector<ndb::Correspondence> stereoMatch(PreprocessedImage& simg, PreprocessedImage& timg, FilterMask& forestmask,
                                               InferenceSettings settings) {
    check_dims(simg, timg, forestmask);
    return depthPriorFast(simg, timg, forestmask, settings);
  }

  // inference.hpp:184-226
  std::vector<ndb::Correspondence> depthPriorFast(PreprocessedImage& src, PreprocessedImage& tar, FilterMask& fastmask,
                                                  InferenceSettings& settings) {
    if (resident_pair(src, tar)) {
      std::lock_guard<std::recursive_mutex> lk(src.resident->cx->mu);
      gpc_ctx* c = src.resident->cx->ctx;
      upload_forest(c, fastmask);
      gpc_settings cs = detail::to_c(settings);
      cs.gradient_threshold = src.resident->thr;
      const int cap = std::max(1, (int)std::min(src.mask.size(), tar.mask.size()));
      std::vector<gpc_correspondence> raw((size_t)cap);
      int n = 0;
      detail::check(c, gpc_correspond_images(c, src.resident->image, tar.resident->image, &cs, raw.data(), cap, &n),
                    "gpc_correspond_images");
      std::vector<ndb::Correspondence> corr;
      corr.reserve((size_t)n);
      for (int i = 0; i < n; i++)
        corr.push_back(ndb::Correspondence(ndb::Point(raw[i].xs, raw[i].ys), ndb::Point(raw[i].xt, raw[i].yt)));
      return corr;
    }
    std::vector<ndb::Descriptor> a = evalFastMaskOnSubsetSSE(src.smooth, src.grad, src.mask, fastmask, settings);
    std::vector<ndb::Descriptor> b = evalFastMaskOnSubsetSSE(tar.smooth, tar.grad, tar.mask, fastmask, settings);
    if (settings.epipolarMode_) {
      for (auto& el : a) el.state |= uint64_t(el.point.y) << 32;
      for (auto& el : b) el.state |= uint64_t(el.point.y) << 32;
    }
    if (settings.useHashtable_) return hashMatch(a, b);          // inference.hpp:204-225
    return findCorrespondences(a, b);
  }

  // inference.hpp:266-292: one Descriptor per entry of idx, state from the fern tests evaluated
  // on the caller's smoothed image (CUDA; gpc_hash_smooth).  `grad` is only a skip hint in the
  // reference (filter.hpp:566) and is not needed here.
  std::vector<ndb::Descriptor> evalFastMaskOnSubsetSSE(ndb::Buffer<uint8_t>& img, ndb::Buffer<uint8_t>& grad,
                                                       std::vector<int>& idx, FilterMask& fastmask,
                                                       InferenceSettings& settings) {
    (void)grad; (void)settings;
    const int w = img.cols(), h = img.rows();
    const uint8_t* pixels = img.data();                     // a lazily fetched image materialises here, before the context lock
    auto cx = detail::runtime().get(w, h);
    std::lock_guard<std::recursive_mutex> lk(cx->mu);
    gpc_ctx* c = cx->ctx;
    upload_forest(c, fastmask);
    std::vector<uint32_t> states(idx.size());
    std::vector<int32_t> idx32(idx.begin(), idx.end());
    detail::check(c, gpc_hash_smooth(c, pixels, w, h, idx32.data(), (int)idx32.size(), states.data()), "gpc_hash_smooth");
    std::vector<ndb::Descriptor> out(idx.size());
    for (size_t j = 0; j < idx.size(); j++)
      out[j] = ndb::Descriptor(ndb::Point(idx[j] % w, idx[j] / w), states[j]);
    return out;
  }

  // inference.hpp:227-254 on explicit descriptor lists (CUDA; gpc_find_correspondences).
  std::vector<ndb::Correspondence> findCorrespondences(std::vector<ndb::Descriptor>& srcStates,
                                                       std::vector<ndb::Descriptor>& tarStates) {
    std::vector<ndb::Correspondence> corr;
    if (srcStates.empty() || tarStates.empty()) return corr;
    auto cx = detail::runtime().get(16, 1);
    std::lock_guard<std::recursive_mutex> lk(cx->mu);
    gpc_ctx* c = cx->ctx;
    std::vector<uint64_t> ks(srcStates.size()), kt(tarStates.size());
    for (size_t i = 0; i < ks.size(); i++) ks[i] = srcStates[i].state;
    for (size_t i = 0; i < kt.size(); i++) kt[i] = tarStates[i].state;
    std::vector<int32_t> pairs(2 * std::min(ks.size(), kt.size()));
    int n = 0;
    detail::check(c, gpc_find_correspondences(c, ks.data(), (int)ks.size(), kt.data(), (int)kt.size(), pairs.data(),
                                              (int)(pairs.size() / 2), &n), "gpc_find_correspondences");
    corr.reserve((size_t)n);
    for (int i = 0; i < n; i++) corr.push_back(ndb::Correspondence(srcStates[pairs[2 * i]].point, tarStates[pairs[2 * i + 1]].point));
    return corr;
  }

  // ndb::Hashmatch<Descriptor> as depthPriorFast drives it (inference.hpp:204-225; CUDA, gpc_hashmatch).
  std::vector<ndb::Correspondence> hashMatch(std::vector<ndb::Descriptor>& srcStates, std::vector<ndb::Descriptor>& tarStates) {
    std::vector<ndb::Correspondence> corr;
    if (srcStates.empty() || tarStates.empty()) return corr;
    auto cx = detail::runtime().get(16, 1);
    std::lock_guard<std::recursive_mutex> lk(cx->mu);
    gpc_ctx* c = cx->ctx;
    std::vector<uint64_t> ks(srcStates.size()), kt(tarStates.size());
    for (size_t i = 0; i < ks.size(); i++) ks[i] = srcStates[i].state;
    for (size_t i = 0; i < kt.size(); i++) kt[i] = tarStates[i].state;
    std::vector<int32_t> pairs(2 * std::min(ks.size(), kt.size()));
    int n = 0;
    detail::check(c, gpc_hashmatch(c, ks.data(), (int)ks.size(), kt.data(), (int)kt.size(), pairs.data(), (int)(pairs.size() / 2), &n),
                  "gpc_hashmatch");
    corr.reserve((size_t)n);
    for (int i = 0; i < n; i++) corr.push_back(ndb::Correspondence(srcStates[pairs[2 * i]].point, tarStates[pairs[2 * i + 1]].point));
    return corr;
  }

 private:
  // both images resident on the SAME context (it need not be the current one: a PreprocessedImage keeps its own alive)
  static bool resident_pair(const PreprocessedImage& a, const PreprocessedImage& b) {
    return a.resident && b.resident && a.resident->image && b.resident->image && a.resident->cx == b.resident->cx &&
           a.resident->thr == b.resident->thr;
  }

  static void check_dims(const PreprocessedImage& simg, const PreprocessedImage& timg, const FilterMask& fm) {
    assert((fm.width == simg.smooth.cols() && fm.height == simg.smooth.rows()) &&
           "Source Image: dimension does not fit dimension of supplied forest mask");
    assert((fm.width == timg.smooth.cols() && fm.height == simg.smooth.rows()) &&
           "Targe Image: dimension does not fit dimension of supplied forest mask");
    (void)simg; (void)timg; (void)fm;
  }

  // FilterMask -> gpc_forest: undo the per-width baking (offset = dx + dy*W with |dx| <= 13).
  static void upload_forest(gpc_ctx* c, const FilterMask& fm) {
    gpc_forest f;
    std::memset(&f, 0, sizeof(f));
    const int W = fm.width;
    f.n_tests = (int)std::min<size_t>(fm.mask.size() / 2, GPC_MAX_TESTS);
    f.type = fm.type;
    auto split = [W](int32_t off, int32_t* dx, int32_t* dy) {
      int y = (int)std::floor((double)off / W + 0.5);
      *dy = y; *dx = off - y * W;
    };
    for (int t = 0; t < f.n_tests; t++) {
      split(fm.mask[2 * t], &f.ix[t], &f.iy[t]);
      split(fm.mask[2 * t + 1], &f.jx[t], &f.jy[t]);
      f.tau[t] = (fm.type != 0 && (size_t)t < fm.tau.size()) ? fm.tau[t] : 0;
    }
    detail::check(c, gpc_set_forest(c, &f), "gpc_set_forest");
  }
};

}  // namespace inference
}  // namespace gpc
#endif

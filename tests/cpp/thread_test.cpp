// thread_test -- the drop-in C++ API from several host threads at once, and across a context regrowth.
//   thread_test <forest> <left.png> <right.png> <small_left.png> <small_right.png> <n_threads> <loops>
// The reference's Forest is stateless and re-entrant (SURVEY.md 8b "Threading"); here the threads share one device
// context behind a lock.  Every thread must get exactly the single-threaded result; images preprocessed before a
// larger image forced a new context must still match (their context stays alive with them).  Prints "ok <n> <n_small>".
#include <atomic>
#include <cstdlib>
#include <iostream>
#include <thread>

#include "gpc/inference.hpp"

namespace gi = gpc::inference;

static bool same(const std::vector<ndb::Support>& a, const std::vector<ndb::Support>& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++) if (a[i].x != b[i].x || a[i].y != b[i].y || a[i].d != b[i].d) return false;
  return true;
}

int main(int argc, char** argv) {
  if (argc < 8) { std::cerr << "usage: thread_test forest left right small_left small_right n_threads loops\n"; return 2; }
  gi::Forest forest;
  gi::InferenceSettings st = gi::InferenceSettings().builder().gradientThreshold(5).verticalTolerance(0).dispHigh(128).epipolarMode(true);
  ndb::Buffer<uint8_t> L, R, l, r;
  if (L.readPNG(argv[2]) || R.readPNG(argv[3]) || l.readPNG(argv[4]) || r.readPNG(argv[5])) return 3;
  const int nt = std::atoi(argv[6]), loops = std::atoi(argv[7]);
  try {
    // small pair first, then the large one: the runtime creates a second, larger context
    gi::Forest::FilterMask fms = forest.readForest(argv[1], l.cols(), l.rows());
    gi::Forest::PreprocessedImage lp = forest.preprocessImage(l, st), rp = forest.preprocessImage(r, st);
    const std::vector<ndb::Support> small0 = forest.rectifiedMatch(lp, rp, fms, st);
    gi::Forest::FilterMask fm = forest.readForest(argv[1], L.cols(), L.rows());
    gi::Forest::PreprocessedImage LP = forest.preprocessImage(L, st), RP = forest.preprocessImage(R, st);
    const std::vector<ndb::Support> big0 = forest.rectifiedMatch(LP, RP, fm, st);
    if (!same(forest.rectifiedMatch(lp, rp, fms, st), small0)) { std::cerr << "small pair differs after the context grew\n"; return 4; }
    if (!same(forest.rectifiedMatch(LP, RP, fm, st), big0)) { std::cerr << "repeated rectifiedMatch differs\n"; return 4; }
    // mixed contexts: a small left image against itself, preprocessed again on the new context
    gi::Forest::PreprocessedImage rp2 = forest.preprocessImage(r, st);
    if (!same(forest.rectifiedMatch(lp, rp2, fms, st), small0)) { std::cerr << "images of two contexts: fallback path differs\n"; return 4; }
    // lazily fetched images equal the eager ones of a hand-built copy
    gi::Forest::PreprocessedImage lq(lp.smooth, lp.grad, lp.mask), rq(rp.smooth, rp.grad, rp.mask);
    if (!same(forest.rectifiedMatch(lq, rq, fms, st), small0)) { std::cerr << "hand-built images differ\n"; return 4; }
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
      th.emplace_back([&, t]() {
        try {
          gi::Forest f2;
          for (int k = 0; k < loops; k++) {
            const bool big = ((t + k) & 1) != 0;
            gi::Forest::PreprocessedImage a = f2.preprocessImage(big ? L : l, st), b = f2.preprocessImage(big ? R : r, st);
            if (!same(f2.rectifiedMatch(a, b, big ? fm : fms, st), big ? big0 : small0)) bad++;
            if (k == 0 && a.smooth.getPixel(20, 20) != (big ? LP : lp).smooth.getPixel(20, 20)) bad++;
          }
        } catch (const std::exception& e) { std::cerr << "thread " << t << ": " << e.what() << "\n"; bad++; }
      });
    for (auto& x : th) x.join();
    if (bad) { std::cerr << bad << " threaded results differ\n"; return 5; }
    std::cout << "ok " << big0.size() << " " << small0.size() << std::endl;
  } catch (const gi::GpcError& e) {
    std::cerr << "GpcError " << e.status << ": " << e.what() << "\n";
    return 10;
  }
  return 0;
}

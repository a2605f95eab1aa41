// api_test -- drives the drop-in C++ API (include/gpc/inference.hpp) and dumps everything it
// returns to a flat int32 file that tests/test_cpp_api.py compares against the oracle.
//   api_test <forest> <left.png> <right.png> <out.bin> <epipolar 0|1> <vt> <dispHigh> <thr>
// File layout (int32): nL nR nS nC nS2 nD | maskL | maskR | supports(x,y,(int)d) | corr(xs,ys,xt,yt)
//                      | supports from hand-built PreprocessedImages | states of the left descriptors
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "gpc/inference.hpp"
#include "gpc/inference.hpp"   // include guard check: the header must tolerate double inclusion

int main(int argc, char** argv) {
  if (argc < 9) { std::cerr << "usage: api_test forest left right out epipolar vt dispHigh thr\n"; return 2; }
  namespace gi = gpc::inference;
  gi::Forest forest;
  gi::InferenceSettings st = gi::InferenceSettings().builder().gradientThreshold((uint8_t)std::atoi(argv[8]))
      .verticalTolerance(std::atoi(argv[6])).dispHigh(std::atoi(argv[7])).epipolarMode(std::atoi(argv[5]) != 0).useHashtable(false).numThreads(3);
  ndb::Buffer<uint8_t> left, right;
  if (left.readPNG(argv[2]) || right.readPNG(argv[3])) return 3;
  try {
    gi::Forest::FilterMask fm = forest.readForest(argv[1], left.cols(), left.rows());
    gi::Forest::PreprocessedImage lp = forest.preprocessImage(left, st), rp = forest.preprocessImage(right, st);
    std::vector<ndb::Support> supp = forest.rectifiedMatch(lp, rp, fm, st);
    std::vector<ndb::Correspondence> corr = forest.stereoMatch(lp, rp, fm, st);
    // hand-built PreprocessedImages (no device handle): evalFastMaskOnSubsetSSE + findCorrespondences path
    gi::Forest::PreprocessedImage lq(lp.smooth, lp.grad, lp.mask), rq(rp.smooth, rp.grad, rp.mask);
    std::vector<ndb::Support> supp2 = forest.rectifiedMatch(lq, rq, fm, st);
    std::vector<ndb::Descriptor> desc = forest.evalFastMaskOnSubsetSSE(lp.smooth, lp.grad, lp.mask, fm, st);
    // useHashtable(true): the reference's hashtable matcher, on resident and on hand-built images
    gi::InferenceSettings hs = st;
    hs.useHashtable(true);
    std::vector<ndb::Support> supp_ht = forest.rectifiedMatch(lp, rp, fm, hs);
    std::vector<ndb::Support> supp_ht2 = forest.rectifiedMatch(lq, rq, fm, hs);
    if (supp_ht.size() != supp_ht2.size()) { std::cerr << "useHashtable(true): resident and hand-built paths differ\n"; return 4; }
    for (size_t i = 0; i < supp_ht.size(); i++)
      if (supp_ht[i].x != supp_ht2[i].x || supp_ht[i].y != supp_ht2[i].y || supp_ht[i].d != supp_ht2[i].d) return 4;
    std::vector<int32_t> out = {(int32_t)lp.mask.size(), (int32_t)rp.mask.size(), (int32_t)supp.size(), (int32_t)corr.size(),
                                (int32_t)supp2.size(), (int32_t)desc.size(), (int32_t)supp_ht.size()};
    for (int v : lp.mask) out.push_back(v);
    for (int v : rp.mask) out.push_back(v);
    for (auto& s : supp) { out.push_back(s.x); out.push_back(s.y); out.push_back((int32_t)s.d); }
    for (auto& c : corr) { out.push_back(c.srcPt.x); out.push_back(c.srcPt.y); out.push_back(c.tarPt.x); out.push_back(c.tarPt.y); }
    for (auto& s : supp2) { out.push_back(s.x); out.push_back(s.y); out.push_back((int32_t)s.d); }
    for (auto& d : desc) out.push_back((int32_t)(uint32_t)d.state);
    for (auto& s : supp_ht) { out.push_back(s.x); out.push_back(s.y); out.push_back((int32_t)s.d); }
    FILE* fp = std::fopen(argv[4], "wb");
    if (!fp) return 5;
    std::fwrite(out.data(), sizeof(int32_t), out.size(), fp);
    std::fclose(fp);
    // PNG round trip of the writer/reader pair
    ndb::Buffer<ndb::RGBColor> vis = ndb::getDisparityVisualization(left, supp);
    vis.writePNGRGB(std::string(argv[4]) + ".png");
    left.writePNG(std::string(argv[4]) + ".gray.png");
    ndb::Buffer<uint8_t> back;
    if (back.readPNG(std::string(argv[4]) + ".gray.png")) return 6;
    for (int y = 0; y < left.height; y++) for (int x = 0; x < left.width; x++) if (back.getPixel(x, y) != left.getPixel(x, y)) return 7;
  } catch (const gi::GpcError& e) {
    std::cerr << "GpcError " << e.status << ": " << e.what() << "\n";
    return 10;
  }
  return 0;
}

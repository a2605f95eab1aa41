// sparsematch -- command-line driver of the B200 Global Patch Collider path.
//
// Same contract as the reference's samples/sparsematch.cpp: three positional arguments
// (<forest path> <left image path> <right image path>), the settings of sparsematch.cpp:29-34
// (gradient threshold 5, vertical tolerance 0, dispHigh 128, epipolar mode, sort matcher), the
// same "tPreprocess ... num matches" statistics line on stdout and a disparity.png overlay.
// The reference's own sparsematch.cpp also compiles unchanged against include/gpc/inference.hpp
// (tests/test_cpp_api.py builds it when the reference tree is present); this file exists so the
// repo ships a driver of its own and adds two optional arguments:
//   sparsematch <forest> <left.png> <right.png> [<out.png> [<repeat>]]
// <repeat> > 1 re-runs the timed window on the resident context and prints the best time.
#include <cstdlib>
#include <iostream>

#include "gpc/inference.hpp"

int main(int argc, char** argv) {
  if (argc < 4) {
    std::cout << "Usage: " << argv[0] << " <forest path> <left image path> <right image path> [<out png> [<repeat>]]" << std::endl;
    return 2;
  }
  const std::string forestPath = argv[1], leftPath = argv[2], rightPath = argv[3];
  const std::string outPath = argc > 4 ? argv[4] : "disparity.png";
  const int repeat = argc > 5 ? std::max(1, std::atoi(argv[5])) : 1;

  namespace gi = gpc::inference;
  gi::Forest forest;
  gi::InferenceSettings settings =
      gi::InferenceSettings().builder().gradientThreshold(5).verticalTolerance(0).dispHigh(128).epipolarMode(true).useHashtable(false);

  ndb::Buffer<uint8_t> left, right;
  if (left.readPNG(leftPath) || right.readPNG(rightPath)) {
    std::cout << "No image data \n";
    return -1;
  }
  std::cout << "Using CUDA kernels (sm_100a) through libgpc_b200" << std::endl;
  try {
    gi::Forest::FilterMask fm = forest.readForest(forestPath, left.cols(), left.rows());
    float bestPre = 1e30f, bestMatch = 1e30f;
    size_t nl = 0, nr = 0;
    std::vector<ndb::Support> supp;
    for (int it = 0; it < repeat; it++) {
      gi::time_point t0 = gi::sysTick();
      gi::Forest::PreprocessedImage lp = forest.preprocessImage(left, settings);
      gi::Forest::PreprocessedImage rp = forest.preprocessImage(right, settings);
      gi::time_point t1 = gi::sysTick();
      supp = forest.rectifiedMatch(lp, rp, fm, settings);
      gi::time_point t2 = gi::sysTick();
      bestPre = std::min(bestPre, gi::tickToMs(t1, t0));
      bestMatch = std::min(bestMatch, gi::tickToMs(t2, t1));
      nl = lp.mask.size(); nr = rp.mask.size();
    }
    std::cout << "tPreprocess: " << bestPre << " ms"
              << ", #candidatesL:" << nl << ", #candidatesR:" << nr << ", tMatch: " << bestMatch << " ms"
              << ", num matches:" << supp.size() << std::endl;
    ndb::Buffer<ndb::RGBColor> vis = ndb::getDisparityVisualization(left, supp);
    vis.writePNGRGB(outPath);
  } catch (const gi::GpcError& e) {
    std::cout << "ERR: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}

// png_test -- CPU-only check of the zlib PNG codec in include/gpc/buffer.hpp.
//   png_test <in.png> <out_gray.png> <out.raw>    (out.raw = height*cols bytes of the Buffer<uint8_t>)
#include <cstdio>
#include "gpc/buffer.hpp"
int main(int argc, char** argv) {
  if (argc < 4) return 2;
  ndb::Buffer<uint8_t> img;
  if (img.readPNG(argv[1])) return 3;
  if (img.cols() % 16 != 0 || img.rows() != img.height || img.cols() < img.width) return 4;
  img.writePNG(argv[2]);
  FILE* fp = std::fopen(argv[3], "wb");
  if (!fp) return 5;
  int32_t dims[3] = {img.width, img.height, img.cols()};
  std::fwrite(dims, 4, 3, fp);
  std::fwrite(img.data(), 1, (size_t)img.rows() * img.cols(), fp);
  std::fclose(fp);
  ndb::Buffer<uint8_t> sm(img.height, img.width, 7);
  sm.clearBoundary();
  if (sm.getPixel(0, 5) != 0 || sm.getPixel(1, 5) != 0 || sm.getPixel(2, 5) != 7 || sm.getPixel(5, 0) != 0 ||
      sm.getPixel(5, img.height - 2) != 0 || sm.getPixel(sm.cols() - 1, 5) != 0) return 6;
  return 0;
}

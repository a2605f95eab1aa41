// gpc/buffer.hpp -- ndb:: image container and POD types of the openGPC inference API, re-provided
// without Eigen or libpng so that code written against the reference's lib/gpc/buffer.hpp
// (samples/sparsematch.cpp in particular) compiles unchanged against the B200 implementation.
//
// Mirrors (names, fields, argument meaning):
//   ndb::RGBColor / Point / Descriptor / Support / Correspondence      buffer.hpp:41-102
//   ndb::Buffer<T>: ctor + ALIGN16 width padding, data/rows/cols, width/height,
//     getPixel/setPixel/set, clearBoundary, convertToRGB               buffer.hpp:142-193, :475-560, :630-654
//   Buffer::readPNG / writePNG / writePNGRGB                           buffer.hpp:197-474
//   ndb::getDisparityVisualization(img, supports)                      buffer.hpp:949-1014
// Differences, all deliberate: storage is zero-initialised (the reference leaves it uninitialised,
// which is what makes its output vary run to run); every function is inline or a
// template so the header can be included from several translation units; PNG I/O is a small
// zlib-based codec (8/16-bit gray and 8-bit RGB, non-interlaced) instead of libpng.
#ifndef GPC_B200_BUFFER_HPP
#define GPC_B200_BUFFER_HPP

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace ndb {

struct RGBColor {
  uint8_t b, g, r;
  RGBColor(uint8_t r, uint8_t g, uint8_t b) : b(b), g(g), r(r) {}
  RGBColor() : b(0), g(0), r(0) {}
};

struct Point {
  int x, y;
  Point(int x, int y) : x(x), y(y) {}
  Point() : x(0), y(0) {}
};

struct Descriptor {
  Point point;
  uint64_t state = 0;
  bool srcDescr = false;
  Descriptor(Point point, uint64_t state) : point(point), state(state) {}
  Descriptor() {}
  bool operator==(const Descriptor& d) const { return state == d.state; }
  bool operator!=(const Descriptor& d) const { return state != d.state; }
  bool operator<(const Descriptor& d) const { return state < d.state; }
  bool operator<=(const Descriptor& d) const { return state <= d.state; }
  bool diffImgs(const Descriptor& d) { return srcDescr != d.srcDescr; }
  int operator%(const int& d) const { return (int)(state % (uint64_t)d); }
};

// layout-identical to gpc_support (include/gpc_b200.h): 12 bytes
struct Support {
  int x, y;
  float d;
  Support(int x, int y, float d) : x(x), y(y), d(d) {}
  Support(int x, int y) : x(x), y(y), d(0.f) {}
  Support() : x(0), y(0), d(0.f) {}
};

struct Correspondence {
  Point srcPt, tarPt;
  Correspondence(Point srcPt, Point tarPt) : srcPt(srcPt), tarPt(tarPt) {}
  Correspondence() {}
};

struct Dimension {
  int w, h;
  Dimension(int w, int h) : w(w), h(h) {}
};

#ifndef ALIGN16
#define ALIGN16(X) ((X) % 16) == 0 ? (X) : (((X) / 16) + 1) * 16
#endif

namespace detail {

inline int align16(int x) { return (x % 16) == 0 ? x : ((x / 16) + 1) * 16; }

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put_be32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

struct PngImage {
  int width = 0, height = 0, channels = 0, bit_depth = 0, color_type = -1;
  std::vector<uint8_t> pixels;   // [height][width*channels*(bit_depth/8)], unfiltered
};

// Minimal PNG reader: non-interlaced, bit depth 8 or 16, colour type 0 (gray), 2 (RGB), 6 (RGBA).
// Returns an empty string on success, else the reference-style error text.
inline std::string png_decode(const std::string& filename, PngImage* out) {
  FILE* fp = std::fopen(filename.c_str(), "rb");
  if (!fp) return "ERR: File" + filename + " could not be opened for reading";
  std::vector<uint8_t> file;
  uint8_t buf[65536];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), fp)) > 0) file.insert(file.end(), buf, buf + n);
  std::fclose(fp);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return "ERR: File" + filename + " is not recognized as a PNG file";
  std::vector<uint8_t> idat;
  int interlace = 0;
  size_t pos = 8;
  bool have_ihdr = false;
  while (pos + 12 <= file.size()) {
    uint32_t len = be32(&file[pos]);
    if (pos + 12 + (size_t)len > file.size()) break;
    const uint8_t* type = &file[pos + 4];
    const uint8_t* data = &file[pos + 8];
    if ((uint32_t)crc32(0L, type, (uInt)(len + 4)) != be32(data + len)) return "ERR: Error during read_image";   // libpng rejects a bad CRC too
    if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
      out->width = (int)be32(data); out->height = (int)be32(data + 4);
      out->bit_depth = data[8]; out->color_type = data[9]; interlace = data[12];
      have_ihdr = true;
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr) return "ERR: Error during init_io";
  if (out->width <= 0 || out->height <= 0 || out->width > (1 << 16) || out->height > (1 << 16))
    return "ERR: Error during read_image";                // IHDR sizes are unsigned 31-bit: no negative / absurd allocations
  switch (out->color_type) {
    case 0: out->channels = 1; break;
    case 2: out->channels = 3; break;
    case 6: out->channels = 4; break;
    default: out->channels = 0; break;
  }
  if (out->channels == 0 || interlace != 0 || (out->bit_depth != 8 && out->bit_depth != 16))
    return "ERR: Error during read_image";
  const int bpp = out->channels * out->bit_depth / 8;
  const size_t stride = (size_t)out->width * bpp;
  std::vector<uint8_t> raw((stride + 1) * (size_t)out->height);
  uLongf raw_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
    return "ERR: Error during read_image";
  out->pixels.assign(stride * (size_t)out->height, 0);
  std::vector<uint8_t> zero(stride, 0);
  for (int y = 0; y < out->height; y++) {
    const uint8_t ft = raw[(stride + 1) * (size_t)y];
    const uint8_t* src = &raw[(stride + 1) * (size_t)y + 1];
    uint8_t* dst = &out->pixels[stride * (size_t)y];
    const uint8_t* up = y ? dst - stride : zero.data();
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= (size_t)bpp ? dst[i - bpp] : 0, b = up[i], c = i >= (size_t)bpp ? up[i - bpp] : 0;
      int pred = 0;
      switch (ft) {
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) / 2; break;
        case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                  pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
        default: pred = 0; break;
      }
      dst[i] = (uint8_t)(src[i] + pred);
    }
  }
  return "";
}

// Minimal PNG writer: 8-bit, colour type 0 (gray) or 2 (RGB), filter 0, one IDAT.
inline bool png_encode(const std::string& filename, const uint8_t* pixels, int width, int height, int channels) {
  const size_t stride = (size_t)width * channels;
  std::vector<uint8_t> raw((stride + 1) * (size_t)height);
  for (int y = 0; y < height; y++) {
    raw[(stride + 1) * (size_t)y] = 0;
    std::memcpy(&raw[(stride + 1) * (size_t)y + 1], pixels + stride * (size_t)y, stride);
  }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  comp.resize(clen);
  std::vector<uint8_t> f = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  auto chunk = [&f](const char* type, const std::vector<uint8_t>& data) {
    put_be32(f, (uint32_t)data.size());
    const size_t start = f.size();
    f.insert(f.end(), type, type + 4);
    f.insert(f.end(), data.begin(), data.end());
    put_be32(f, (uint32_t)crc32(0L, &f[start], (uInt)(f.size() - start)));
  };
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, (uint32_t)width); put_be32(ihdr, (uint32_t)height);
  ihdr.push_back(8); ihdr.push_back(channels == 3 ? 2 : 0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk("IHDR", ihdr);
  chunk("IDAT", comp);
  chunk("IEND", {});
  FILE* fp = std::fopen(filename.c_str(), "wb");
  if (!fp) return false;
  const bool ok = std::fwrite(f.data(), 1, f.size(), fp) == f.size();
  std::fclose(fp);
  return ok;
}

}  // namespace detail

// Row-major image container; the column count is padded to a multiple of 16 (ALIGN16,
// buffer.hpp:142-193) while width/height keep the visible size.
template <class T>
class Buffer {
 public:
  int width = 0;
  int height = 0;

  Buffer() {}
  Buffer(const int r, const int c) : width(c), height(r), rows_(r), cols_(detail::align16(c)), v_((size_t)rows_ * cols_) {}
  Buffer(const int r, const int c, T color) : width(c), height(r), rows_(r), cols_(detail::align16(c)), v_((size_t)rows_ * cols_, color) {}
  // copies of a buffer whose pixels have not been fetched yet stay lazy (each copy fetches into its own storage)
  Buffer(const Buffer& o) { *this = o; }
  Buffer(Buffer&& o) noexcept { *this = std::move(o); }
  Buffer& operator=(const Buffer& o) {
    if (this == &o) return *this;
    width = o.width; height = o.height; rows_ = o.rows_; cols_ = o.cols_; lazy_ = o.lazy_;
    const bool p = o.pending_.load(std::memory_order_acquire);
    if (p) v_.clear(); else v_ = o.v_;
    pending_.store(p, std::memory_order_release);
    return *this;
  }
  Buffer& operator=(Buffer&& o) noexcept {
    if (this == &o) return *this;
    width = o.width; height = o.height; rows_ = o.rows_; cols_ = o.cols_; lazy_ = std::move(o.lazy_);
    v_ = std::move(o.v_);
    pending_.store(o.pending_.load(std::memory_order_acquire), std::memory_order_release);
    return *this;
  }

  // B200 addition: the pixels may be produced on first access (Forest::preprocessImage leaves smooth / grad on the
  // device until somebody reads them).  `fill` receives data(); dimensions are known from the start, the storage
  // itself is allocated on first access too.
  static Buffer lazy(const int r, const int c, std::function<void(T*)> fill) {
    Buffer b;
    b.width = c; b.height = r; b.rows_ = r; b.cols_ = detail::align16(c);
    b.lazy_ = std::make_shared<std::function<void(T*)>>(std::move(fill));
    b.pending_.store(true, std::memory_order_release);
    return b;
  }
  void setLazyFill(std::function<void(T*)> fill) {
    lazy_ = std::make_shared<std::function<void(T*)>>(std::move(fill));
    pending_.store(true, std::memory_order_release);
  }
  bool isLazy() const { return pending_.load(std::memory_order_acquire); }

  T* data() { materialize(); return v_.data(); }
  const T* data() const { materialize(); return v_.data(); }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  long size() const { return (long)rows_ * cols_; }
  T& operator()(int r, int c) { materialize(); return v_[(size_t)r * cols_ + c]; }
  const T& operator()(int r, int c) const { materialize(); return v_[(size_t)r * cols_ + c]; }

  // Eigen-style resize: contents unspecified afterwards (here: zero)
  void resize(int r, int c) { drop_lazy(); rows_ = r; cols_ = c; v_.assign((size_t)r * c, T()); }
  void conservativeResize(int r, int c) {
    materialize();
    std::vector<T> nv((size_t)r * c, T());
    for (int y = 0; y < std::min(r, rows_); y++)
      for (int x = 0; x < std::min(c, cols_); x++) nv[(size_t)y * c + x] = v_[(size_t)y * cols_ + x];
    v_.swap(nv); rows_ = r; cols_ = c;
  }

  void setPixel(int x, int y, T color) { materialize(); v_[(size_t)cols_ * y + x] = color; }
  T getPixel(int x, int y) const { materialize(); return v_[(size_t)cols_ * y + x]; }
  void set(T color) { drop_lazy(); v_.assign((size_t)rows_ * cols_, color); }
  Dimension getDimension() { return Dimension(cols_, rows_); }

  // buffer.hpp:630-654
  void clearBoundary() {
    materialize();
    const int h = height, w = width, wa = cols_;
    for (int x = 0; x < 2; x++) for (int y = 0; y < h; y++) v_[(size_t)y * wa + x] = T();
    for (int x = 0; x < w; x++) v_[x] = T();
    for (int x = 0; x < w; x++) for (int y = h - 2; y < h; y++) v_[(size_t)y * wa + x] = T();
    for (int y = 0; y < h; y++) v_[(size_t)y * wa + (wa - 1)] = T();
  }

  Buffer<RGBColor> convertToRGB() const {
    Buffer<RGBColor> out(height, width);
    for (int y = 0; y < height; y++)
      for (int x = 0; x < width; x++) {
        const uint8_t c = (uint8_t)getPixel(x, y);
        out.setPixel(x, y, RGBColor(c, c, c));
      }
    return out;
  }

  // buffer.hpp:197-318: gray as is, RGB -> (r+g+b)/3, 16-bit truncated to T; returns 1 on error.
  // The padding columns (width..cols-1) are zero here (uninitialised in the reference).
  int readPNG(std::string filename) {
    detail::PngImage img;
    const std::string err = detail::png_decode(filename, &img);
    if (!err.empty()) { std::cout << err << std::endl; return 1; }
    drop_lazy();
    width = img.width; height = img.height;
    rows_ = height; cols_ = detail::align16(width);
    v_.assign((size_t)rows_ * cols_, T());
    if (img.channels == 4 || img.channels == 0) {
      std::cout << "ERR: found something other than gray or 3 channel color image(" << img.color_type << ") aborting!" << std::endl;
      return 1;
    }
    const size_t stride = (size_t)img.width * img.channels * (img.bit_depth / 8);
    for (int y = 0; y < height; y++) {
      const uint8_t* row = &img.pixels[stride * (size_t)y];
      for (int x = 0; x < width; x++) {
        int val;
        if (img.bit_depth == 16) val = ((int)row[x * 2] << 8) + row[x * 2 + 1];          // :280-288 (reads as gray)
        else if (img.channels == 1) val = row[x];
        else val = (row[3 * x] + row[3 * x + 1] + row[3 * x + 2]) / 3;                   // :299
        v_[(size_t)y * cols_ + x] = (T)val;
      }
    }
    return 0;
  }

  void writePNG(std::string filename) {
    std::vector<uint8_t> px((size_t)width * height);
    for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) px[(size_t)y * width + x] = (uint8_t)getPixel(x, y);
    if (!detail::png_encode(filename, px.data(), width, height, 1))
      std::cout << "ERR: File" << filename << " could not be opened for writing" << std::endl;
  }

  void writePNGRGB(std::string filename) { writeRGBImpl(filename, *this); }

 private:
  static void writeRGBImpl(const std::string& filename, const Buffer<RGBColor>& b) {
    std::vector<uint8_t> px((size_t)b.width * b.height * 3);
    for (int y = 0; y < b.height; y++)
      for (int x = 0; x < b.width; x++) {
        const RGBColor c = b.getPixel(x, y);
        uint8_t* p = &px[((size_t)y * b.width + x) * 3];
        p[0] = c.r; p[1] = c.g; p[2] = c.b;
      }
    if (!detail::png_encode(filename, px.data(), b.width, b.height, 3))
      std::cout << "ERR: File" << filename << " could not be opened for writing" << std::endl;
  }
  template <class U> static void writeRGBImpl(const std::string&, const Buffer<U>&) {
    std::cout << "ERR: writePNGRGB needs a Buffer<RGBColor>" << std::endl;
  }

  // Reading a PreprocessedImage from several threads is safe in the reference (plain const reads), so the first access
  // that fetches the pixels is made safe here: double-checked under one lock per element type.  The fill may take the
  // device-context lock; nothing enters materialize() while holding that lock (inference.hpp touches lazy buffers first).
  void materialize() const {
    if (!pending_.load(std::memory_order_acquire)) return;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!pending_.load(std::memory_order_relaxed)) return;
    if (v_.size() != (size_t)rows_ * cols_) v_.assign((size_t)rows_ * cols_, T());
    if (lazy_) (*lazy_)(v_.data());
    pending_.store(false, std::memory_order_release);
  }
  void drop_lazy() { pending_.store(false, std::memory_order_release); lazy_.reset(); }

  int rows_ = 0, cols_ = 0;
  mutable std::vector<T> v_;
  mutable std::shared_ptr<std::function<void(T*)>> lazy_;
  mutable std::atomic<bool> pending_{false};
};

// buffer.hpp:949-1014: supports drawn over the gray image with the KITTI disparity colour map.  The eight-entry table
// and the interpolation below ARE Geiger's KITTI stereo devkit colour map as the reference uses it: a restatement kept
// value for value because disparity.png has to come out pixel-identical to the reference's (host-side visualisation only).
inline Buffer<RGBColor> getDisparityVisualization(Buffer<uint8_t>& srcImg, std::vector<Support>& support) {
  const float min_disparity = 0.f, max_disparity = 128.f;
  Buffer<RGBColor> vis = srcImg.convertToRGB();
  const float map[8][4] = {{0, 0, 1, 185}, {1, 0, 0, 114}, {1, 0, 1, 174}, {0, 1, 0, 114},
                           {0, 1, 1, 185}, {1, 1, 0, 114}, {1, 1, 1, 0},   {0, 0, 0, 114}};
  float sum = 0;
  for (int i = 0; i < 8; ++i) sum += map[i][3];
  float weights[8], cumsum[8];
  cumsum[0] = 0;
  for (int i = 0; i < 7; ++i) {
    weights[i] = sum / map[i][3];
    cumsum[i + 1] = cumsum[i] + map[i][3] / sum;
  }
  for (auto& s : support) {
    if (s.x < 0 || s.y < 0 || s.x >= vis.width || s.y >= vis.height) continue;
    const float value = std::max(0.f, std::min(0.8f, (s.d - min_disparity) / (max_disparity - min_disparity)));
    int bin;
    for (bin = 0; bin < 7; ++bin)
      if (value < cumsum[bin + 1]) break;
    const float w = 1.0f - (value - cumsum[bin]) * weights[bin];
    const uint8_t r = (uint8_t)((w * map[bin][0] + (1.0f - w) * map[bin + 1][0]) * 255.0f);
    const uint8_t g = (uint8_t)((w * map[bin][1] + (1.0f - w) * map[bin + 1][1]) * 255.0f);
    const uint8_t b = (uint8_t)((w * map[bin][2] + (1.0f - w) * map[bin + 1][2]) * 255.0f);
    vis.setPixel(s.x, s.y, RGBColor(r, g, b));
  }
  return vis;
}

}  // namespace ndb
#endif

// Stand-in for <opencv4/opencv2/opencv.hpp>: core + cv::imwrite for single-channel PNG.  OpenCV's default PNG settings
// are mirrored (filter SUB, zlib level 1, strategy RLE; a CV_32F image is converted to 8 bits first) so the reference's
// timed span costs what it would cost with the real library; the bytes are not claimed identical (PNG is lossless,
// parity is on decoded pixels).  An optional hook lets the test shim capture the matrices the reference hands to
// imwrite.  See ../README.md.
#pragma once
#include <functional>
#include <zlib.h>
#include "core.hpp"

namespace cv {
namespace stub {
typedef std::function<bool(const std::string&, const Mat&)> ImwriteHook;
inline ImwriteHook& imwrite_hook() { static ImwriteHook h; return h; }

inline void be32(std::vector<uchar>& v, std::uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
inline void chunk(std::vector<uchar>& out, const char* tag, const uchar* p, std::size_t n) {
  be32(out, static_cast<std::uint32_t>(n));
  std::size_t at = out.size();
  out.insert(out.end(), tag, tag + 4);
  if (n) out.insert(out.end(), p, p + n);
  be32(out, static_cast<std::uint32_t>(crc32(0L, out.data() + at, static_cast<uInt>(n + 4))));
}
inline bool png_gray8(const Mat& m, std::vector<uchar>& out) {
  static const uchar magic[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  out.assign(magic, magic + 8);
  std::vector<uchar> hdr; be32(hdr, m.cols); be32(hdr, m.rows);
  const uchar tail[5] = {8, 0, 0, 0, 0}; hdr.insert(hdr.end(), tail, tail + 5);
  chunk(out, "IHDR", hdr.data(), hdr.size());
  std::vector<uchar> raw(static_cast<std::size_t>(m.rows) * (m.cols + 1));
  for (int r = 0; r < m.rows; r++) {
    uchar* d = &raw[static_cast<std::size_t>(r) * (m.cols + 1)]; const uchar* s = m.ptr(r);
    d[0] = 1;                                                      // PNG_FILTER_SUB
    for (int c = 0; c < m.cols; c++) d[1 + c] = static_cast<uchar>(s[c] - (c ? s[c - 1] : 0));
  }
  z_stream zs; std::memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, Z_BEST_SPEED, Z_DEFLATED, 15, 8, Z_RLE) != Z_OK) return false;
  std::vector<uchar> z(deflateBound(&zs, static_cast<uLong>(raw.size())));
  zs.next_in = raw.data(); zs.avail_in = static_cast<uInt>(raw.size()); zs.next_out = z.data(); zs.avail_out = static_cast<uInt>(z.size());
  int rc = deflate(&zs, Z_FINISH); std::size_t zn = zs.total_out; deflateEnd(&zs);
  if (rc != Z_STREAM_END) return false;
  chunk(out, "IDAT", z.data(), zn);
  chunk(out, "IEND", nullptr, 0);
  return true;
}
}  // namespace stub

inline bool imwrite(const std::string& filename, const Mat& img, const std::vector<int>& = std::vector<int>()) {
  if (stub::imwrite_hook()) return stub::imwrite_hook()(filename, img);   // the matrix as the caller passed it
  Mat m = img;
  if (m.type() != CV_8U) img.convertTo(m, CV_8U);
  std::vector<uchar> png;
  if (!stub::png_gray8(m, png)) return false;
  FILE* fp = std::fopen(filename.c_str(), "wb");
  if (!fp) return false;
  bool ok = std::fwrite(png.data(), 1, png.size(), fp) == png.size();
  std::fclose(fp);
  return ok;
}
}  // namespace cv

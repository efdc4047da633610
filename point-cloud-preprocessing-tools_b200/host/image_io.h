// image_io.h — file encoders of the BEV stage, off the GPU critical path (run on the host encode pool).
//
//   write_png_gray8 : replaces cv::imwrite(png, CV_8UC1 Mat)  (BatchMultiBevGen.cpp:318, :361).  Any valid PNG is a
//                     faithful replacement — PNG is lossless and parity is checked on decoded pixels.
//   format_csv_u8   : replaces `f_csv << cv::format(single_bev, cv::Formatter::FMT_CSV)` (:371).  OpenCV's CSVFormatter
//                     (modules/core/src/out.cpp): every value printed with "%3d", values separated by ", ", rows
//                     separated by "\n", and a final "\n" epilogue when cols > 1.
//   format_csv_f32  : same formatter with set32fPrecision(4) => "%.4g" (CloudManip.cpp:97-103).
//   f32_to_u8_sat   : cv::Mat::convertTo(CV_8U) = round-half-to-even + saturate, used by imwrite on a CV_32F Mat
//                     (CloudManip.cpp:108).
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace imgio {

inline void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

inline void png_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
  put_be32(out, (uint32_t)n);
  size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  if (n) out.insert(out.end(), data, data + n);
  uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4));
  put_be32(out, crc);
}

// 8-bit grayscale PNG, filter type 0 on every scanline, one IDAT.  level: zlib level (1 = fast; BEV layers are sparse).
inline bool encode_png_gray8(const uint8_t* pix, int w, int h, std::vector<uint8_t>& out, int level = 1) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  out.clear(); out.insert(out.end(), sig, sig + 8);
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
  ihdr.push_back(8); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);  // depth 8, gray, deflate, adaptive, no interlace
  png_chunk(out, "IHDR", ihdr.data(), ihdr.size());
  std::vector<uint8_t> raw((size_t)h * (w + 1));
  for (int y = 0; y < h; y++) { raw[(size_t)y * (w + 1)] = 0; memcpy(&raw[(size_t)y * (w + 1) + 1], pix + (size_t)y * w, w); }
  // Z_RLE: match distance 1 only - what OpenCV's PNG writer asks of libpng by default (IMWRITE_PNG_STRATEGY_RLE) and far
  // cheaper than hash chains on occupancy layers that are mostly runs of 0
  z_stream zs; memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level, Z_DEFLATED, 15, 8, Z_RLE) != Z_OK) return false;
  std::vector<uint8_t> z(deflateBound(&zs, (uLong)raw.size()));
  zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size(); zs.next_out = z.data(); zs.avail_out = (uInt)z.size();
  const int rc = deflate(&zs, Z_FINISH);
  const size_t cap = zs.total_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) return false;
  png_chunk(out, "IDAT", z.data(), cap);
  png_chunk(out, "IEND", nullptr, 0);
  return true;
}

inline bool write_bytes(const std::string& path, const void* p, size_t n) {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  bool ok = n == 0 || fwrite(p, 1, n, fp) == n;
  fclose(fp);
  return ok;
}

inline bool write_png_gray8(const std::string& path, const uint8_t* pix, int w, int h, int level = 1) {
  std::vector<uint8_t> out;
  if (!encode_png_gray8(pix, w, h, out, level)) return false;
  return write_bytes(path, out.data(), out.size());
}

inline std::string format_csv_u8(const uint8_t* m, int rows, int cols) {
  struct Lut { char t[256][4]; Lut() { for (int i = 0; i < 256; i++) snprintf(t[i], 4, "%3d", i); } };
  static const Lut L;                     // magic static: initialised once, thread-safe (called from the encode pool)
  const char (*lut)[4] = L.t;
  std::string s;
  s.reserve((size_t)rows * cols * 5 + 2);
  for (int r = 0; r < rows; r++) {
    for (int c = 0; c < cols; c++) {
      s.append(lut[m[(size_t)r * cols + c]], 3);
      if (c + 1 < cols) s.append(", ");
    }
    if (r + 1 < rows) s.push_back('\n');
  }
  if (cols > 1) s.push_back('\n');
  return s;
}

inline std::string format_csv_f32(const float* m, int rows, int cols, int prec = 4) {
  std::string s; char fmt[16], buf[64];
  snprintf(fmt, sizeof fmt, "%%.%dg", prec);
  for (int r = 0; r < rows; r++) {
    for (int c = 0; c < cols; c++) {
      float v = m[(size_t)r * cols + c];
      if (std::isnan(v)) snprintf(buf, sizeof buf, "nan");                       // out.cpp prints nan / inf / -inf
      else if (std::isinf(v)) snprintf(buf, sizeof buf, "%s", v > 0 ? "inf" : "-inf");
      else snprintf(buf, sizeof buf, fmt, (double)v);
      s.append(buf);
      if (c + 1 < cols) s.append(", ");
    }
    if (r + 1 < rows) s.push_back('\n');
  }
  if (cols > 1) s.push_back('\n');
  return s;
}

inline uint8_t f32_to_u8_sat(float v) {   // cv::saturate_cast<uchar>(float): cvRound (half to even) then clamp
  if (!(v == v)) return 0;
  double r = std::nearbyint((double)v);
  return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}

}  // namespace imgio

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

// A deflate encoder for run-heavy data (RFC 1951, one fixed-Huffman block inside a zlib wrapper): every byte is a literal
// followed by distance-1 matches covering the rest of its run - the matches libpng's Z_RLE strategy finds (OpenCV's default
// PNG strategy), without zlib's per-byte state machine.  The occupancy layers are mostly runs of 0, so the encoder is bound
// by the 8-byte run scan.  Any inflater reproduces the input exactly (tests/test_image_codec.py decodes with cv2 / zlib).
struct BitSink {
  std::vector<uint8_t>& out; uint64_t acc = 0; int n = 0;
  explicit BitSink(std::vector<uint8_t>& o) : out(o) {}
  inline void put(uint32_t bits, int cnt) {          // LSB-first packing
    acc |= (uint64_t)bits << n; n += cnt;
    while (n >= 8) { out.push_back((uint8_t)acc); acc >>= 8; n -= 8; }
  }
  inline void flush() { if (n > 0) { out.push_back((uint8_t)acc); acc = 0; n = 0; } }
};
struct RleTables {
  uint16_t lit_code[256]; uint8_t lit_bits[256];     // Huffman codes are sent MSB first: stored bit-reversed
  uint32_t len_code[259]; uint8_t len_bits[259];     // length code + its extra bits + the 5-bit distance code 0 (distance 1)
  static uint32_t rev(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i); return r; }
  RleTables() {
    for (int b = 0; b < 256; b++) {
      if (b < 144) { lit_code[b] = (uint16_t)rev(0x30 + b, 8); lit_bits[b] = 8; }
      else { lit_code[b] = (uint16_t)rev(0x190 + (b - 144), 9); lit_bits[b] = 9; }
    }
    static const int base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const int extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    for (int len = 3; len <= 258; len++) {
      int c = 28; while (base[c] > len) c--;
      if (len < 258 && c == 28) c = 27;              // 258 has its own code; 227..257 belong to code 284
      const int sym = 257 + c;
      uint32_t code; int nb;
      if (sym < 280) { code = rev(sym - 256, 7); nb = 7; } else { code = rev(0xC0 + (sym - 280), 8); nb = 8; }
      code |= (uint32_t)(len - base[c]) << nb; nb += extra[c];   // extra bits, LSB first
      nb += 5;                                       // distance code 0 = distance 1: five 0 bits
      len_code[len] = code; len_bits[len] = (uint8_t)nb;
    }
  }
};
inline void deflate_rle(const uint8_t* raw, size_t n, std::vector<uint8_t>& out) {
  static const RleTables T;                           // magic static: thread-safe
  out.push_back(0x78); out.push_back(0x01);           // zlib header: deflate, 32 K window, fastest
  BitSink bs(out);
  bs.put(1, 1); bs.put(1, 2);                         // BFINAL = 1, BTYPE = 01 (fixed Huffman)
  size_t i = 0;
  while (i < n) {
    const uint8_t b = raw[i++];
    bs.put(T.lit_code[b], T.lit_bits[b]);
    size_t j = i;
    const uint64_t pat = 0x0101010101010101ull * b;
    while (j + 8 <= n) { uint64_t w; memcpy(&w, raw + j, 8); if (w != pat) break; j += 8; }
    while (j < n && raw[j] == b) j++;
    size_t run = j - i;
    while (run >= 3) {
      size_t l = run > 258 ? 258 : run;
      if (run - l > 0 && run - l < 3) l = run - 3;    // never strand 1 or 2 bytes behind a full-length match
      bs.put(T.len_code[l], T.len_bits[l]);
      i += l; run -= l;
    }
  }
  bs.put(0, 7);                                       // end of block (symbol 256: seven 0 bits)
  bs.flush();
  const uint32_t ad = (uint32_t)adler32(adler32(0L, Z_NULL, 0), raw, (uInt)n);
  out.push_back(ad >> 24); out.push_back(ad >> 16); out.push_back(ad >> 8); out.push_back(ad);
}

// 8-bit grayscale PNG, filter type 0 on every scanline, one IDAT.  level <= 1: the run-length encoder above; higher levels:
// zlib with the Z_RLE strategy (what OpenCV's PNG writer asks of libpng by default, IMWRITE_PNG_STRATEGY_RLE).
inline bool encode_png_gray8(const uint8_t* pix, int w, int h, std::vector<uint8_t>& out, int level = 1) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  out.clear(); out.insert(out.end(), sig, sig + 8);
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
  ihdr.push_back(8); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);  // depth 8, gray, deflate, adaptive, no interlace
  png_chunk(out, "IHDR", ihdr.data(), ihdr.size());
  std::vector<uint8_t> raw((size_t)h * (w + 1));
  for (int y = 0; y < h; y++) { raw[(size_t)y * (w + 1)] = 0; memcpy(&raw[(size_t)y * (w + 1) + 1], pix + (size_t)y * w, w); }
  std::vector<uint8_t> z;
  if (level <= 1) {
    z.reserve(raw.size() / 8 + 64);
    deflate_rle(raw.data(), raw.size(), z);
  } else {
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, 15, 8, Z_RLE) != Z_OK) return false;
    z.resize(deflateBound(&zs, (uLong)raw.size()));
    zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size(); zs.next_out = z.data(); zs.avail_out = (uInt)z.size();
    const int rc = deflate(&zs, Z_FINISH);
    z.resize(zs.total_out);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
  }
  png_chunk(out, "IDAT", z.data(), z.size());
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
  struct Lut { char t[256][5]; Lut() { for (int i = 0; i < 256; i++) { char b[8]; snprintf(b, sizeof b, "%3d, ", i); memcpy(t[i], b, 5); } } };
  static const Lut L;                     // magic static: initialised once, thread-safe (called from the encode pool)
  if (rows <= 0 || cols <= 0) return std::string();
  // every value is "%3d" (3 bytes for 0..255) + ", ", the last of a row + "\n" instead: 5 bytes per value, one pass
  std::string s((size_t)rows * cols * 5 - (size_t)rows, '\n');
  char* o = &s[0];
  for (int r = 0; r < rows; r++) {
    const uint8_t* p = m + (size_t)r * cols;
    for (int c = 0; c + 1 < cols; c++) { memcpy(o, L.t[p[c]], 5); o += 5; }
    memcpy(o, L.t[p[cols - 1]], 3); o += 3;
    *o++ = '\n';                          // row separator; after the last row it is the epilogue FMT_CSV writes when cols > 1
  }
  if (cols == 1) s.pop_back();            // a single column has no epilogue (out.cpp)
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

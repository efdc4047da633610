// Test helper (CPU only): runs the product's file encoders (host/image_io.h) on raw input files.
// usage: image_probe <u8_224x224.bin> <f32_201x201.bin> <out_prefix>
//   writes <out_prefix>.png (8-bit gray of the u8 image), <out_prefix>.csv (FMT_CSV of the u8 image),
//          <out_prefix>_f.csv (FMT_CSV, "%.4g", of the f32 image), <out_prefix>_f.png (f32 -> u8 as cv::imwrite converts it)
#include <cstdio>
#include <string>
#include <vector>

#include "image_io.h"

static std::vector<uint8_t> slurp(const char* p) {
  std::vector<uint8_t> b; FILE* f = fopen(p, "rb"); if (!f) return b;
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); b.resize((size_t)n);
  if (n && fread(b.data(), 1, (size_t)n, f) != (size_t)n) b.clear();
  fclose(f); return b;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  std::vector<uint8_t> a = slurp(argv[1]), b = slurp(argv[2]);
  if (a.size() != 224 * 224 || b.size() != 201 * 201 * 4) return 1;
  std::string pre = argv[3];
  if (!imgio::write_png_gray8(pre + ".png", a.data(), 224, 224, 1)) return 1;
  std::string t = imgio::format_csv_u8(a.data(), 224, 224);
  if (!imgio::write_bytes(pre + ".csv", t.data(), t.size())) return 1;
  const float* f = reinterpret_cast<const float*>(b.data());
  std::string tf = imgio::format_csv_f32(f, 201, 201, 4);
  if (!imgio::write_bytes(pre + "_f.csv", tf.data(), tf.size())) return 1;
  std::vector<uint8_t> u(201 * 201);
  for (int i = 0; i < 201 * 201; i++) u[i] = imgio::f32_to_u8_sat(f[i]);
  return imgio::write_png_gray8(pre + "_f.png", u.data(), 201, 201, 6) ? 0 : 1;
}

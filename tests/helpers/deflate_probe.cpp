// Test helper: stdin-free probe of imgio::deflate_rle (host/image_io.h).  argv[1] = input file, argv[2] = output file (zlib stream).
#include "image_io.h"

#include <cstdio>
#include <vector>

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  std::vector<uint8_t> in;
  uint8_t buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + n);
  fclose(f);
  std::vector<uint8_t> out;
  imgio::deflate_rle(in.data(), in.size(), out);
  return imgio::write_bytes(argv[2], out.data(), out.size()) ? 0 : 4;
}

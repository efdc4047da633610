// Test helper (CPU only): parses a PCD file with the product's host reader (host/pcd_io.h) and dumps what it saw, so that
// tests/test_pcd_codec.py can compare it with what the Python writer put in.  usage: pcd_probe <in.pcd> <out.bin>
//   stdout: "n=<points> packed=<0|1> stride=<bytes> off=<x,y,z,intensity,row,col,t,label>"
//   out.bin: the 8 SoA arrays (x y z intensity f32, row col u16, t u32, label i16), each of n elements, concatenated
#include <cstdio>
#include <string>

#include "pcd_io.h"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::vector<uint8_t> buf; std::string err;
  if (!pcdio::read_file(argv[1], buf, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  pcdio::Header h;
  if (!pcdio::parse_header(buf, h)) { fprintf(stderr, "bad header\n"); return 1; }
  pcdio::PackedLayout L;
  const bool packed = pcdio::packed_layout(h, L);
  pcdio::Cloud c;
  if (!pcdio::decode(buf, h, argv[1], c, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  printf("n=%zu packed=%d stride=%d off=%d,%d,%d,%d,%d,%d,%d,%d\n", c.size(), packed ? 1 : 0, L.stride, L.off[0], L.off[1], L.off[2], L.off[3],
         L.off[4], L.off[5], L.off[6], L.off[7]);
  if (packed) {   // the interleaved view must give the same values as the by-name decode
    const uint8_t* pay = buf.data() + h.payload_pos;
    for (size_t i = 0; i < c.size(); i++) {
      float x, y, z, it; uint16_t r, cc; uint32_t t; int16_t l;
      pcdio::packed_get(pay, L, i, x, y, z, it, r, cc, t, l);
      if (memcmp(&x, &c.x[i], 4) || memcmp(&y, &c.y[i], 4) || memcmp(&z, &c.z[i], 4) || memcmp(&it, &c.intensity[i], 4) || r != c.row[i] ||
          cc != c.col[i] || t != c.t[i] || l != c.label[i]) { fprintf(stderr, "packed_get differs at %zu\n", i); return 3; }
    }
  }
  FILE* fp = fopen(argv[2], "wb");
  if (!fp) return 1;
  const size_t n = c.size();
  fwrite(c.x.data(), 4, n, fp); fwrite(c.y.data(), 4, n, fp); fwrite(c.z.data(), 4, n, fp); fwrite(c.intensity.data(), 4, n, fp);
  fwrite(c.row.data(), 2, n, fp); fwrite(c.col.data(), 2, n, fp); fwrite(c.t.data(), 4, n, fp); fwrite(c.label.data(), 2, n, fp);
  fclose(fp);
  // round trip through the writer: header + packed 26-byte records
  std::string out = std::string(argv[2]) + ".pcd";
  return pcdio::save_binary(out, c) ? 0 : 1;
}

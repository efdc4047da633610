// pcd_io.h — PCD v0.7 reader/writer for pcl::PointXYZIRCT clouds (BatchMultiBevGen.h:43-66), dependency-free.
//
// Replaces pcl::io::loadPCDFile (BatchMultiBevGen.cpp:730, CloudManip.cpp:117) and pcl::io::savePCDFileBinary
// (BatchMultiBevGen.cpp:756, CloudManip.cpp:139-140).  PCL itself is not vendored by the reference, so the format
// is restated from the PCD v0.7 specification / PCL 1.10 behaviour:
//   * reader: DATA ascii | binary | binary_compressed (LZF, field-major payload); fields are mapped BY NAME with
//     arbitrary order, SIZE/TYPE/COUNT, unknown fields and "_" padding skipped, missing fields left zero
//     (as pcl::fromPCLPointCloud2 does); POINTS wins when WIDTH*HEIGHT disagrees or is 0.
//   * writer: the header PCDWriter::generateHeader<PointXYZIRCT> produces + packed 26-byte records
//     (x y z intensity f32 | row col u16 | t u32 | label i16), WIDTH = n, HEIGHT = 1, VIEWPOINT 0 0 0 1 0 0 0.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace pcdio {

struct Cloud {   // SoA; all vectors have the same length
  std::vector<float> x, y, z, intensity;
  std::vector<uint16_t> row, col;
  std::vector<uint32_t> t;
  std::vector<int16_t> label;
  size_t size() const { return x.size(); }
  void resize(size_t n) { x.resize(n); y.resize(n); z.resize(n); intensity.resize(n); row.resize(n); col.resize(n); t.resize(n); label.resize(n); }
};

struct Field { std::string name; int size = 4; char type = 'F'; int count = 1; int offset = 0; };

// LZF decompression (Marc Lehmann's format as used by PCL's binary_compressed): literal runs and back references.
inline bool lzf_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  const uint8_t* ip = in; const uint8_t* const in_end = in + in_len;
  uint8_t* op = out; uint8_t* const out_end = out + out_len;
  while (ip < in_end) {
    unsigned ctrl = *ip++;
    if (ctrl < 32) {                      // literal run of ctrl+1 bytes
      ctrl++;
      if (op + ctrl > out_end || ip + ctrl > in_end) return false;
      memcpy(op, ip, ctrl); op += ctrl; ip += ctrl;
    } else {                              // back reference
      unsigned len = ctrl >> 5;
      if (ip >= in_end) return false;
      if (len == 7) { len += *ip++; if (ip >= in_end) return false; }
      const uint8_t* ref = op - ((ctrl & 0x1f) << 8) - 1 - *ip++;
      if (ref < out || op + len + 2 > out_end) return false;
      len += 2;
      for (unsigned i = 0; i < len; i++) op[i] = ref[i];   // may overlap: byte-wise
      op += len;
    }
  }
  return op == out_end;
}

template <typename T> inline T rd(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }

inline double field_value(const uint8_t* p, const Field& f) {
  switch (f.type) {
    case 'F': return f.size == 4 ? (double)rd<float>(p) : (f.size == 8 ? rd<double>(p) : 0.0);
    case 'U': return f.size == 1 ? (double)rd<uint8_t>(p) : f.size == 2 ? (double)rd<uint16_t>(p) : f.size == 4 ? (double)rd<uint32_t>(p) : (double)rd<uint64_t>(p);
    case 'I': return f.size == 1 ? (double)rd<int8_t>(p) : f.size == 2 ? (double)rd<int16_t>(p) : f.size == 4 ? (double)rd<int32_t>(p) : (double)rd<int64_t>(p);
  }
  return 0.0;
}

inline void store(Cloud& c, size_t i, int which, double v) {
  switch (which) {
    case 0: c.x[i] = (float)v; break; case 1: c.y[i] = (float)v; break; case 2: c.z[i] = (float)v; break;
    case 3: c.intensity[i] = (float)v; break; case 4: c.row[i] = (uint16_t)(int64_t)v; break; case 5: c.col[i] = (uint16_t)(int64_t)v; break;
    case 6: c.t[i] = (uint32_t)(int64_t)v; break; case 7: c.label[i] = (int16_t)(int64_t)v; break;
  }
}

inline int which_field(const std::string& n) {
  static const char* names[8] = {"x", "y", "z", "intensity", "row", "col", "t", "label"};
  for (int i = 0; i < 8; i++) if (n == names[i]) return i;
  return -1;
}

// Parsed PCD header: fields with their byte offsets inside a record, point count, DATA kind, payload position.
struct Header {
  std::vector<Field> fields;
  size_t n = 0; int rec = 0; std::string data_kind; size_t payload_pos = 0;
};

inline bool read_file(const std::string& path, std::vector<uint8_t>& buf, std::string* err = nullptr) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) { if (err) *err = "cannot open " + path; return false; }
  fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
  if (sz < 0) { fclose(fp); if (err) *err = "cannot stat " + path; return false; }
  buf.resize((size_t)sz);
  if (sz && fread(buf.data(), 1, (size_t)sz, fp) != (size_t)sz) { fclose(fp); if (err) *err = "short read " + path; return false; }
  fclose(fp);
  return true;
}

inline bool parse_header(const std::vector<uint8_t>& buf, Header& h) {
  h = Header();
  std::vector<Field>& fields = h.fields;
  size_t pos = 0, width = 0, height = 0, points = 0; bool have_points = false;
  while (pos < buf.size()) {
    size_t e = pos; while (e < buf.size() && buf[e] != '\n') e++;
    std::string line((const char*)&buf[pos], e - pos); pos = e + 1;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line); std::string key; ss >> key;
    if (key == "FIELDS" || key == "COLUMNS") { std::string n; while (ss >> n) { Field f; f.name = n; fields.push_back(f); } }
    else if (key == "SIZE") { for (auto& f : fields) ss >> f.size; }
    else if (key == "TYPE") { for (auto& f : fields) ss >> f.type; }
    else if (key == "COUNT") { for (auto& f : fields) ss >> f.count; }
    else if (key == "WIDTH") ss >> width;
    else if (key == "HEIGHT") ss >> height;
    else if (key == "POINTS") { ss >> points; have_points = true; }
    else if (key == "DATA") { ss >> h.data_kind; break; }
  }
  if (fields.empty() || h.data_kind.empty()) return false;
  h.n = have_points ? points : width * height;
  int rec = 0; for (auto& f : fields) { f.offset = rec; rec += f.size * f.count; }
  h.rec = rec; h.payload_pos = pos;
  return true;
}

// Interleaved-record view of a DATA binary payload for the GPU de-interleave (bevgen_process_packed_host): possible
// when every PointXYZIRCT field that is present has the point type's own scalar type (BatchMultiBevGen.h:56-66), so
// no value conversion is needed; any field order / extra fields / padding.  off: x y z intensity row col t label.
struct PackedLayout {
  int stride = 0; int off[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  bool operator==(const PackedLayout& o) const { return stride == o.stride && !memcmp(off, o.off, sizeof off); }
};
inline bool packed_layout(const Header& h, PackedLayout& L) {
  static const int cs[8] = {4, 4, 4, 4, 2, 2, 4, 2}; static const char ct[8] = {'F', 'F', 'F', 'F', 'U', 'U', 'U', 'I'};
  L = PackedLayout();
  if (h.data_kind != "binary" || h.rec < 1 || h.rec > 256) return false;
  L.stride = h.rec;
  for (auto& f : h.fields) {
    int w = which_field(f.name);
    if (w < 0) continue;
    if (f.size != cs[w] || f.type != ct[w] || f.count != 1 || L.off[w] >= 0) return false;
    L.off[w] = f.offset;
  }
  return true;
}
// Field values of record i of an interleaved payload (absent fields read as 0).
inline void packed_get(const uint8_t* payload, const PackedLayout& L, size_t i, float& x, float& y, float& z, float& inten,
                       uint16_t& row, uint16_t& col, uint32_t& t, int16_t& label) {
  const uint8_t* p = payload + i * (size_t)L.stride;
  x = L.off[0] >= 0 ? rd<float>(p + L.off[0]) : 0.f; y = L.off[1] >= 0 ? rd<float>(p + L.off[1]) : 0.f;
  z = L.off[2] >= 0 ? rd<float>(p + L.off[2]) : 0.f; inten = L.off[3] >= 0 ? rd<float>(p + L.off[3]) : 0.f;
  row = L.off[4] >= 0 ? rd<uint16_t>(p + L.off[4]) : (uint16_t)0; col = L.off[5] >= 0 ? rd<uint16_t>(p + L.off[5]) : (uint16_t)0;
  t = L.off[6] >= 0 ? rd<uint32_t>(p + L.off[6]) : 0u; label = L.off[7] >= 0 ? rd<int16_t>(p + L.off[7]) : (int16_t)0;
}

inline bool decode(const std::vector<uint8_t>& buf, const Header& h, const std::string& path, Cloud& c, std::string* err = nullptr);

// Returns false (and sets err) if the file cannot be read; the reference ignores loadPCDFile's status and carries
// on with an empty cloud (BatchMultiBevGen.cpp:730), which callers reproduce by using the empty `c`.
inline bool load(const std::string& path, Cloud& c, std::string* err = nullptr) {
  c.resize(0);
  std::vector<uint8_t> buf;
  if (!read_file(path, buf, err)) return false;
  Header h;
  if (!parse_header(buf, h)) { if (err) *err = "bad PCD header in " + path; return false; }
  return decode(buf, h, path, c, err);
}

inline bool decode(const std::vector<uint8_t>& buf, const Header& h, const std::string& path, Cloud& c, std::string* err) {
  const std::vector<Field>& fields = h.fields;
  const std::string& data_kind = h.data_kind;
  size_t n = h.n; const int rec = h.rec; const size_t pos = h.payload_pos;
  c.resize(n);   // value-initialised: missing fields stay zero

  if (data_kind == "ascii") {
    const char* p = (const char*)buf.data() + pos; const char* end = (const char*)buf.data() + buf.size();
    for (size_t i = 0; i < n; i++) {
      for (auto& f : fields) {
        int w = which_field(f.name);
        for (int k = 0; k < f.count; k++) {
          while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
          if (p >= end) { c.resize(i); return true; }
          char* q; double v = strtod(p, &q); if (q == p) { q = (char*)p; while (q < end && *q != ' ' && *q != '\n') q++; }   // "nan" etc. handled by strtod
          p = q;
          if (w >= 0 && k == 0) store(c, i, w, v);
        }
      }
    }
    return true;
  }
  const uint8_t* payload = buf.data() + pos; size_t avail = buf.size() - pos;
  std::vector<uint8_t> unpacked;
  bool field_major = false;
  if (data_kind == "binary_compressed") {
    if (avail < 8) { if (err) *err = "truncated compressed PCD " + path; c.resize(0); return false; }
    uint32_t csz = rd<uint32_t>(payload), usz = rd<uint32_t>(payload + 4);
    if (avail < 8 + (size_t)csz) { if (err) *err = "truncated compressed PCD " + path; c.resize(0); return false; }
    unpacked.resize(usz);
    if (usz && !lzf_decompress(payload + 8, csz, unpacked.data(), usz)) { if (err) *err = "LZF error in " + path; c.resize(0); return false; }
    payload = unpacked.data(); avail = usz; field_major = true;
  } else if (data_kind != "binary") { if (err) *err = "unknown DATA kind in " + path; c.resize(0); return false; }
  if (avail < n * (size_t)rec) n = avail / (size_t)(rec ? rec : 1), c.resize(n);
  // fast path: the layout savePCDFileBinary writes for PointXYZIRCT
  bool canon = !field_major && rec == 26 && fields.size() == 8;
  static const char* cn[8] = {"x", "y", "z", "intensity", "row", "col", "t", "label"};
  static const int cs[8] = {4, 4, 4, 4, 2, 2, 4, 2}; static const char ct[8] = {'F', 'F', 'F', 'F', 'U', 'U', 'U', 'I'};
  for (int i = 0; canon && i < 8; i++) canon = fields[i].name == cn[i] && fields[i].size == cs[i] && fields[i].type == ct[i] && fields[i].count == 1;
  if (canon) {
    for (size_t i = 0; i < n; i++) {
      const uint8_t* p = payload + i * 26;
      c.x[i] = rd<float>(p); c.y[i] = rd<float>(p + 4); c.z[i] = rd<float>(p + 8); c.intensity[i] = rd<float>(p + 12);
      c.row[i] = rd<uint16_t>(p + 16); c.col[i] = rd<uint16_t>(p + 18); c.t[i] = rd<uint32_t>(p + 20); c.label[i] = rd<int16_t>(p + 24);
    }
    return true;
  }
  for (auto& f : fields) {
    int w = which_field(f.name);
    if (w < 0) continue;
    for (size_t i = 0; i < n; i++) {
      const uint8_t* p = field_major ? payload + (size_t)f.offset * n + i * (size_t)(f.size * f.count) : payload + i * (size_t)rec + f.offset;
      // same-type fast stores keep integer/float bit patterns exact; other types go through double like PCL's cast
      if (f.type == 'F' && f.size == 4 && w <= 3) { float v = rd<float>(p); (w == 0 ? c.x : w == 1 ? c.y : w == 2 ? c.z : c.intensity)[i] = v; }
      else store(c, i, w, field_value(p, f));
    }
  }
  return true;
}

inline std::string header(size_t n) {
  std::ostringstream o;
  o << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity row col t label\n"
       "SIZE 4 4 4 4 2 2 4 2\nTYPE F F F F U U U I\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH " << n << "\nHEIGHT 1\n"
       "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  return o.str();
}

// Serialises n packed records into `out` (header + n*26 bytes).
inline void pack_record(uint8_t* p, float x, float y, float z, float inten, uint16_t row, uint16_t col, uint32_t t, int16_t label) {
  memcpy(p, &x, 4); memcpy(p + 4, &y, 4); memcpy(p + 8, &z, 4); memcpy(p + 12, &inten, 4);
  memcpy(p + 16, &row, 2); memcpy(p + 18, &col, 2); memcpy(p + 20, &t, 4); memcpy(p + 24, &label, 2);
}

inline bool write_file(const std::string& path, const std::vector<uint8_t>& bytes) {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  bool ok = bytes.empty() || fwrite(bytes.data(), 1, bytes.size(), fp) == bytes.size();
  fclose(fp);
  return ok;
}

inline bool save_binary(const std::string& path, const Cloud& c) {
  std::string h = header(c.size());
  std::vector<uint8_t> out(h.size() + c.size() * 26);
  memcpy(out.data(), h.data(), h.size());
  for (size_t i = 0; i < c.size(); i++)
    pack_record(out.data() + h.size() + i * 26, c.x[i], c.y[i], c.z[i], c.intensity[i], c.row[i], c.col[i], c.t[i], c.label[i]);
  return write_file(path, out);
}

}  // namespace pcdio

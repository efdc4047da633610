// Stand-in for <pcl/io/pcd_io.h>: loadPCDFile / savePCDFileBinary for any point type registered with
// POINT_CLOUD_REGISTER_POINT_STRUCT.  PCD v0.7 as PCL 1.10 reads and writes it, restated by us (third-party behaviour,
// NOT pinned by the reference — ../README.md):
//   reader  DATA ascii | binary | binary_compressed (LZF, field-major); a file field feeds a point member when name,
//           TYPE, SIZE and COUNT all match (pcl::FieldMatches), anything else is skipped and the member stays zero;
//           POINTS wins over WIDTH*HEIGHT.
//   writer  PCDWriter::generateHeader<PointT> + writeBinary: registered fields in registration order, packed (padding
//           squeezed out), WIDTH/HEIGHT of the cloud, VIEWPOINT 0 0 0 1 0 0 0.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

namespace pcl { namespace io {

namespace stub_detail {
struct FileField { std::string name; int size = 4; char type = 'F'; int count = 1; std::size_t offset = 0; };

inline bool slurp(const std::string& path, std::vector<unsigned char>& buf) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) return false;
  unsigned char tmp[1 << 16]; std::size_t k;
  while ((k = std::fread(tmp, 1, sizeof tmp, fp)) > 0) buf.insert(buf.end(), tmp, tmp + k);
  std::fclose(fp);
  return true;
}

// liblzf stream: control byte < 32 = literal run, otherwise (length, distance) back reference
inline bool unlzf(const unsigned char* src, std::size_t n_src, unsigned char* dst, std::size_t n_dst) {
  std::size_t i = 0, o = 0;
  while (i < n_src) {
    unsigned c = src[i++];
    if (c < 32) {
      std::size_t run = c + 1;
      if (i + run > n_src || o + run > n_dst) return false;
      std::memcpy(dst + o, src + i, run); i += run; o += run;
    } else {
      std::size_t len = c >> 5;
      if (len == 7) { if (i >= n_src) return false; len += src[i++]; }
      if (i >= n_src) return false;
      std::size_t dist = ((c & 31u) << 8) + src[i++] + 1;
      len += 2;
      if (dist > o || o + len > n_dst) return false;
      for (std::size_t k = 0; k < len; k++, o++) dst[o] = dst[o - dist];
    }
  }
  return o == n_dst;
}
}  // namespace stub_detail

template <class PointT>
int loadPCDFile(const std::string& file_name, pcl::PointCloud<PointT>& cloud) {
  using namespace stub_detail;
  cloud.clear();
  std::vector<unsigned char> buf;
  if (!slurp(file_name, buf)) { std::cerr << "[pcl::PCDReader::read] Could not find file '" << file_name << "'.\n"; return -1; }
  std::vector<FileField> ff;
  std::size_t pos = 0, width = 0, height = 0, points = 0; bool have_points = false; std::string kind;
  while (pos < buf.size() && kind.empty()) {
    std::size_t e = pos; while (e < buf.size() && buf[e] != '\n') e++;
    std::string line(reinterpret_cast<const char*>(buf.data()) + pos, e - pos); pos = e + 1;
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    std::istringstream ss(line); std::string key;
    if (!(ss >> key) || key[0] == '#') continue;
    if (key == "FIELDS" || key == "COLUMNS") { std::string n; while (ss >> n) { FileField f; f.name = n; ff.push_back(f); } }
    else if (key == "SIZE") { for (std::size_t i = 0; i < ff.size(); i++) ss >> ff[i].size; }
    else if (key == "TYPE") { for (std::size_t i = 0; i < ff.size(); i++) ss >> ff[i].type; }
    else if (key == "COUNT") { for (std::size_t i = 0; i < ff.size(); i++) ss >> ff[i].count; }
    else if (key == "WIDTH") ss >> width;
    else if (key == "HEIGHT") ss >> height;
    else if (key == "POINTS") { ss >> points; have_points = true; }
    else if (key == "DATA") ss >> kind;
  }
  if (ff.empty() || kind.empty()) { std::cerr << "[pcl::PCDReader::readHeader] No points to read\n"; return -1; }
  std::size_t rec = 0; for (std::size_t i = 0; i < ff.size(); i++) { ff[i].offset = rec; rec += static_cast<std::size_t>(ff[i].size) * ff[i].count; }
  std::size_t n = have_points ? points : width * height;
  const std::vector<pcl::stub::Field> pf = pcl::traits::fieldList<PointT>::get();
  // member index for each file field, or -1 (pcl::FieldMatches: same name, datatype and count)
  std::vector<int> map(ff.size(), -1);
  for (std::size_t i = 0; i < ff.size(); i++)
    for (std::size_t j = 0; j < pf.size(); j++)
      if (ff[i].name == pf[j].name && ff[i].type == pf[j].type && static_cast<std::size_t>(ff[i].size) == pf[j].size && ff[i].count == 1) map[i] = static_cast<int>(j);
  for (std::size_t j = 0; j < pf.size(); j++) {
    bool found = false; for (std::size_t i = 0; i < ff.size(); i++) found = found || map[i] == static_cast<int>(j);
    if (!found) std::cerr << "Failed to find match for field '" << pf[j].name << "'.\n";
  }

  if (kind == "ascii") {
    cloud.points.resize(n);
    const char* p = reinterpret_cast<const char*>(buf.data()) + pos; const char* end = reinterpret_cast<const char*>(buf.data()) + buf.size();
    std::size_t i = 0;
    for (; i < n; i++) {
      bool short_line = false;
      for (std::size_t f = 0; f < ff.size() && !short_line; f++)
        for (int k = 0; k < ff[f].count; k++) {
          while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
          if (p >= end) { short_line = true; break; }
          const char* q = p; while (q < end && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r') q++;
          std::string tok(p, q - p); p = q;
          if (map[f] < 0) continue;
          unsigned char* dst = reinterpret_cast<unsigned char*>(&cloud.points[i]) + pf[map[f]].offset;
          if (ff[f].type == 'F') { if (ff[f].size == 4) { float v = std::strtof(tok.c_str(), nullptr); std::memcpy(dst, &v, 4); } else { double v = std::strtod(tok.c_str(), nullptr); std::memcpy(dst, &v, 8); } }
          else if (ff[f].type == 'U') { unsigned long long v = std::strtoull(tok.c_str(), nullptr, 10); std::memcpy(dst, &v, ff[f].size); }   // little endian
          else { long long v = std::strtoll(tok.c_str(), nullptr, 10); std::memcpy(dst, &v, ff[f].size); }
        }
      if (short_line) break;
    }
    cloud.points.resize(i);
  } else {
    const unsigned char* payload = buf.data() + pos; std::size_t avail = buf.size() - pos;
    std::vector<unsigned char> raw; bool field_major = false;
    if (kind == "binary_compressed") {
      if (avail < 8) return -1;
      std::uint32_t csz, usz; std::memcpy(&csz, payload, 4); std::memcpy(&usz, payload + 4, 4);
      if (avail < 8 + static_cast<std::size_t>(csz)) return -1;
      raw.resize(usz);
      if (usz && !unlzf(payload + 8, csz, raw.data(), usz)) { std::cerr << "[pcl::PCDReader::read] LZF decompression error\n"; return -1; }
      payload = raw.data(); avail = usz; field_major = true;
    } else if (kind != "binary") return -1;
    if (rec && avail < n * rec) n = avail / rec;
    cloud.points.resize(n);
    for (std::size_t f = 0; f < ff.size(); f++) {
      if (map[f] < 0) continue;
      const pcl::stub::Field& m = pf[map[f]];
      for (std::size_t i = 0; i < n; i++) {
        const unsigned char* src = field_major ? payload + ff[f].offset * n + i * m.size : payload + i * rec + ff[f].offset;
        std::memcpy(reinterpret_cast<unsigned char*>(&cloud.points[i]) + m.offset, src, m.size);
      }
    }
  }
  cloud.width = static_cast<std::uint32_t>(cloud.points.size()); cloud.height = 1;
  if (width * height == cloud.points.size() && height > 0) { cloud.width = static_cast<std::uint32_t>(width); cloud.height = static_cast<std::uint32_t>(height); }
  return 0;
}

template <class PointT>
int savePCDFileBinary(const std::string& file_name, const pcl::PointCloud<PointT>& cloud) {
  const std::vector<pcl::stub::Field> pf = pcl::traits::fieldList<PointT>::get();
  std::ostringstream h;
  h << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS";
  for (std::size_t j = 0; j < pf.size(); j++) h << " " << pf[j].name;
  h << "\nSIZE"; for (std::size_t j = 0; j < pf.size(); j++) h << " " << pf[j].size;
  h << "\nTYPE"; for (std::size_t j = 0; j < pf.size(); j++) h << " " << pf[j].type;
  h << "\nCOUNT"; for (std::size_t j = 0; j < pf.size(); j++) h << " 1";
  h << "\nWIDTH " << cloud.width << "\nHEIGHT " << cloud.height << "\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << cloud.points.size() << "\nDATA binary\n";
  std::size_t rec = 0; for (std::size_t j = 0; j < pf.size(); j++) rec += pf[j].size;
  const std::string hs = h.str();
  std::vector<unsigned char> out(hs.size() + rec * cloud.points.size());
  std::memcpy(out.data(), hs.data(), hs.size());
  unsigned char* o = out.data() + hs.size();
  for (std::size_t i = 0; i < cloud.points.size(); i++)
    for (std::size_t j = 0; j < pf.size(); j++) { std::memcpy(o, reinterpret_cast<const unsigned char*>(&cloud.points[i]) + pf[j].offset, pf[j].size); o += pf[j].size; }
  FILE* fp = std::fopen(file_name.c_str(), "wb");
  if (!fp) { std::cerr << "[pcl::PCDWriter::writeBinary] Error during open!\n"; return -1; }
  bool ok = out.empty() || std::fwrite(out.data(), 1, out.size(), fp) == out.size();
  std::fclose(fp);
  return ok ? 0 : -1;
}

} }  // namespace pcl::io

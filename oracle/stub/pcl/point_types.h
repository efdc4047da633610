// Stand-in for <pcl/point_types.h>: the two macros a custom point struct needs (BatchMultiBevGen.h:43-66).
// POINT_CLOUD_REGISTER_POINT_STRUCT records (name, offset, size, PCD type letter) per field so the PCD reader/writer in
// pcl/io/pcd_io.h can map fields by name, as PCL's does.  See ../README.md.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include <Eigen/Core>

// Headers the real <pcl/...> tree pulls in transitively and the reference relies on without including them itself
// (std::ofstream, std::chrono, std::unique_ptr, std::tie, strcmp, access(), std::sort, std::for_each).
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <tuple>
#include <unistd.h>

namespace pcl { namespace stub {
// what getArray3fMap() / getNormalVector3fMap() return in PCL (an Eigen::Map over three floats): assignment copies x, y, z
struct Map3f {
  float* p;
  Map3f& operator=(const Map3f& o) { p[0] = o.p[0]; p[1] = o.p[1]; p[2] = o.p[2]; return *this; }
};
} }
#define PCL_ADD_POINT4D \
  union EIGEN_ALIGN16 { float data[4]; struct { float x; float y; float z; }; }; \
  inline ::pcl::stub::Map3f getArray3fMap() { return ::pcl::stub::Map3f{data}; }

namespace pcl {
struct Normal {
  union EIGEN_ALIGN16 { float data_n[4]; struct { float normal_x, normal_y, normal_z; }; };
  float curvature;
  Normal() : curvature(0.f) { data_n[0] = data_n[1] = data_n[2] = data_n[3] = 0.f; }
  inline stub::Map3f getNormalVector3fMap() { return stub::Map3f{data_n}; }
};
struct PointXYZ {                      // pcl/impl/point_types.hpp: x = y = z = 0, data[3] = 1
  PCL_ADD_POINT4D
  PointXYZ() { data[0] = data[1] = data[2] = 0.f; data[3] = 1.f; }
};
struct PointNormal {
  PCL_ADD_POINT4D
  union EIGEN_ALIGN16 { float data_n[4]; struct { float normal_x, normal_y, normal_z; }; };
  float curvature;
  PointNormal() : curvature(0.f) { data[0] = data[1] = data[2] = 0.f; data[3] = 1.f; data_n[0] = data_n[1] = data_n[2] = data_n[3] = 0.f; }
  inline stub::Map3f getNormalVector3fMap() { return stub::Map3f{data_n}; }
};
namespace stub {
struct Field { std::string name; std::size_t offset; std::size_t size; char type; };
template <class T> struct pcd_type;
template <> struct pcd_type<float> { static const char v = 'F'; };
template <> struct pcd_type<double> { static const char v = 'F'; };
template <> struct pcd_type<std::uint8_t> { static const char v = 'U'; };
template <> struct pcd_type<std::uint16_t> { static const char v = 'U'; };
template <> struct pcd_type<std::uint32_t> { static const char v = 'U'; };
template <> struct pcd_type<std::int8_t> { static const char v = 'I'; };
template <> struct pcd_type<std::int16_t> { static const char v = 'I'; };
template <> struct pcd_type<std::int32_t> { static const char v = 'I'; };
template <class T> inline void add_field(std::vector<Field>& v, const char* name, std::size_t off) {
  v.push_back(Field{name, off, sizeof(T), pcd_type<T>::v});
}
}  // namespace stub
namespace traits { template <class P> struct fieldList; }
}  // namespace pcl

// The field list is a Boost.PP sequence "(type, member, tag)(type, member, tag)…"; walk it with two macros that
// hand over to each other.
#define PCL_STUB_CAT_(a, b) a##b
#define PCL_STUB_CAT(a, b) PCL_STUB_CAT_(a, b)
#define PCL_STUB_F_A(type, member, tag) ::pcl::stub::add_field<type>(v, #tag, offsetof(P_, member)); PCL_STUB_F_B
#define PCL_STUB_F_B(type, member, tag) ::pcl::stub::add_field<type>(v, #tag, offsetof(P_, member)); PCL_STUB_F_A
#define PCL_STUB_F_A_END
#define PCL_STUB_F_B_END
#define POINT_CLOUD_REGISTER_POINT_STRUCT(PT, seq)                                             \
  namespace pcl { namespace traits { template <> struct fieldList<PT> {                        \
    static std::vector< ::pcl::stub::Field> get() {                                            \
      typedef PT P_; std::vector< ::pcl::stub::Field> v; PCL_STUB_CAT(PCL_STUB_F_A seq, _END)  \
      return v; } }; } }

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

#define PCL_ADD_POINT4D \
  union EIGEN_ALIGN16 { float data[4]; struct { float x; float y; float z; }; };

namespace pcl {
struct Normal { float normal_x, normal_y, normal_z, curvature; };
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

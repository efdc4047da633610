// Stand-in for <pcl/point_cloud.h>: points vector (value-initialising resize, as std::vector does and PCL relies on),
// width / height bookkeeping of PCL 1.10's PointCloud::resize, shared_ptr Ptr.  See ../README.md.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include <Eigen/Core>

// PCL 1.10's Ptr types are boost::shared_ptr and the reference names that type directly (TopPartRegistration.cpp:364)
namespace boost { using std::shared_ptr; using std::make_shared; }

namespace pcl {
template <class PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT> > Ptr;
  typedef std::shared_ptr<const PointCloud<PointT> > ConstPtr;
  std::vector<PointT, Eigen::aligned_allocator<PointT> > points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  void resize(std::size_t n) {
    points.resize(n);                      // new elements are value-initialised: all fields zero
    if (width * height != n) { width = static_cast<std::uint32_t>(n); height = 1; }
  }
  void clear() { points.clear(); width = height = 0; }
  void reserve(std::size_t n) { points.reserve(n); }
  void push_back(const PointT& p) { points.push_back(p); width = static_cast<std::uint32_t>(points.size()); height = 1; }   // PCL 1.10 point_cloud.h:548-553
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl

// Stand-in for <pcl/common/transforms.h>: transformPointCloud(in, out, Affine3f) with PCL >= 1.9's SSE2 operation
// order  m0*x + (m1*y + (m2*z + t))  per output coordinate, all other fields copied.  Third-party
// behaviour restated by us — NOT pinned by the reference.  See ../README.md.
#pragma once
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>

namespace pcl {
template <class PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Affine3f& tf) {
  if (&in != &out) { out.points = in.points; out.width = in.width; out.height = in.height; out.is_dense = in.is_dense; }
  for (std::size_t i = 0; i < in.points.size(); i++) {
    const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    float r[3];
    for (int k = 0; k < 3; k++) {
      const float a = tf(k, 2) * z, b = a + tf(k, 3);      // built with -ffp-contract=off: no FMA is formed
      const float c = tf(k, 1) * y, d = c + b;
      const float e = tf(k, 0) * x; r[k] = e + d;
    }
    out.points[i].x = r[0]; out.points[i].y = r[1]; out.points[i].z = r[2];
  }
}
}  // namespace pcl

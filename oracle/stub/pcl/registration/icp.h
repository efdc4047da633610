// Stand-in for <pcl/registration/icp.h>: the ICP objects of TopPartRegistration.cpp:183-237 (out of scope: only
// extractTopAndFlatten, :79-141, is called through oracle/_ref) as shapes that compile and never converge.  See ../../README.md.
#pragma once
#include <pcl/point_cloud.h>
#include <Eigen/Core>
namespace pcl {
template <class Src, class Tgt> struct IterativeClosestPoint {
  void setMaxCorrespondenceDistance(double) {}
  void setMaximumIterations(int) {}
  void setTransformationEpsilon(double) {}
  void setEuclideanFitnessEpsilon(double) {}
  void setInputSource(const typename PointCloud<Src>::Ptr&) {}
  void setInputTarget(const typename PointCloud<Tgt>::Ptr&) {}
  void align(PointCloud<Src>&, const Eigen::Matrix4f&) {}
  bool hasConverged() const { return false; }
  double getFitnessScore() const { return 0.0; }
  Eigen::Matrix4f getFinalTransformation() const { return Eigen::Matrix4f::Identity(); }
};
template <class Src, class Tgt> struct IterativeClosestPointWithNormals : IterativeClosestPoint<Src, Tgt> {};
}  // namespace pcl

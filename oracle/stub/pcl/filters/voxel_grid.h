// Stand-in for <pcl/filters/voxel_grid.h>: TopPartRegistration.cpp's main() down-samples with it (:283-292, :330-341); only
// extractTopAndFlatten (:79-141) is ever called through oracle/_ref, so the filter is a shape that compiles.  See ../../README.md.
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <class PointT> struct VoxelGrid {
  void setLeafSize(float, float, float) {}
  void setInputCloud(const typename PointCloud<PointT>::Ptr&) {}
  void filter(PointCloud<PointT>&) {}
};
}  // namespace pcl

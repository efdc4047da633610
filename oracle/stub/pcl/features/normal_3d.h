// Stand-in for <pcl/features/normal_3d.h>: brings in the search tree type TopPartRegistration.cpp:151-154 names (out of scope,
// compiled only).  See ../../README.md.
#pragma once
#include <memory>
#include <pcl/point_cloud.h>
namespace pcl { namespace search {
template <class PointT> struct KdTree { typedef std::shared_ptr<KdTree<PointT> > Ptr; };
} }

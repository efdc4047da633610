// Stand-in for <pcl/io/io.h>: concatenateFields as TopPartRegistration.cpp:161 calls it (out of scope, compiled only); the hot
// path uses nothing from this header.  See oracle/stub/README.md.
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <class A, class B, class C> void concatenateFields(const PointCloud<A>&, const PointCloud<B>&, PointCloud<C>&) {}
}

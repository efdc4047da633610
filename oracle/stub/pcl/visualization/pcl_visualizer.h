// Stand-in for <pcl/visualization/pcl_visualizer.h>.  The hot path uses nothing of the viewer, but the real header
// matters to the reference in two indirect ways that are reproduced here:
//   * VTK's vtkIOStream.h exports std::ios into the global namespace — BatchMultiBevGen.cpp:384 writes bare `ios::in`, BatchCloudManip.cpp:324 bare `endl`;
//   * VTK <= 8's vtkSetGet.h includes <math.h>, libstdc++'s wrapper of which exports the float overloads of
//     atan2 / sqrt / round into the global namespace, which decides how BatchMultiBevGen.cpp:173 binds
//     (-DSTUB_NO_MATH_H builds the other possibility).  See ../README.md.
#pragma once
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <pcl/point_cloud.h>
#ifndef STUB_NO_MATH_H
#include <math.h>
#else
#include <cmath>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#endif
// vtkIOStream.h's export list (BatchCloudManip.cpp:324 writes bare `endl`, MulranPointCloudSelect.cpp:105-106 bare `ifstream`)
using std::cerr; using std::cin; using std::cout; using std::endl; using std::ends; using std::ios; using std::istream; using std::ostream;
using std::fstream; using std::ifstream; using std::ofstream; using std::istringstream; using std::ostringstream; using std::stringstream;
using std::dec; using std::hex; using std::setfill; using std::setprecision; using std::setw;

namespace pcl { namespace visualization {
enum RenderingProperties { PCL_VISUALIZER_POINT_SIZE = 0, PCL_VISUALIZER_COLOR = 4 };
template <class PointT> struct PointCloudColorHandlerCustom {
  PointCloudColorHandlerCustom(const typename PointCloud<PointT>::Ptr&, double, double, double) {}
};
// A viewer that is already closed: CloudManip.cpp:155's loop ends at once (the interactive window is out of scope).
struct PCLVisualizer {
  explicit PCLVisualizer(const std::string&) {}
  template <class PointT, class H> bool addPointCloud(const typename PointCloud<PointT>::Ptr&, const H&, const std::string&) { return true; }
  template <class C, class H> bool addPointCloud(const C&, const H&, const std::string&) { return true; }
  void addCoordinateSystem(double, const std::string&, int) {}
  void setBackgroundColor(double, double, double, int = 0) {}
  bool setPointCloudRenderingProperties(int, double, const std::string&) { return true; }
  // TopPartRegistration.cpp:365-374 (its viewer is out of scope as well: compiled, never run)
  template <class PointT> bool addPointCloud(const typename PointCloud<PointT>::Ptr&, const std::string&) { return true; }
  bool setPointCloudRenderingProperties(int, double, double, double, const std::string&) { return true; }
  template <class PointT, class PointNT> bool addPointCloudNormals(const typename PointCloud<PointT>::Ptr&, const typename PointCloud<PointNT>::Ptr&,
                                                                    int, float, const std::string&) { return true; }
  bool wasStopped() const { return true; }
  void spinOnce(int = 1) {}
};
} }

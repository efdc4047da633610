// ref_top_part_shim.cpp — builds oracle/_ref/libtoppart_ref.so from the reference's own TopPartRegistration.cpp, compiled
// unmodified where it lies against oracle/stub (TEST INFRASTRUCTURE ONLY; see oracle/stub/README.md).
//   ref_extract_top_and_flatten   extractTopAndFlatten (TopPartRegistration.cpp:79-141): per 20 m cell the round(0.2 * count)
//                                 highest non-ground points, z dropped — SURVEY 8(f)-4, what bevgen_top_flatten replaces
// Everything else in that translation unit (normals, the two ICP stages, the viewer, main) is out of scope; it is compiled
// against shapes that never run.  include/Normal2dEstimation.h (a PCL feature class of the reference's own) is kept out by
// its include guard and replaced by a shape with the four members addNormal (:144-180) names.
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/features/normal_3d.h>
#define _NORMAL2DESTIMATION_H
struct Normal2dEstimation {
  void setInputCloud(const pcl::PointCloud<pcl::PointXYZ>::Ptr&) {}
  void setSearchMethod(const pcl::search::KdTree<pcl::PointXYZ>::Ptr&) {}
  void setRadiusSearch(double) {}
  void compute(const pcl::PointCloud<pcl::Normal>::Ptr&) {}
};
#define main ref_top_part_main_impl
#include "TopPartRegistration.cpp"   // found through -I/root/reference
#undef main
#include <cstdint>

#define REF_API extern "C" __attribute__((visibility("default")))

// n input points (x, y, z, label; the other fields play no part) -> the flattened cloud's x, y (z is 0 by construction, :137);
// returns the number of output points, -1 if a z of the output is not 0
REF_API int64_t ref_extract_top_and_flatten(int64_t n, const float* x, const float* y, const float* z, const int16_t* label,
                                            float* out_x, float* out_y, int64_t cap) {
  pcl::PointCloud<pcl::PointXYZIRCT>::Ptr in(new pcl::PointCloud<pcl::PointXYZIRCT>());
  in->points.resize(n);
  for (int64_t i = 0; i < n; i++) { auto& p = in->points[i]; p.x = x[i]; p.y = y[i]; p.z = z[i]; p.label = label[i]; }
  pcl::PointCloud<pcl::PointXYZ>::Ptr out(new pcl::PointCloud<pcl::PointXYZ>());
  extractTopAndFlatten(in, out);
  const int64_t m = (int64_t)out->points.size();
  for (int64_t i = 0; i < m && i < cap; i++) {
    if (out->points[i].z != 0.0f) return -1;
    out_x[i] = out->points[i].x; out_y[i] = out->points[i].y;
  }
  return m;
}

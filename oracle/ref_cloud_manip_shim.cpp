// ref_cloud_manip_shim.cpp — builds oracle/_ref/libcloudmanip_ref.so from the reference's own CloudManip.cpp, compiled
// unmodified where it lies against oracle/stub (TEST INFRASTRUCTURE ONLY; see oracle/stub/README.md).
//   ref_save_as_mat        saveAsMat (CloudManip.cpp:79-109): the 201x201 float grid the reference hands to imwrite,
//                          and the CSV text it writes
//   ref_cloud_manip_main   the tool's main() (:111-161) with the viewer stubbed out as already closed
//   ref_cloud_manip_matrix the Affine3f main() builds (:119-126) as the stub's Eigen subset evaluates it (third-party
//                          arithmetic restated by us) + pcl::transformPointCloud of the stub on n points
#define main ref_cloud_manip_main_impl
#include "CloudManip.cpp"   // found through -I/root/reference
#undef main
#include <cstdint>
#include <cstring>

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int ref_save_as_mat(int64_t n, const float* x, const float* y, const float* z, float interval, float* grid_out,
                            const char* csv_path) {
  pcl::PointCloud<PointType>::Ptr cloud(new pcl::PointCloud<PointType>());
  cloud->points.resize(n);
  for (int64_t i = 0; i < n; i++) { cloud->points[i].x = x[i]; cloud->points[i].y = y[i]; cloud->points[i].z = z[i]; }
  cv::Mat got;
  cv::stub::imwrite_hook() = [&got](const std::string&, const cv::Mat& m) { got = m.clone(); return true; };
  saveAsMat(cloud, std::string(csv_path), interval);
  cv::stub::imwrite_hook() = nullptr;
  if (got.empty() || got.type() != CV_32F) return -1;
  for (int r = 0; r < got.rows; r++) std::memcpy(grid_out + (size_t)r * got.cols, got.ptr(r), sizeof(float) * got.cols);
  return got.rows;
}

REF_API int ref_cloud_manip_main(int argc, char** argv) { return ref_cloud_manip_main_impl(argc, argv); }

REF_API void ref_cloud_manip_matrix(float tx, float ty, float tz, float theta_deg, float* rt12, int64_t n, const float* x,
                                    const float* y, const float* z, float* ox, float* oy, float* oz) {
  Eigen::Affine3f transform = Eigen::Affine3f::Identity();
  transform.translation() << tx, ty, tz;
  float theta = theta_deg / 180.0f * M_PI;
  transform.rotate(Eigen::AngleAxisf(theta, Eigen::Vector3f::UnitZ()));
  for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) rt12[4 * i + j] = transform(i, j);
  pcl::PointCloud<PointType> in, out;
  in.points.resize(n);
  for (int64_t i = 0; i < n; i++) { in.points[i].x = x[i]; in.points[i].y = y[i]; in.points[i].z = z[i]; }
  pcl::transformPointCloud(in, out, transform);
  for (int64_t i = 0; i < n; i++) { ox[i] = out.points[i].x; oy[i] = out.points[i].y; oz[i] = out.points[i].z; }
}

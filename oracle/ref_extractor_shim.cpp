// ref_extractor_shim.cpp — builds oracle/_ref/lib{mulran,oxford,kitti}select_ref.so (and the _dbl variants) from the
// reference's own MulranPointCloudSelect.cpp / OxfordPointCloudSelect.cpp / KittiPointCloudSelect.cpp, each compiled unmodified
// where it lies against oracle/stub (TEST INFRASTRUCTURE ONLY; see oracle/stub/README.md).  One of -DREF_MULRAN, -DREF_OXFORD,
// -DREF_KITTI selects the translation unit (the three define the same global names).
//   ref_extract_point_cloud   initDirectories(root) + extractPointCloud(timestamp): reads the scan file the extractor would read
//                             (MulRan sensor_data/Ouster/%010ld.bin, Oxford velodyne_left/%010ld.bin, KITTI velodyne/%06d.bin) and
//                             returns the cloud it builds, fields as they stand in the returned points — SURVEY 8(f)-2, what
//                             bevgen_project replaces (Mulran :95-130, Oxford :146-224, Kitti :156-246)
// Everything else in those translation units (pose files, interpolation, keyframe selection, main) is out of scope: compiled,
// never run.
#define main ref_extractor_main_impl
#if defined(REF_MULRAN)
#include "MulranPointCloudSelect.cpp"   // found through -I/root/reference
#elif defined(REF_OXFORD)
#include "OxfordPointCloudSelect.cpp"
#elif defined(REF_KITTI)
#include "KittiPointCloudSelect.cpp"
#else
#error "one of REF_MULRAN / REF_OXFORD / REF_KITTI"
#endif
#undef main
#include <cstdint>

#define REF_API extern "C" __attribute__((visibility("default")))

// returns the number of points of the cloud extractPointCloud returned; the first min(n, cap) are copied out
REF_API int64_t ref_extract_point_cloud(const char* root_dir, int64_t timestamp, int64_t cap, float* x, float* y, float* z,
                                        float* intensity, uint16_t* row, uint16_t* col, int16_t* label) {
  initDirectories(std::string(root_dir));
  pcl::PointCloud<pcl::PointXYZIRCT>::Ptr cloud = extractPointCloud(timestamp);
  const int64_t n = (int64_t)cloud->points.size();
  for (int64_t i = 0; i < n && i < cap; i++) {
    const pcl::PointXYZIRCT& p = cloud->points[i];
    x[i] = p.x; y[i] = p.y; z[i] = p.z; intensity[i] = p.intensity; row[i] = p.row; col[i] = p.col; label[i] = p.label;
  }
  return n;
}

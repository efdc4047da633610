// ref_batch_cloud_manip_shim.cpp — builds oracle/_ref/libbatchcloudmanip_ref.so from the reference's own
// BatchCloudManip.cpp (the HDL-64E-only predecessor tool), compiled unmodified against oracle/stub.
// TEST INFRASTRUCTURE ONLY (oracle/stub/README.md).
//   ref_bcm_frame   getOrderedCloud (:47-62, no bounds check) -> markGroundPoints (:64-199) -> saveAsMat (:201-239):
//                   labels [S] and the 201x201 float bird-view map the reference hands to imwrite
//   ref_bcm_main    the tool's main() (:269-335)
#define main ref_bcm_main_impl
#include "BatchCloudManip.cpp"   // found through -I/root/reference
#undef main
#include <cstdint>
#include <cstring>

#define REF_API extern "C" __attribute__((visibility("default")))
static bool g_nb = false;

REF_API int ref_bcm_frame(int64_t n, const float* x, const float* y, const float* z, const float* intensity, const uint16_t* row,
                          const uint16_t* col, const int16_t* label, int16_t* label_out, float* bvm_out, const char* out_prefix) {
  if (!g_nb) { setNeighbors(); g_nb = true; }
  pcl::PointCloud<pcl::PointXYZIRCT>::Ptr in(new pcl::PointCloud<pcl::PointXYZIRCT>()), ord(new pcl::PointCloud<pcl::PointXYZIRCT>());
  in->points.resize(n);
  for (int64_t i = 0; i < n; i++) {
    pcl::PointXYZIRCT& p = in->points[i];
    p.x = x[i]; p.y = y[i]; p.z = z[i]; p.intensity = intensity[i]; p.row = row[i]; p.col = col[i]; p.label = label[i];
  }
  cv::Mat ground_mat, got;
  getOrderedCloud(in, ord);
  markGroundPoints(ord, ground_mat);
  cv::stub::imwrite_hook() = [&got](const std::string&, const cv::Mat& m) { got = m.clone(); return true; };
  saveAsMat(ord, std::string(out_prefix), 1.0f);
  cv::stub::imwrite_hook() = nullptr;
  const int64_t S = (int64_t)ord->points.size();
  for (int64_t s = 0; s < S; s++) label_out[s] = ord->points[s].label;
  if (got.empty() || got.type() != CV_32F) return -1;
  for (int r = 0; r < got.rows; r++) std::memcpy(bvm_out + (size_t)r * got.cols, got.ptr(r), sizeof(float) * got.cols);
  return (int)S;
}

REF_API int ref_bcm_main(int argc, char** argv) { four_neighbor_iterator_.clear(); g_nb = true; return ref_bcm_main_impl(argc, argv); }

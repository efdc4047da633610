// ref_bevgen_shim.cpp — builds oracle/_ref/libbevgen_ref.so from the reference's OWN translation unit,
// /root/reference/BatchMultiBevGen.cpp (+ src/Utility.cpp), compiled unmodified where it lies against the stand-in
// headers in oracle/stub/ (see oracle/stub/README.md for what that pins and what it does not).
// TEST INFRASTRUCTURE ONLY: nothing from the reference is copied into this repository — the #include below pulls the
// file from /root/reference at build time, and the resulting .so is git-ignored (it travels to the GPU box prebuilt).
//
// Exposed entry points (plain C, for ctypes):
//   ref_set_sensor      parseSensorType + getSensorParams -> the TU's global sensor_params_   (BatchMultiBevGen.cpp:718-719)
//   ref_frame           getOrderedCloud -> markGroundPoints -> computeAndSaveMultiBev -> computeAndSaveSingleBev on one
//                       frame (:735-747), returning the ordered cloud, ground_mat, labels and the images / files the
//                       reference's own code produced
//   ref_labels          selectMajorFrames + getKeyFrameLabel (:502-636)
//   ref_read_poses      readKeyframePose (:381-460)
//   ref_list_pcd        getPcdFileNames (:469-494)
//   ref_save_labels     saveLabels (:645-661)
//   ref_bevgen_main     the reference's main() (:664-771), whole program
//   ref_math_overloads  which overloads the unqualified atan2 / sqrt / abs calls of :173,:179 bind to in this build
#define main ref_bevgen_main_impl
#include "BatchMultiBevGen.cpp"   // found through -I/root/reference
#undef main

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdlib.h>
#include <sys/stat.h>
#include <type_traits>

#define REF_API extern "C" __attribute__((visibility("default")))

static bool g_neighbors_set = false;
static void ensure_neighbors() { if (!g_neighbors_set) { setNeighbors(); g_neighbors_set = true; } }

REF_API int ref_set_sensor(const char* name, int32_t* out4 /* N_SCAN, Horizon_SCAN, GROUND_UPPER_SCAN, HEIGHT_RES bits */) {
  SensorType t = parseSensorType(std::string(name));
  if (t == SensorType::UNKNOWN) return -1;
  sensor_params_ = getSensorParams(t);
  ensure_neighbors();
  if (out4) {
    out4[0] = sensor_params_.N_SCAN; out4[1] = sensor_params_.Horizon_SCAN; out4[2] = sensor_params_.GROUND_UPPER_SCAN;
    std::memcpy(&out4[3], &sensor_params_.HEIGHT_RES, 4);
  }
  return 0;
}

// 0 = float overload, 1 = double, as seen from the same global scope the reference's calls are written in
REF_API int ref_math_overloads(int* atan2_is_double, int* sqrt_is_double, int* abs_is_double, int* round_is_double) {
  float f = 1.0f;
  *atan2_is_double = std::is_same<decltype(atan2(f, f)), double>::value;
  *sqrt_is_double = std::is_same<decltype(sqrt(f)), double>::value;
  *abs_is_double = std::is_same<decltype(abs(f)), float>::value ? 0 : (std::is_same<decltype(abs(f)), double>::value ? 1 : 2 /* int! */);
  *round_is_double = std::is_same<decltype(round(f)), double>::value;
  return 0;
}

static std::string g_tmp_dir;
static const std::string& tmp_dir() {
  if (g_tmp_dir.empty()) {
    char tpl[] = "/tmp/bevgen_ref_XXXXXX";
    const char* d = mkdtemp(tpl);
    g_tmp_dir = std::string(d ? d : "/tmp") + "/";
  }
  return g_tmp_dir;
}

static bool slurp(const std::string& p, std::vector<unsigned char>& b) {
  FILE* fp = std::fopen(p.c_str(), "rb");
  if (!fp) return false;
  unsigned char tmp[1 << 16]; size_t k; b.clear();
  while ((k = std::fread(tmp, 1, sizeof tmp, fp)) > 0) b.insert(b.end(), tmp, tmp + k);
  std::fclose(fp);
  return true;
}

// One frame through the reference's functions.  Inputs: SoA of n points (t may be NULL -> 0).
// Outputs (any may be NULL): ord_* [S] the ordered cloud after markGroundPoints (label_out = its labels),
// ground_mat [S] int8 as markGroundPoints leaves it, single [224*224], multi [24*224*224] = the Mats handed to
// cv::imwrite, bin_equal = 1 iff the .bin the reference wrote holds the same bytes as those 24 Mats,
// csv_out/csv_cap: the CSV text the reference wrote (truncated to csv_cap), *csv_len its full length.
REF_API int ref_frame(int64_t n, const float* x, const float* y, const float* z, const float* intensity,
                      const uint16_t* row, const uint16_t* col, const uint32_t* t, const int16_t* label,
                      float* ord_x, float* ord_y, float* ord_z, float* ord_i, uint16_t* ord_row, uint16_t* ord_col,
                      uint32_t* ord_t, int16_t* label_out, int8_t* ground_mat_out, uint8_t* single, uint8_t* multi,
                      int* bin_equal, char* csv_out, int64_t csv_cap, int64_t* csv_len) {
  ensure_neighbors();
  pcl::PointCloud<pcl::PointXYZIRCT>::Ptr cloud_unordered(new pcl::PointCloud<pcl::PointXYZIRCT>());
  pcl::PointCloud<pcl::PointXYZIRCT>::Ptr cloud_ordered(new pcl::PointCloud<pcl::PointXYZIRCT>());
  cloud_unordered->points.resize(n);
  for (int64_t i = 0; i < n; i++) {
    pcl::PointXYZIRCT& p = cloud_unordered->points[i];
    p.x = x[i]; p.y = y[i]; p.z = z[i]; p.intensity = intensity[i]; p.row = row[i]; p.col = col[i];
    p.t = t ? t[i] : 0u; p.label = label[i];
  }
  cloud_unordered->width = (uint32_t)n; cloud_unordered->height = 1;

  const std::string dir = tmp_dir();
  output_multi_bvm_bin_dir_ = dir; output_multi_bvm_img_dir_ = dir; output_single_bvm_img_dir_ = dir; output_single_bvm_csv_dir_ = dir;
  ::mkdir((dir + "f").c_str(), 0777);   // computeAndSaveMultiBev would fork `mkdir -p` for it (:303-306)

  std::vector<cv::Mat> captured;
  cv::stub::imwrite_hook() = [&captured](const std::string&, const cv::Mat& m) { captured.push_back(m.clone()); return true; };

  cv::Mat ground_mat;
  getOrderedCloud(cloud_unordered, cloud_ordered);      // :735
  markGroundPoints(cloud_ordered, ground_mat);          // :736
  computeAndSaveMultiBev(cloud_ordered, "f", 1.0f);     // :746
  computeAndSaveSingleBev(cloud_ordered, "f", 1.0f);    // :747
  cv::stub::imwrite_hook() = nullptr;

  const int64_t S = (int64_t)cloud_ordered->points.size();
  for (int64_t s = 0; s < S; s++) {
    const pcl::PointXYZIRCT& p = cloud_ordered->points[s];
    if (ord_x) ord_x[s] = p.x; if (ord_y) ord_y[s] = p.y; if (ord_z) ord_z[s] = p.z; if (ord_i) ord_i[s] = p.intensity;
    if (ord_row) ord_row[s] = p.row; if (ord_col) ord_col[s] = p.col; if (ord_t) ord_t[s] = p.t; if (label_out) label_out[s] = p.label;
  }
  if (ground_mat_out)
    for (int r = 0; r < ground_mat.rows; r++) std::memcpy(ground_mat_out + (int64_t)r * ground_mat.cols, ground_mat.ptr(r), ground_mat.cols);
  if (captured.size() != 25) return -2;
  const size_t cell = (size_t)captured[0].rows * captured[0].cols;
  std::vector<unsigned char> all(24 * cell);
  for (int l = 0; l < 24; l++) std::memcpy(all.data() + l * cell, captured[l].ptr(0), cell);
  if (multi) std::memcpy(multi, all.data(), all.size());
  if (single) std::memcpy(single, captured[24].ptr(0), cell);
  std::vector<unsigned char> f;
  if (bin_equal) *bin_equal = slurp(dir + "f.bin", f) && f == all;
  if (csv_len) {
    *csv_len = slurp(dir + "f.csv", f) ? (int64_t)f.size() : -1;
    if (csv_out && *csv_len > 0) std::memcpy(csv_out, f.data(), (size_t)std::min<int64_t>(*csv_len, csv_cap));
  }
  std::remove((dir + "f.bin").c_str()); std::remove((dir + "f.csv").c_str());
  return (int)S;
}

// Timing entry for bench.py's reference arm: the four calls of the reference's hot loop body (:735-747) over nf frames,
// `iters` times, clouds built once outside the timed span.  The PNG / CSV encoders are stubbed out (imwrite hook that
// does nothing, cv::format returning "") because the GPU arm it is compared with has no file encoders either; the
// reference's own .bin write (:307-314) goes to a tmpfs directory.  Returns the seconds spent inside the four calls.
REF_API double ref_bench(int nf, const int64_t* offs, const float* x, const float* y, const float* z, const float* intensity,
                         const uint16_t* row, const uint16_t* col, const int16_t* label, int iters, uint64_t* checksum) {
  ensure_neighbors();
  std::vector<pcl::PointCloud<pcl::PointXYZIRCT>::Ptr> clouds(nf);
  for (int f = 0; f < nf; f++) {
    clouds[f].reset(new pcl::PointCloud<pcl::PointXYZIRCT>());
    const int64_t o = offs[f], n = offs[f + 1] - o;
    clouds[f]->points.resize(n);
    for (int64_t i = 0; i < n; i++) {
      pcl::PointXYZIRCT& p = clouds[f]->points[i];
      p.x = x[o + i]; p.y = y[o + i]; p.z = z[o + i]; p.intensity = intensity[o + i]; p.row = row[o + i]; p.col = col[o + i]; p.label = label[o + i];
    }
  }
  std::string dir = tmp_dir();
  if (access("/dev/shm", W_OK) == 0) { char tpl[] = "/dev/shm/bevgen_ref_XXXXXX"; const char* d = mkdtemp(tpl); if (d) dir = std::string(d) + "/"; }
  output_multi_bvm_bin_dir_ = dir; output_multi_bvm_img_dir_ = dir; output_single_bvm_img_dir_ = dir; output_single_bvm_csv_dir_ = dir;
  ::mkdir((dir + "f").c_str(), 0777);
  uint64_t sum = 0;
  cv::stub::imwrite_hook() = [&sum](const std::string&, const cv::Mat& m) { sum += m.at<uint8_t>(112, 112); return true; };
  cv::stub::format_enabled() = false;
  const auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < iters; it++)
    for (int f = 0; f < nf; f++) {
      pcl::PointCloud<pcl::PointXYZIRCT>::Ptr cloud_ordered(new pcl::PointCloud<pcl::PointXYZIRCT>());
      cv::Mat ground_mat;
      getOrderedCloud(clouds[f], cloud_ordered);
      markGroundPoints(cloud_ordered, ground_mat);
      computeAndSaveMultiBev(cloud_ordered, "f", 1.0f);
      computeAndSaveSingleBev(cloud_ordered, "f", 1.0f);
      sum += (uint64_t)cloud_ordered->points[cloud_ordered->points.size() / 2].label & 0xFFFF;
    }
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  cv::stub::imwrite_hook() = nullptr;
  cv::stub::format_enabled() = true;
  std::remove((dir + "f.bin").c_str()); std::remove((dir + "f.csv").c_str()); ::rmdir((dir + "f").c_str());
  if (dir != tmp_dir()) ::rmdir(dir.c_str());
  if (checksum) *checksum = sum;
  return dt;
}

static std::vector<Pose6f> poses_from_xyz(int K, const float* xyz) {
  std::vector<Pose6f> v(K);
  for (int i = 0; i < K; i++) { v[i].x = xyz[3 * i]; v[i].y = xyz[3 * i + 1]; v[i].z = xyz[3 * i + 2]; v[i].roll = v[i].pitch = v[i].yaw = 0.f; }
  return v;
}

// selectMajorFrames + getKeyFrameLabel.  major_out [K] (first M valid), labels_out [K*M] dense row-major or NULL.
// Pass labels_out = NULL first to learn M.  Returns M.
REF_API int ref_labels(int K, const float* xyz, int32_t* major_out, float* labels_out) {
  std::vector<Pose6f> poses = poses_from_xyz(K, xyz);
  std::vector<int32_t> major = selectMajorFrames(poses);
  const int M = (int)major.size();
  for (int j = 0; j < M; j++) major_out[j] = major[j];
  if (labels_out) {
    std::vector<LabelType> lab = getKeyFrameLabel(poses, major);
    for (int i = 0; i < K; i++) std::memcpy(labels_out + (size_t)i * M, lab[i].data(), sizeof(float) * M);
  }
  return M;
}

// saveLabels on a dense K x M table
REF_API int ref_save_labels(int K, int M, const float* labels, const char* path) {
  std::vector<LabelType> lab(K, LabelType(M));
  for (int i = 0; i < K; i++) std::memcpy(lab[i].data(), labels + (size_t)i * M, sizeof(float) * M);
  saveLabels(lab, std::string(path));
  return 0;
}

// readKeyframePose: xyz_out [cap*3]; returns the number of poses read
REF_API int ref_read_poses(const char* path, float* xyz_out, int cap) {
  std::vector<Pose6f> poses = readKeyframePose(std::string(path));
  for (int i = 0; i < (int)poses.size() && i < cap; i++) { xyz_out[3 * i] = poses[i].x; xyz_out[3 * i + 1] = poses[i].y; xyz_out[3 * i + 2] = poses[i].z; }
  return (int)poses.size();
}

// getPcdFileNames: names joined with '\n' into out (cap bytes); returns the count
REF_API int ref_list_pcd(const char* dir, char* out, int64_t cap) {
  std::vector<std::string> names;
  getPcdFileNames(std::string(dir), names);
  std::string all;
  for (size_t i = 0; i < names.size(); i++) { all += names[i]; all += '\n'; }
  if (out && cap > 0) { size_t k = std::min<size_t>(all.size(), (size_t)cap - 1); std::memcpy(out, all.data(), k); out[k] = 0; }
  return (int)names.size();
}

REF_API float ref_get_distance(const float* a, const float* b) {
  Pose6f p, q; p.x = a[0]; p.y = a[1]; p.z = a[2]; q.x = b[0]; q.y = b[1]; q.z = b[2];
  return getDistance(p, q);
}

REF_API void ref_belonging_grid(float x, float y, int* sr, int* sc) {
  pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>());
  c->points.resize(1); c->points[0].x = x; c->points[0].y = y;
  std::pair<int, int> g = getBelongingGrid(c, 0);
  *sr = g.first; *sc = g.second;
}

// The reference's whole program.  argv as the tool takes it: {"batch_multi_bev_gen", root, sensor}.
REF_API int ref_bevgen_main(int argc, char** argv) {
  four_neighbor_iterator_.clear(); g_neighbors_set = false;   // main() calls setNeighbors() itself (:712)
  int rc = ref_bevgen_main_impl(argc, argv);
  g_neighbors_set = true;
  return rc;
}

// cloud_manip — replacement of the reference tool (CloudManip.cpp:111-141) without the interactive VTK viewer
// (:143-158, out of scope):  cloud_manip <cloud.pcd> <tx> <ty> <tz> <theta_deg>
// loads the PCD, applies Translation * RotZ(theta) to every point on the GPU (bevgen_cloud_manip: PCL's
// transformPointCloud op order), rasterises the 201x201 float max-height grid of the input and of the output cloud
// (saveAsMat :79-109) and writes <name>_input.csv/.png, <name>_output.csv/.png, <name>_input.pcd, <name>_output.pcd
// into the current directory, like the reference.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "bevgen.h"
#include "image_io.h"
#include "pcd_io.h"

static void save_as_mat(const float* grid, const std::string& mat_filename) {
  const int M = BEVGEN_MANIP_GRID;
  std::string txt = imgio::format_csv_f32(grid, M, M, 4);                    // set32fPrecision(4) (:97-103)
  if (!imgio::write_bytes(mat_filename, txt.data(), txt.size())) std::cerr << "Can not open file: " << mat_filename << "\n";
  std::vector<uint8_t> u8((size_t)M * M);
  for (int i = 0; i < M * M; i++) u8[i] = imgio::f32_to_u8_sat(grid[i]);      // imwrite converts CV_32F to 8U (:108)
  imgio::write_png_gray8(mat_filename + ".png", u8.data(), M, M);
}

int main(int argc, char** argv) {
  if (argc < 6) { std::cerr << "Usage: " << argv[0] << " <cloud.pcd> <tx> <ty> <tz> <theta_deg>\n"; return 1; }
  std::string input_filename(argv[1]);
  pcdio::Cloud in; std::string err;
  if (!pcdio::load(input_filename, in, &err)) std::cerr << "[pcd] " << err << std::endl;   // the reference ignores the status (:117)
  // Eigen::Affine3f: translation, then rotate(AngleAxisf(theta, UnitZ)) (:119-126), all in float
  float trans_x = std::stof(argv[2]), trans_y = std::stof(argv[3]), trans_z = std::stof(argv[4]);
  float theta = (float)(std::stof(argv[5]) / 180.0f * M_PI);                 // float / float * double -> float (:124)
  std::cout << "rotating yaw radiance: " << theta << "\n";
  // AngleAxis::toRotationMatrix for axis (0,0,1): c on the x/y diagonal, (1-c)*1*1 + c at (2,2), -/+ s off-diagonal
  float s = std::sin(theta), c = std::cos(theta);
  float m22 = (1.0f - c) * 1.0f * 1.0f + c;
  float rt[12] = {c, -s, 0.0f, trans_x, s, c, 0.0f, trans_y, 0.0f, 0.0f, m22, trans_z};

  bevgen_params p;
  bevgen_sensor_params("HDL_64E", &p);   // cloud_manip does not depend on the sensor table; any context works
  bevgen_ctx* ctx = nullptr;
  if (bevgen_create(&ctx, 0, &p, 1024, 1) != 0) { std::cerr << "bevgen_create: " << bevgen_last_error() << std::endl; return 1; }
  const size_t n = in.size();
  pcdio::Cloud out = in;
  std::vector<float> gi(BEVGEN_MANIP_GRID * BEVGEN_MANIP_GRID), go(gi.size());
  // an empty / unreadable cloud (the reference ignores loadPCDFile's status, :117) still yields the two zero grids and
  // the empty csv / png / pcd files: the library accepts NULL point arrays when n == 0
  if (bevgen_cloud_manip(ctx, (int64_t)n, rt, in.x.data(), in.y.data(), in.z.data(), out.x.data(), out.y.data(), out.z.data(), gi.data(), go.data()) != 0) {
    std::cerr << "bevgen_cloud_manip: " << bevgen_last_error() << std::endl; return 1;
  }
  bevgen_destroy(ctx);
  std::string short_name = input_filename.substr(input_filename.find_last_of('/') + 1);   // splitString(...).back() (:130-131)
  save_as_mat(gi.data(), short_name + "_input.csv");                        // :136
  save_as_mat(go.data(), short_name + "_output.csv");                       // :137
  pcdio::save_binary(short_name + "_input.pcd", in);                        // :139
  pcdio::save_binary(short_name + "_output.pcd", out);                      // :140
  return 0;
}

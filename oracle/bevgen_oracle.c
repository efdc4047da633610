/*
 * bevgen_oracle.c — CPU ORACLE for the batch_multi_bev_gen hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (CUDA library, CLI) links, loads or calls this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * It restates, expression by expression, the algorithm of the reference
 * (soytony/Point-Cloud-Preprocessing-Tools @ d94040e):
 *     BatchMultiBevGen.cpp:94-117   getOrderedCloud
 *     BatchMultiBevGen.cpp:119-252  markGroundPoints   (+ BatchMultiBevGen.h:73-99 getBelongingGrid)
 *     BatchMultiBevGen.cpp:261-321  computeAndSaveMultiBev   (binning part :278-292)
 *     BatchMultiBevGen.cpp:331-373  computeAndSaveSingleBev  (binning part :342-356)
 *     BatchMultiBevGen.cpp:502-566  selectMajorFrames  (+ src/Utility.cpp:43-49 getDistance)
 *     BatchMultiBevGen.cpp:575-636  getKeyFrameLabel
 *     src/Utility.cpp:72-124        parseSensorType / getSensorParams
 *     CloudManip.cpp:79-109,119-128 saveAsMat / rigid transform (pcl::transformPointCloud)
 *   widened rows:
 *     BatchCloudManip.cpp:201-226                saveAsMat with the label filter (oracle_bvm)
 *     TopPartRegistration.cpp:79-141             extractTopAndFlatten (oracle_top_flatten)
 *     MulranPointCloudSelect.cpp:112-126         row / col projection (oracle_project_mulran)
 *     OxfordPointCloudSelect.cpp:201-219         row / col projection (oracle_project_oxford)
 *     KittiPointCloudSelect.cpp:188-243          ring detection + column (oracle_project_kitti)
 *
 * PARITY PIN STATUS.  The reference ships no tests, golden vectors or fixtures, and its own build system cannot run
 * here (PCL, OpenCV C++, Eigen, Boost, VTK are absent).  PINNED TO THE REFERENCE'S OWN SOURCE TEXT all the same: its
 * translation units compile UNMODIFIED, where they lie under /root/reference, against the stand-in headers of
 * oracle/stub/ (recipe oracle/Makefile, target `ref`; shims oracle/ref_*_shim.cpp; outputs oracle/_ref/, git-ignored):
 *     libbevgen_ref.so / _dbl.so   BatchMultiBevGen.cpp + src/Utility.cpp   order, ground, both BEVs, poses, labels, main()
 *     libcloudmanip_ref.so         CloudManip.cpp                           saveAsMat, the tool's main()
 *     libbatchcloudmanip_ref.so    BatchCloudManip.cpp                      oracle_bvm
 *     libtoppart_ref.so            TopPartRegistration.cpp                  oracle_top_flatten
 *     lib{mulran,oxford,kitti}select_ref.so / _dbl.so   {Mulran,Oxford,Kitti}PointCloudSelect.cpp: extractPointCloud, run on scan
 *                                  files in the datasets' layouts            oracle_project_mulran / _oxford / _kitti
 *     libnanoflann_ref.so          include/nanoflann.hpp                    the KD-tree of the label stage
 * tests/test_reference_source_pin.py + tests/test_oracle_labels_ref.py hold every function above to those builds bit
 * for bit (both overload sets of the unqualified atan2 / sqrt / round); tests/golden/make_*_golden.py generate from them
 * the vectors the GPU box (no /root/reference) checks oracle and CUDA against (tests/test_golden_vectors.py).
 * No function of this file is left "parity unpinned".
 * Also kept: one hand-derived known-answer test per quirk of SURVEY §8a.1, and an independent line-by-line Python
 * restatement (oracle/bevgen_oracle_py.py) that must agree bit for bit (tests/test_oracle_cross.py).
 *
 * Third-party semantics restated (not under /root/reference, versions unpinned by the reference):
 *   PCL  PointCloud::resize value-initialises (all-zero records);  transformPointCloud (PCL >= 1.9, SSE2
 *        path) evaluates  m0*x + (m1*y + (m2*z + t))  per output component, no FMA.
 *   OpenCV cv::Mat / (cv::divide) on CV_32F is IEEE single division; Mat::zeros / 0.01*Mat::ones.
 *   glibc 2.39 libm  atan2f / sqrtf (float overloads are normative, see SURVEY §8a.1-G3;
 *        -DORACLE_DOUBLE_LIBM selects the C double overloads instead).
 *
 * Build: gcc -O3 -fno-fast-math -ffp-contract=off (x86-64 SSE2, no FMA contraction — the reference is
 * compiled -O3 without -march, CMakeLists.txt:10, so every float op is an individually rounded IEEE op).
 *
 * float -> int conversions: the reference binary executes cvttss2si/cvttsd2si, which return INT_MIN
 * ("integer indefinite") for NaN and out-of-range values.  C leaves that undefined, so it is spelled
 * out in x86_cvtt() below and used wherever the reference converts a floating value to int.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- constants of the reference ------------------------------------------------------------------ */
enum { GROUND_HEIGHT_GRID_ROWS = 75, GROUND_HEIGHT_GRID_COLS = 50 }; /* BatchMultiBevGen.cpp:25-26 */

typedef struct {
  int32_t n_scan, horizon_scan, ground_upper_scan;
  float height_res;
} oracle_sensor;

static int32_t x86_cvtt(double v) { /* cvttsd2si semantics */
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
  return (int32_t)v;
}

/* src/Utility.cpp:72-89 (substring match, first hit wins in this order) + :92-124 */
ORACLE_API int oracle_sensor_params(const char *sensor_str, oracle_sensor *out) {
  if (strstr(sensor_str, "HDL_32E")) {
    out->n_scan = 32; out->horizon_scan = 1056; out->ground_upper_scan = 20; out->height_res = 0.5f;
    return 0;
  } else if (strstr(sensor_str, "HDL_64E")) {
    out->n_scan = 64; out->horizon_scan = 2083; out->ground_upper_scan = 50; out->height_res = 0.25f;
    return 1;
  } else if (strstr(sensor_str, "OS1_64")) {
    out->n_scan = 64; out->horizon_scan = 1024; out->ground_upper_scan = 31; out->height_res = 1.0f;
    return 2;
  }
  return -1; /* UNKNOWN: the reference continues with uninitialised params; callers must treat as error */
}

/* The reference point record, BatchMultiBevGen.h:43-53 (only the fields the path reads/writes). */
typedef struct {
  float x, y, z, intensity;
  int16_t label;
} opoint;

/* ---- getOrderedCloud, BatchMultiBevGen.cpp:94-117 ------------------------------------------------ */
/* Output: S-slot SoA cloud (value-initialised => zeros, label 0) and owner[slot] = 1 + index of the
 * input point that ended up in the slot (serial loop => the LAST writer wins), 0 for an empty slot. */
ORACLE_API void oracle_order(const oracle_sensor *sp, int64_t n_in, const float *x, const float *y, const float *z,
                             const float *intensity, const uint16_t *row, const uint16_t *col, const int16_t *label,
                             float *ox, float *oy, float *oz, float *oi, int16_t *olabel, uint32_t *owner) {
  const int64_t S = (int64_t)sp->n_scan * sp->horizon_scan;
  memset(ox, 0, S * sizeof(float)); memset(oy, 0, S * sizeof(float)); memset(oz, 0, S * sizeof(float));
  memset(oi, 0, S * sizeof(float)); memset(olabel, 0, S * sizeof(int16_t)); memset(owner, 0, S * sizeof(uint32_t));
  for (int64_t i = 0; i < n_in; i++) {
    int row_idx = row[i]; /* :103-104 (uint16 -> int, never negative) */
    int col_idx = col[i];
    if (row_idx < 0 || row_idx >= sp->n_scan) continue;       /* :106 */
    if (col_idx < 0 || col_idx >= sp->horizon_scan) continue; /* :109 */
    int64_t point_idx = (int64_t)row_idx * sp->horizon_scan + col_idx; /* :113 */
    ox[point_idx] = x[i]; oy[point_idx] = y[i]; oz[point_idx] = z[i]; oi[point_idx] = intensity[i];
    olabel[point_idx] = label[i];
    owner[point_idx] = (uint32_t)(i + 1);
  }
}

/* ---- getBelongingGrid, BatchMultiBevGen.h:73-99 -------------------------------------------------- */
static void belonging_grid(float px, float py, int *sr, int *sc) {
  float normalized_x = (float)((double)px + 75.0); /* float + double literal -> double add -> float store */
  float normalized_y = (float)((double)py + 50.0);
  int sector_row_idx = x86_cvtt(floor((double)normalized_x / 2.0));
  int sector_col_idx = x86_cvtt(floor((double)normalized_y / 2.0));
  if (sector_row_idx >= 75) sector_row_idx = 75 - 1;
  if (sector_row_idx < 0) sector_row_idx = 0;
  if (sector_col_idx >= 50) sector_col_idx = 50 - 1;
  if (sector_col_idx < 0) sector_col_idx = 0;
  *sr = sector_row_idx; *sc = sector_col_idx;
}

/* ---- markGroundPoints, BatchMultiBevGen.cpp:119-252 ---------------------------------------------- */
/* In/out: olabel (ground slots set to 0).  Optional outputs (may be NULL): gm_out [S] int8 = ground_mat
 * AFTER loop 1 (before loop 3 clears it), gm_final [S] = ground_mat at function exit,
 * avg_out [75*50] = ground_grid_avg_heights after the divide. */
ORACLE_API void oracle_mark_ground(const oracle_sensor *sp, const float *ox, const float *oy, const float *oz,
                                   const float *oi, int16_t *olabel, int8_t *gm_out, int8_t *gm_final, float *avg_out) {
  const int N = sp->n_scan, H = sp->horizon_scan, G = sp->ground_upper_scan;
  const int64_t S = (int64_t)N * H;
  int8_t *ground_mat = (int8_t *)calloc(S, 1);                 /* :123 Mat::zeros CV_8S */
  float avg[GROUND_HEIGHT_GRID_ROWS * GROUND_HEIGHT_GRID_COLS]; /* :133 zeros */
  float num[GROUND_HEIGHT_GRID_ROWS * GROUND_HEIGHT_GRID_COLS]; /* :135 0.01 * ones (double 0.01 -> float) */
  for (int i = 0; i < GROUND_HEIGHT_GRID_ROWS * GROUND_HEIGHT_GRID_COLS; i++) { avg[i] = 0.0f; num[i] = (float)(0.01 * 1.0); }

  /* loop 1, :139-184 */
  for (int col_idx = 0; col_idx < H; col_idx++) {
    for (int row_idx = N - 1; row_idx > N - G - 1; row_idx--) {
      int64_t lowerInd = (int64_t)row_idx * H + col_idx;
      int64_t upperInd = (int64_t)(row_idx - 1) * H + col_idx;
      if (oi[upperInd] == -1) {                                /* :146 */
        int tmp_col_idx = (col_idx + 2) % H;
        upperInd = (int64_t)(row_idx - 1) * H + tmp_col_idx;
      }
      if (oi[upperInd] == -1) {                                /* :151 — C++ %, negative for col<2 */
        int tmp_col_idx = (col_idx - 2) % H;
        upperInd = (int64_t)(row_idx - 1) * H + tmp_col_idx;
      }
      if (oi[upperInd] == -1 && row_idx >= 2) {                /* :157 */
        int tmp_row_idx = row_idx - 2;
        upperInd = (int64_t)tmp_row_idx * H + col_idx;
      }
      if (oi[lowerInd] == -1 || oi[upperInd] == -1) {          /* :162 */
        ground_mat[(int64_t)row_idx * H + col_idx] = -1;
        continue;
      }
      float diffX = ox[upperInd] - ox[lowerInd];               /* :169-171 */
      float diffY = oy[upperInd] - oy[lowerInd];
      float diffZ = oz[upperInd] - oz[lowerInd];
#ifdef ORACLE_DOUBLE_LIBM
      float angle = (float)(atan2((double)diffZ, sqrt((double)(diffX * diffX + diffY * diffY))) * 180.0 / M_PI);
#else
      float angle = (float)((double)atan2f(diffZ, sqrtf(diffX * diffX + diffY * diffY)) * 180.0 / M_PI); /* :173 */
#endif
      float sensorMountAngle = 0.0f;
      if (fabsf(angle - sensorMountAngle) <= 10.0f) {          /* :179 */
        ground_mat[(int64_t)row_idx * H + col_idx] = 1;
        ground_mat[(int64_t)(row_idx - 1) * H + col_idx] = 1;
      }
    }
  }
  if (gm_out) memcpy(gm_out, ground_mat, S);

  /* loop 2, :187-208 */
  for (int row_idx = 0; row_idx < N; row_idx++) {
    for (int col_idx = 0; col_idx < H; col_idx++) {
      if (ground_mat[(int64_t)row_idx * H + col_idx] != 1) continue;
      int sector_row = 0, sector_col = 0;
      int64_t point_index = (int64_t)row_idx * H + col_idx;
      belonging_grid(ox[point_index], oy[point_index], &sector_row, &sector_col);
      avg[sector_row * GROUND_HEIGHT_GRID_COLS + sector_col] += oz[point_index];                     /* :198 */
      num[sector_row * GROUND_HEIGHT_GRID_COLS + sector_col] =
          num[sector_row * GROUND_HEIGHT_GRID_COLS + sector_col] + 1;                                /* :205 */
    }
  }
  for (int i = 0; i < GROUND_HEIGHT_GRID_ROWS * GROUND_HEIGHT_GRID_COLS; i++) avg[i] = avg[i] / num[i]; /* :210 */
  if (avg_out) memcpy(avg_out, avg, sizeof(avg));

  /* loop 3, :216-250; neighbour order from setNeighbors :73-84 */
  static const int nb[4][2] = {{-1, 0}, {0, 1}, {0, -1}, {1, 0}};
  for (int row_idx = 0; row_idx < N; row_idx++) {
    for (int col_idx = 0; col_idx < H; col_idx++) {
      int sector_row = 0, sector_col = 0;
      int64_t point_index = (int64_t)row_idx * H + col_idx;
      belonging_grid(ox[point_index], oy[point_index], &sector_row, &sector_col);
      for (int k = 0; k < 4; k++) {
        int neighbor_sector_row = sector_row + nb[k][0];
        int neighbor_sector_col = sector_col + nb[k][1];
        if (neighbor_sector_row < 0 || neighbor_sector_row >= 75 || neighbor_sector_col < 0 || neighbor_sector_col >= 50)
          continue;
        float d = oz[point_index] - avg[neighbor_sector_row * GROUND_HEIGHT_GRID_COLS + neighbor_sector_col];
        if ((double)d > 0.30) {                                /* :236-237 float - float, compared with a double literal */
          ground_mat[point_index] = 0;
          break;
        }
      }
      if (ground_mat[point_index] == 1) olabel[point_index] = 0; /* :244-245 */
    }
  }
  if (gm_final) memcpy(gm_final, ground_mat, S);
  free(ground_mat);
}

/* ---- binning shared by :278-292 and :342-356 ----------------------------------------------------- */
enum { MAX_RANGE = 112, MAT_SIZE = 224, NUM_BEV_LAYERS = 24 }; /* :266-268 with interval = 1.0f */
static const float LIDAR_TO_GROUND_HEIGHT = 2.0f;               /* :269 */

/* computeAndSaveMultiBev binning, :278-292.  multi = 24 * 224 * 224 bytes, layer-major (the .bin layout). */
ORACLE_API void oracle_multi_bev(const oracle_sensor *sp, const float *ox, const float *oy, const float *oz,
                                 const int16_t *olabel, uint8_t *multi) {
  const int64_t S = (int64_t)sp->n_scan * sp->horizon_scan;
  const float interval = 1.0f;
  memset(multi, 0, (size_t)NUM_BEV_LAYERS * MAT_SIZE * MAT_SIZE);
  for (int64_t i = 0; i < S; i++) {
    int x = x86_cvtt(round((double)((ox[i] + (float)MAX_RANGE) / interval) + 0.5)); /* :279 */
    int y = x86_cvtt(round((double)((oy[i] + (float)MAX_RANGE) / interval) + 0.5)); /* :280 */
    int layer_idx = x86_cvtt((double)roundf(oz[i] / sp->height_res + LIDAR_TO_GROUND_HEIGHT)); /* :281 */
    if (x < 0 || x >= MAT_SIZE || y < 0 || y >= MAT_SIZE || layer_idx < 0 || layer_idx >= NUM_BEV_LAYERS || olabel[i] == 0)
      continue;
    uint8_t *cell = &multi[(size_t)layer_idx * MAT_SIZE * MAT_SIZE + (size_t)x * MAT_SIZE + y];
    if (*cell == 0) *cell = 255;                               /* :289-291 */
  }
}

/* computeAndSaveSingleBev binning, :342-356.  single = 224 * 224 bytes. */
ORACLE_API void oracle_single_bev(const oracle_sensor *sp, const float *ox, const float *oy, const float *oz,
                                  const int16_t *olabel, uint8_t *single) {
  const int64_t S = (int64_t)sp->n_scan * sp->horizon_scan;
  const float interval = 1.0f;
  memset(single, 0, (size_t)MAT_SIZE * MAT_SIZE);
  for (int64_t i = 0; i < S; i++) {
    int x = x86_cvtt(round((double)((ox[i] + (float)MAX_RANGE) / interval) + 0.5)); /* :343 */
    int y = x86_cvtt(round((double)((oy[i] + (float)MAX_RANGE) / interval) + 0.5)); /* :344 */
    int height = x86_cvtt((double)(oz[i] + LIDAR_TO_GROUND_HEIGHT) * 4.0);           /* :345 */
    height = height < 0 ? 0 : height; height = height > 255 ? 255 : height;          /* :346 */
    if (x < 0 || x >= MAT_SIZE || y < 0 || y >= MAT_SIZE || olabel[i] == 0) continue;
    uint8_t *cell = &single[(size_t)x * MAT_SIZE + y];
    if (*cell < height) *cell = (uint8_t)height;              /* :353-355 */
  }
}

/* One frame of the hot loop, BatchMultiBevGen.cpp:735-747 (without the file encoders). */
ORACLE_API void oracle_frame(const oracle_sensor *sp, int64_t n_in, const float *x, const float *y, const float *z,
                             const float *intensity, const uint16_t *row, const uint16_t *col, const int16_t *label,
                             int16_t *label_out, uint32_t *owner_out, uint8_t *single, uint8_t *multi) {
  const int64_t S = (int64_t)sp->n_scan * sp->horizon_scan;
  float *buf = (float *)malloc(4 * S * sizeof(float));
  float *ox = buf, *oy = buf + S, *oz = buf + 2 * S, *oi = buf + 3 * S;
  oracle_order(sp, n_in, x, y, z, intensity, row, col, label, ox, oy, oz, oi, label_out, owner_out);
  oracle_mark_ground(sp, ox, oy, oz, oi, label_out, NULL, NULL, NULL);
  oracle_multi_bev(sp, ox, oy, oz, label_out, multi);
  oracle_single_bev(sp, ox, oy, oz, label_out, single);
  free(buf);
}

/* ---- batch driver with host threads (bench.py --impl reference / cpu_baseline) ------------------- */
/* Frames are independent (BatchMultiBevGen.cpp:727-757 carries no state) so the fairest many-core CPU arm
 * is one frame per thread at a time.  offsets[f]..offsets[f+1] delimit frame f in the concatenated SoA. */
typedef struct {
  const oracle_sensor *sp; int n_frames; const int64_t *offsets;
  const float *x, *y, *z, *intensity; const uint16_t *row, *col; const int16_t *label;
  int16_t *label_out; uint32_t *owner_out; uint8_t *single, *multi;
  int next; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  const int64_t S = (int64_t)j->sp->n_scan * j->sp->horizon_scan;
  for (;;) {
    pthread_mutex_lock(&j->mu); int f = j->next++; pthread_mutex_unlock(&j->mu);
    if (f >= j->n_frames) break;
    int64_t o = j->offsets[f], n = j->offsets[f + 1] - o;
    oracle_frame(j->sp, n, j->x + o, j->y + o, j->z + o, j->intensity + o, j->row + o, j->col + o, j->label + o,
                 j->label_out + f * S, j->owner_out + f * S, j->single + (size_t)f * MAT_SIZE * MAT_SIZE,
                 j->multi + (size_t)f * NUM_BEV_LAYERS * MAT_SIZE * MAT_SIZE);
  }
  return NULL;
}

ORACLE_API void oracle_frames(const oracle_sensor *sp, int n_frames, const int64_t *offsets, const float *x, const float *y,
                              const float *z, const float *intensity, const uint16_t *row, const uint16_t *col,
                              const int16_t *label, int16_t *label_out, uint32_t *owner_out, uint8_t *single, uint8_t *multi,
                              int n_threads) {
  batch_job j = {sp, n_frames, offsets, x, y, z, intensity, row, col, label, label_out, owner_out, single, multi, 0,
                 PTHREAD_MUTEX_INITIALIZER};
  if (n_threads <= 1) { batch_worker(&j); return; }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
  for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, batch_worker, &j);
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th);
}

/* ---- labels ------------------------------------------------------------------------------------- */
/* Squared L2 exactly as nanoflann L2_Adaptor::evalMetric does for dim 3 (include/nanoflann.hpp:383-407:
 * the 4-wide loop is skipped, three tail iterations of result += diff*diff with diff = query - data). */
static float d2_nanoflann(const float *q, const float *m) {
  float result = 0.0f;
  for (int d = 0; d < 3; d++) { const float diff0 = q[d] - m[d]; result += diff0 * diff0; }
  return result;
}

/* k-NN by exhaustive scan with KNNResultSet::addPoint semantics (nanoflann.hpp:175-203; strict '>' so an
 * equal distance never displaces an earlier hit).  Visiting order here is ascending index, so ties resolve
 * to the LOWEST index; the KD-tree's visiting order may differ on exact ties. */
static void knn_scan(const float *pts, int M, const float *q, int k, size_t *idx, float *dist) {
  int count = 0;
  dist[k - 1] = 3.402823466e+38f; /* KNNResultSet::init */
  for (int j = 0; j < M; j++) {
    float d = d2_nanoflann(q, pts + 3 * j);
    int i;
    for (i = count; i > 0; --i) {
      if (dist[i - 1] > d) { if (i < k) { dist[i] = dist[i - 1]; idx[i] = idx[i - 1]; } }
      else break;
    }
    if (i < k) { dist[i] = d; idx[i] = (size_t)j; }
    if (count < k) count++;
  }
}

/* selectMajorFrames, BatchMultiBevGen.cpp:502-566.  xyz = K*3 floats (Pose6f x,y,z).  major_idx must hold K ints.
 * If overlap_nn != NULL it receives, per keyframe, -1 (major), -2 (early-skipped by step 1) or the index
 * (into the major list) of the overlapping nearest major printed at :553-555. Returns M. */
ORACLE_API int oracle_select_major(int K, const float *xyz, int32_t *major_idx, int32_t *overlap_nn) {
  const float MAJOR_FRAME_INTERVAL = 20.0f;
  if (K <= 0) return 0;
  float *major_pos = (float *)malloc(sizeof(float) * 3 * K);
  int M = 0;
  major_idx[M] = 0; memcpy(major_pos, xyz, 3 * sizeof(float)); M++;  /* :518-520 */
  if (overlap_nn) overlap_nn[0] = -1;
  for (int frame_idx = 1; frame_idx < K; frame_idx++) {
    const float *last = xyz + 3 * major_idx[M - 1];
    const float *cur = xyz + 3 * frame_idx;
    /* getDistance, src/Utility.cpp:43-49 */
    float diff_x = cur[0] - last[0], diff_y = cur[1] - last[1], diff_z = cur[2] - last[2];
    float dist_to_last_major = sqrtf(diff_x * diff_x + diff_y * diff_y + diff_z * diff_z);
    if (dist_to_last_major < MAJOR_FRAME_INTERVAL) { if (overlap_nn) overlap_nn[frame_idx] = -2; continue; } /* :528 */
    size_t cand[1] = {0}; float d[1];
    knn_scan(major_pos, M, cur, 1, cand, d);                            /* :534-550 */
    if (d[0] < MAJOR_FRAME_INTERVAL * MAJOR_FRAME_INTERVAL) {           /* :552 */
      if (overlap_nn) overlap_nn[frame_idx] = (int32_t)cand[0];
      continue;
    }
    major_idx[M] = frame_idx; memcpy(major_pos + 3 * M, cur, 3 * sizeof(float)); M++; /* :561-562 */
    if (overlap_nn) overlap_nn[frame_idx] = -1;
  }
  free(major_pos);
  return M;
}

/* getKeyFrameLabel, BatchMultiBevGen.cpp:575-636.  labels = K*M floats, dense, row-major (zero-filled here).
 * Optional sparse outputs (may be NULL): nn_idx [K*2] int32, nn_w [K*2] float (w1 = 0 and idx1 = -1 for one-hot rows). */
ORACLE_API void oracle_labels(int K, const float *xyz, int M, const int32_t *major_idx, float *labels, int32_t *nn_idx,
                              float *nn_w) {
  float *major_pos = (float *)malloc(sizeof(float) * 3 * (M > 0 ? M : 1));
  for (int j = 0; j < M; j++) memcpy(major_pos + 3 * j, xyz + 3 * major_idx[j], 3 * sizeof(float)); /* :585-591 */
  if (labels) memset(labels, 0, sizeof(float) * (size_t)K * M);                                       /* :578 */
  for (int key_frame_idx = 0; key_frame_idx < K; key_frame_idx++) {
    size_t cand[2] = {0, 0}; float d[2] = {0.0f, 0.0f};  /* std::vector value-init, :604-605 */
    knn_scan(major_pos, M, xyz + 3 * key_frame_idx, 2, cand, d);
    if (key_frame_idx == major_idx[cand[0]]) {                                                   /* :616 */
      if (labels) labels[(size_t)key_frame_idx * M + cand[0]] = 1.0f;
      if (nn_idx) { nn_idx[2 * key_frame_idx] = (int32_t)cand[0]; nn_idx[2 * key_frame_idx + 1] = -1; }
      if (nn_w) { nn_w[2 * key_frame_idx] = 1.0f; nn_w[2 * key_frame_idx + 1] = 0.0f; }
    } else {
      float weight_0 = (float)(1.0 / ((double)d[0] + 1e-5));   /* :623 1.0f / (float + double) -> double -> float */
      float weight_1 = (float)(1.0 / ((double)d[1] + 1e-5));   /* :624 */
      float sum_weights = weight_0 + weight_1;
      weight_0 /= sum_weights; weight_1 /= sum_weights;
      if (labels) {
        labels[(size_t)key_frame_idx * M + cand[0]] = weight_0; /* :629 */
        labels[(size_t)key_frame_idx * M + cand[1]] = weight_1; /* :630 (M==1: overwrites index 0) */
      }
      if (nn_idx) { nn_idx[2 * key_frame_idx] = (int32_t)cand[0]; nn_idx[2 * key_frame_idx + 1] = (int32_t)cand[1]; }
      if (nn_w) { nn_w[2 * key_frame_idx] = weight_0; nn_w[2 * key_frame_idx + 1] = weight_1; }
    }
  }
  free(major_pos);
}

/* ---- cloud_manip (BASELINE config #5) ------------------------------------------------------------ */
/* pcl::transformPointCloud(cloud_in, cloud_out, Affine3f) as called at CloudManip.cpp:128; rt = row-major
 * 3x4 [R|t].  PCL >= 1.9 SSE2 Transformer<float>::se3: p0 + (p1 + (p2 + c3)) with pK = src[K] * colK. */
ORACLE_API void oracle_transform(int64_t n, const float *rt, const float *x, const float *y, const float *z, float *tx,
                                 float *ty, float *tz) {
  for (int64_t i = 0; i < n; i++) {
    float px = x[i], py = y[i], pz = z[i];
    float o[3];
    for (int r = 0; r < 3; r++) {
      float p0 = px * rt[4 * r + 0], p1 = py * rt[4 * r + 1], p2 = pz * rt[4 * r + 2];
      float t2 = p2 + rt[4 * r + 3];
      float t1 = p1 + t2;
      o[r] = p0 + t1;
    }
    tx[i] = o[0]; ty[i] = o[1]; tz[i] = o[2];
  }
}

/* saveAsMat binning, CloudManip.cpp:79-95 (called with interval = 1.0f at :134-137 => MAT_SIZE = 201). */
ORACLE_API void oracle_save_as_mat(int64_t n, const float *x, const float *y, const float *z, float *cart_bv /*201*201*/) {
  const int MAXR = 100; const float interval = 1.0f;
  const int MS = (int)(MAXR * 2 / interval + 1);
  for (int i = 0; i < MS * MS; i++) cart_bv[i] = 0.0f;
  for (int64_t i = 0; i < n; i++) {
    int xi = x86_cvtt(round((double)((x[i] + (float)MAXR) / interval) + 0.5));
    int yi = x86_cvtt(round((double)((y[i] + (float)MAXR) / interval) + 0.5));
    if (xi < 0 || xi >= MS || yi < 0 || yi >= MS) continue;
    float v = z[i] + 2.0f;
    if (v > cart_bv[xi * MS + yi]) cart_bv[xi * MS + yi] = v;
  }
}

/* batch_cloud_manip's saveAsMat, BatchCloudManip.cpp:201-226, called with interval_res = 1.0f (:310) on the ordered,
 * ground-marked cloud (:305-320) => MAT_SIZE = 201 (static, fixed by the first call).  Differs from CloudManip.cpp's
 * version only by the `pi.label == 0` skip (:214).  x/y/z/label: the S slots of the ordered cloud. */
ORACLE_API void oracle_bvm(int64_t n, const float *x, const float *y, const float *z, const int16_t *label,
                           float *cart_bv /*201*201*/) {
  const int MAXR = 100; const float interval = 1.0f;
  const int MS = (int)(MAXR * 2 / interval + 1);                                   /* :208 */
  for (int i = 0; i < MS * MS; i++) cart_bv[i] = 0.0f;                             /* :209 cv::Mat::zeros */
  for (int64_t i = 0; i < n; i++) {
    int xi = x86_cvtt(round((double)((x[i] + (float)MAXR) / interval) + 0.5));     /* :211 */
    int yi = x86_cvtt(round((double)((y[i] + (float)MAXR) / interval) + 0.5));     /* :212 */
    if (xi < 0 || xi >= MS || yi < 0 || yi >= MS || label[i] == 0) continue;       /* :214 */
    float v = z[i] + 2.0f;
    if (v > cart_bv[xi * MS + yi]) cart_bv[xi * MS + yi] = v;                      /* :218-220 */
  }
}

/* Projection step of the keyframe extractors.  atan2 / sqrt / round are called unqualified on float
 * arguments; as for the ground criterion (8a.1-G3) the float overloads are normative (-DORACLE_DOUBLE_LIBM: double). */
#ifdef ORACLE_DOUBLE_LIBM
#define P_ATAN2(a, b) atan2((double)(a), (double)(b))
#define P_SQRT(a) sqrt((double)(a))
#else
#define P_ATAN2(a, b) ((double)atan2f((a), (b)))
#define P_SQRT(a) sqrtf(a)
#endif
static uint16_t u16_cast(float v) { return (uint16_t)(x86_cvtt((double)v) & 0xFFFF); }   /* cvttss2si + truncation */

/* MulranPointCloudSelect.cpp:112-126 */
ORACLE_API void oracle_project_mulran(int64_t n, const float *x, const float *y, uint16_t *row, uint16_t *col) {
  for (int64_t k = 0; k < n; k++) {
    row[k] = (uint16_t)(k % 64);                                                    /* :120 */
    float azimuthal_angle = (float)(P_ATAN2(y[k], x[k]) / M_PI * 180.0f);           /* :121 */
    if (azimuthal_angle > 360.0f) azimuthal_angle = azimuthal_angle - 360.0f;       /* :122 */
    else if (azimuthal_angle < 0.0f) azimuthal_angle = azimuthal_angle + 360.0f;    /* :123 */
    col[k] = u16_cast(roundf(azimuthal_angle / 360.0f * 1024));                     /* :125 */
  }
}

/* OxfordPointCloudSelect.cpp:201-219; x and z are negated in place. */
ORACLE_API void oracle_project_oxford(int64_t n, float *x, const float *y, float *z, uint16_t *row, uint16_t *col) {
  for (int64_t k = 0; k < n; k++) {
    x[k] = -x[k]; z[k] = -z[k];                                                     /* :203-204 */
    float elevation_angle = (float)((double)P_ATAN2(z[k], P_SQRT(x[k] * x[k] + y[k] * y[k])) / M_PI * 180.0f);   /* :208 */
    int row_idx = x86_cvtt(round((-elevation_angle + 10.67) / 1.3335));             /* :209 */
    row_idx = row_idx > 0 ? row_idx : 0; row_idx = row_idx < 31 ? row_idx : 31;     /* :210 */
    row[k] = (uint16_t)row_idx;
    float azimuthal_angle = (float)(P_ATAN2(y[k], x[k]) / M_PI * 180.0f);           /* :213 */
    if (azimuthal_angle > 360.0f) azimuthal_angle = azimuthal_angle - 360.0f;
    else if (azimuthal_angle < 0.0f) azimuthal_angle = azimuthal_angle + 360.0f;
    uint16_t c = u16_cast(roundf(azimuthal_angle / 360.0f * 1056));                 /* :216 */
    if (c >= 1056) c -= 1056;                                                       /* :217 (:218 is dead code for an unsigned) */
    col[k] = c;
  }
}

/* KittiPointCloudSelect.cpp:188-243: azimuth per point, ring detection, column index.  row/col = 0xFFFF for the points
 * the extractor does not place into the structured cloud (point 0; rings outside 0..63). */
ORACLE_API void oracle_project_kitti(int64_t n, const float *x, const float *y, uint16_t *row, uint16_t *col) {
  const int N_SCAN = 64, Horizon_SCAN = 2083;                                       /* :148-149 */
  if (n <= 0) return;
  float *azimuth_angle = (float *)malloc((size_t)n * sizeof(float));
  for (int64_t i = 0; i < n; i++) azimuth_angle[i] = (float)(P_ATAN2(y[i], x[i]) / M_PI * 180.0f);   /* :191-194 */
  int32_t ring_idx = azimuth_angle[0] > 0 ? 0 : -1;                                 /* :199-204 */
  int num_points_on_this_ring = 0;
  row[0] = col[0] = 0xFFFF;
  for (int64_t i = 1; i < n; i++) {                                                 /* :211 */
    if (azimuth_angle[i - 1] <= 0 && azimuth_angle[i] > 0) {                        /* :213 */
      if (ring_idx == -1) { ring_idx = 0; num_points_on_this_ring = 0; }
      else if ((float)num_points_on_this_ring > (float)Horizon_SCAN * 0.60f) { ring_idx++; num_points_on_this_ring = 0; }
    }
    float a = azimuth_angle[i];                                                     /* makeAngleSemiPositive :137-146 */
    if (a >= 360.0f) a = a - 360.0f; else if (a < 0) a = a + 360.0f;
    int col_idx = x86_cvtt(round((double)a / (360.0 / Horizon_SCAN)));              /* :226 */
    row[i] = col[i] = 0xFFFF;
    if (ring_idx >= 0 && ring_idx < N_SCAN) {                                       /* :228 */
      if (col_idx >= Horizon_SCAN) col_idx = col_idx - Horizon_SCAN;
      else if (col_idx < 0) col_idx = col_idx + Horizon_SCAN;
      if (col_idx >= 0 && col_idx < Horizon_SCAN) { row[i] = (uint16_t)ring_idx; col[i] = (uint16_t)col_idx; }
    }
    num_points_on_this_ring++;                                                      /* :241 */
  }
  free(azimuth_angle);
}

/* extractTopAndFlatten, TopPartRegistration.cpp:79-141.  out_index[k] = index of the k-th output point, out_x/out_y its
 * coordinates (z is set to 0 by the reference, :135).  Returns the number of output points.  The reference sorts every
 * cell with std::sort (order of equal heights unspecified); this restatement uses a stable order (height descending,
 * then input index), which is one of the orders std::sort may produce. */
typedef struct { float z; int idx; } top_item;
static int top_cmp(const void *a, const void *b) {
  const top_item *p = (const top_item *)a, *q = (const top_item *)b;
  if (p->z > q->z) return -1;                                                      /* :130-132: z_1 > z_2 first */
  if (q->z > p->z) return 1;
  return (p->idx > q->idx) - (p->idx < q->idx);
}
ORACLE_API int64_t oracle_top_flatten(int64_t n, const float *x, const float *y, const float *z, const int16_t *label,
                                      float *out_x, float *out_y, uint32_t *out_index) {
  const int NUM_GRID_X = 10, NUM_GRID_Y = 10;                                       /* :83-84 */
  const float MAX_RADIUS_X = 100.0f, MAX_RADIUS_Y = 100.0f;                         /* :85-86 */
  const float GRID_RES_X = 2.0f * MAX_RADIUS_X / NUM_GRID_X, GRID_RES_Y = 2.0f * MAX_RADIUS_Y / NUM_GRID_Y;   /* :87-88 */
  const int MIN_GRID_POINTS_SIZE = 20;                                              /* :90 */
  int *cell = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
  int cnt[100] = {0};
  for (int64_t i = 0; i < n; i++) {
    cell[i] = -1;
    if (label[i] == 0) continue;                                                    /* :99-101 */
    int grid_x = x86_cvtt((double)roundf((x[i] + MAX_RADIUS_X) / GRID_RES_X));      /* :104 */
    int grid_y = x86_cvtt((double)roundf((y[i] + MAX_RADIUS_Y) / GRID_RES_Y));      /* :105 */
    if (grid_x < 0 || grid_x >= NUM_GRID_X || grid_y < 0 || grid_y >= NUM_GRID_Y) continue;   /* :107-109 */
    cell[i] = grid_x * NUM_GRID_Y + grid_y; cnt[cell[i]]++;
  }
  top_item *items = (top_item *)malloc((size_t)(n > 0 ? n : 1) * sizeof(top_item));
  int64_t m = 0;
  for (int c = 0; c < NUM_GRID_X * NUM_GRID_Y; c++) {                               /* :116-117 */
    int num_points_in_grid = cnt[c];
    int num_points_needed = x86_cvtt((double)roundf(0.2f * (float)num_points_in_grid));   /* :123 */
    if (num_points_in_grid < MIN_GRID_POINTS_SIZE) continue;                        /* :124-126 */
    int k = 0;
    for (int64_t i = 0; i < n; i++) if (cell[i] == c) { items[k].z = z[i]; items[k].idx = (int)i; k++; }
    qsort(items, (size_t)k, sizeof(top_item), top_cmp);                             /* :129-132 */
    for (int j = 0; j < num_points_needed; j++) {                                   /* :134-139 */
      out_x[m] = x[items[j].idx]; out_y[m] = y[items[j].idx]; out_index[m] = (uint32_t)items[j].idx; m++;
    }
  }
  free(items); free(cell);
  return m;
}

/* Exposed for tests: the float libm the oracle was built against. */
ORACLE_API float oracle_atan2f(float y, float x) { return atan2f(y, x); }
ORACLE_API float oracle_angle_deg(float dz, float dx, float dy) {
  return (float)((double)atan2f(dz, sqrtf(dx * dx + dy * dy)) * 180.0 / M_PI);
}

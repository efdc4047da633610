/*
 * bevgen.h — C-ABI of libbevgen_cuda.so: the B200 (sm_100a) implementation of the batch_multi_bev_gen hot path.
 *
 * The reference (soytony/Point-Cloud-Preprocessing-Tools @ d94040e) has no library / plugin / FFI interface for
 * this path: it sits behind a CLI (BatchMultiBevGen.cpp:664-771) and a directory contract.  This header is the
 * boundary a host program (our C++ `batch_multi_bev_gen` CLI, or the reference's own main() — see INTEGRATION.md)
 * binds instead of calling the reference's free functions.  Each entry point names the reference code it replaces.
 *
 * Conventions: plain C types only; every function returns 0 on success or a negative code and leaves a message
 * retrievable through bevgen_last_error() (thread-local).  One context per device; a context is used from one
 * host thread at a time; contexts are independent (thread-per-GPU sharding, no collective).  There is NO CPU
 * fallback: if no CUDA device / sm_100 kernel image is usable, bevgen_create() fails.
 *
 * Frames are passed as concatenated SoA arrays ("points") plus offsets[F+1]: frame f owns [offsets[f], offsets[f+1]).
 * Field meaning = pcl::PointXYZIRCT (BatchMultiBevGen.h:43-66); `t` never goes to the GPU.
 */
#ifndef BEVGEN_H_
#define BEVGEN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BEVGEN_API __attribute__((visibility("default")))
#else
#define BEVGEN_API
#endif

#define BEVGEN_GRID_SIZE 224        /* MAX_RANGE*2/interval, BatchMultiBevGen.cpp:266-267,336-337 */
#define BEVGEN_MAX_RANGE 112
#define BEVGEN_NUM_LAYERS 24        /* BatchMultiBevGen.cpp:268 */
#define BEVGEN_SECTOR_ROWS 75       /* BatchMultiBevGen.cpp:25 */
#define BEVGEN_SECTOR_COLS 50       /* BatchMultiBevGen.cpp:26 */
#define BEVGEN_MANIP_GRID 201       /* CloudManip.cpp:81-82 with interval 1.0f */

typedef struct bevgen_ctx bevgen_ctx;

/* SensorParams (include/Utility.h:30-36) + the literals of the BEV stage + optional rigid transform. */
typedef struct bevgen_params {
  int32_t n_scan, horizon_scan, ground_upper_scan; /* src/Utility.cpp:92-124 */
  float height_res;
  int32_t grid_size;      /* must be 224 */
  int32_t max_range;      /* must be 112 */
  int32_t n_layers;       /* must be 24  */
  float lidar_to_ground;  /* must be 2.0f, BatchMultiBevGen.cpp:269 */
  float rt[12];           /* row-major 3x4 [R|t]; applied to every point before ordering when has_transform != 0   */
  int32_t has_transform;  /* (pcl::transformPointCloud semantics, CloudManip.cpp:119-128); 0 = identity, skipped    */
} bevgen_params;

/* Concatenated SoA input (host or device pointers, depending on the call). */
typedef struct bevgen_points {
  const float *x, *y, *z, *intensity;
  const uint16_t *row, *col;
  const int16_t *label;
} bevgen_points;

/* Per-frame outputs, frame-major. S = n_scan*horizon_scan.
 *   label      [F][S]           labels of the ordered cloud after markGroundPoints (0 = ground or empty slot)
 *   winner_bits                 one bit per INPUT point: set iff the point is the last writer of its (row, col) slot in
 *                               getOrderedCloud's serial loop (:102-116), i.e. the record that ends up in the ordered
 *                               cloud savePCDFileBinary writes (:756); slot = row*Horizon_SCAN + col is the caller's own
 *                               data.  Frame f's bits start at 32-bit word (offsets[f] >> 5) + f; point i of the frame is
 *                               bit (i & 31) of word (i >> 5) from there.  bevgen_winner_words(n_total, F) words in all;
 *                               words between two frames are unspecified.
 *   single_bev [F][224][224]    computeAndSaveSingleBev matrix (:340-356)
 *   multi_bev  [F][24][224][224] computeAndSaveMultiBev layers = the .bin payload (:271-292, :307-314)
 *   bvm        [F][201][201] f32 OPTIONAL (NULL = not computed): batch_cloud_manip's bird-view map, saveAsMat of
 *                               BatchCloudManip.cpp:201-226 with interval 1.0f on the ground-removed ordered cloud
 *                               (max of z + 2.0f per 1 m cell, label == 0 skipped) - SURVEY 8(f)-3                */
typedef struct bevgen_outputs {
  int16_t *label;
  uint32_t *winner_bits;
  uint8_t *single_bev;
  uint8_t *multi_bev;
  float *bvm;
} bevgen_outputs;

/* Number of 32-bit words of bevgen_outputs.winner_bits for F frames holding n_total points. */
static inline size_t bevgen_winner_words(int64_t n_total, int n_frames) { return (size_t)(n_total >> 5) + (size_t)n_frames + 1; }

/* parseSensorType + getSensorParams (src/Utility.cpp:72-124): substring match "HDL_32E" / "HDL_64E" / "OS1_64";
 * fills the BEV literals and an identity transform.  Returns the SensorType enum value (0,1,2) or -1 if unknown
 * (the reference prints "Unknown sensor type" and runs on uninitialised params; here it is an error). */
BEVGEN_API int bevgen_sensor_params(const char *sensor_type, bevgen_params *out);

/* Replaces the per-process globals of BatchMultiBevGen.cpp:29-37.  max_points_per_frame bounds n_in of one frame,
 * max_frames_per_batch sizes the device scratch (frames processed per launch wave). */
BEVGEN_API int bevgen_create(bevgen_ctx **ctx, int device, const bevgen_params *params, int max_points_per_frame,
                  int max_frames_per_batch);
BEVGEN_API void bevgen_destroy(bevgen_ctx *ctx);
BEVGEN_API const char *bevgen_last_error(void);

/* Pinned host memory for staging (cudaHostAlloc); process_host/submit run fully async only on such buffers. */
BEVGEN_API void *bevgen_host_alloc(size_t bytes);
BEVGEN_API void bevgen_host_free(void *p);
/* The same, write-combined (cudaHostAllocWriteCombined): for INPUT staging the host only ever writes sequentially and the
 * copy engine reads - the PCIe reads are not snooped against the CPU caches, which matters when the GPUs of a box
 * stage at once.  Never read such a buffer on the host (uncached reads); free with bevgen_host_free. */
BEVGEN_API void *bevgen_host_alloc_wc(size_t bytes);

/* The serial hot loop body BatchMultiBevGen.cpp:735-747 (getOrderedCloud + markGroundPoints + both BEVs, without
 * the file encoders) for n_frames frames.
 *   _host:   `in` / `out` are HOST buffers; H2D on a copy stream, kernels on the compute stream and D2H on a third
 *            stream are pipelined over chunks of max_frames_per_batch frames; returns when `out` is complete.
 *   _device: `in` / `out` are DEVICE buffers on the context's device (offsets stays a host array); work is
 *            enqueued on the context's compute stream; call bevgen_sync() before reading `out`.
 * Every array of `in` and `out` is required (a NULL one is refused with -1, nothing is launched); only `out->bvm` is
 * optional, and a batch without a single point may come with NULL point arrays.                               */
BEVGEN_API int bevgen_process_host(bevgen_ctx *ctx, int n_frames, const int64_t *offsets, const bevgen_points *in,
                        const bevgen_outputs *out);
BEVGEN_API int bevgen_process_device(bevgen_ctx *ctx, int n_frames, const int64_t *offsets, const bevgen_points *in,
                          const bevgen_outputs *out);
BEVGEN_API int bevgen_sync(bevgen_ctx *ctx);

/* ---- compact staging format: the host path with the fewest bytes over PCIe ----------------------------------------
 * The loop body BatchMultiBevGen.cpp:735-747 needs, per input point, x / y / z, the slot row*Horizon_SCAN + col
 * (:106-113), whether intensity == -1 (:146-165) and whether label != 0 (:285, :349) - 16 bytes instead of the 22 of
 * the SoA form; and its results are, per frame, which slots became ground (:244-245, one bit per slot; every other
 * slot keeps the label of the point that owns it, which the caller still holds), the winner bits, the single BEV and
 * the multi BEV, whose 24 layers are 0/255 images, i.e. 24 bits per cell.  Same kernels, same results, 2.0 MB in /
 * 0.23 MB out per HDL_64E frame instead of 2.6 MB / 1.5 MB.  The host-side packers / expanders below restate nothing of
 * the algorithm: they only change the representation. */
#define BEVGEN_META_INVALID 0x00FFFFFFu   /* slot field of a point getOrderedCloud drops (row >= N_SCAN or col >= Horizon_SCAN) */
#define BEVGEN_META_NEG1 (1u << 24)       /* intensity == -1 */
#define BEVGEN_META_LABELED (1u << 25)    /* label != 0 */
typedef struct bevgen_points_compact {
  const float *x, *y, *z;
  const uint32_t *meta;                   /* bevgen_pack_meta() per point */
} bevgen_points_compact;
typedef struct bevgen_outputs_compact {
  uint32_t *ground_bits;   /* [F][(S+31)/32]: bit s&31 of word s>>5 set iff slot s is ground after markGroundPoints (label -> 0) */
  uint32_t *winner_bits;   /* as in bevgen_outputs */
  uint8_t *single_bev;     /* [F][224][224] */
  uint8_t *multi_planes;   /* [F][3][224][224]: bit (l & 7) of plane (l >> 3) set iff layer l of the cell is occupied (255) */
} bevgen_outputs_compact;
static inline uint32_t bevgen_pack_meta(const bevgen_params *p, uint16_t row, uint16_t col, float intensity, int16_t label) {
  const uint32_t slot = (row < p->n_scan && col < p->horizon_scan) ? (uint32_t)row * (uint32_t)p->horizon_scan + col : BEVGEN_META_INVALID;
  return slot | (intensity == -1.0f ? BEVGEN_META_NEG1 : 0u) | (label != 0 ? BEVGEN_META_LABELED : 0u);
}
/* multi_planes of one frame -> the 24 layers of the .bin payload (:307-314) */
static inline void bevgen_expand_multi(const uint8_t *planes, uint8_t *multi_bev) {
  for (int l = 0; l < BEVGEN_NUM_LAYERS; l++) {
    const uint8_t *pl = planes + (size_t)(l >> 3) * BEVGEN_GRID_SIZE * BEVGEN_GRID_SIZE;
    uint8_t *o = multi_bev + (size_t)l * BEVGEN_GRID_SIZE * BEVGEN_GRID_SIZE;
    for (int i = 0; i < BEVGEN_GRID_SIZE * BEVGEN_GRID_SIZE; i++) o[i] = ((pl[i] >> (l & 7)) & 1) ? 255 : 0;
  }
}
/* Same contract as bevgen_process_host (chunked H2D | kernels | D2H pipeline; pinned buffers make it fully async). */
BEVGEN_API int bevgen_process_host_compact(bevgen_ctx *ctx, int n_frames, const int64_t *offsets,
                                           const bevgen_points_compact *in, const bevgen_outputs_compact *out);

/* SURVEY 8(f)-1 — packed-record staging: the frames arrive as the interleaved records of a binary PCD payload
 * (what pcl::io::loadPCDFile parses at BatchMultiBevGen.cpp:730 and savePCDFileBinary writes at :756) and the
 * de-interleave into SoA runs on the GPU, so host staging is one copy of the file payload into pinned memory.
 * Field types are those of pcl::PointXYZIRCT (BatchMultiBevGen.h:43-66): x, y, z, intensity f32; row, col u16;
 * label i16.  Offsets are bytes inside a record (any order, any padding, no alignment requirement); -1 = the field is
 * absent and reads as 0 (pcl::fromPCLPointCloud2 leaves a missing field value-initialised). */
typedef struct bevgen_record_layout {
  int32_t stride;                              /* bytes per record, 1..256 */
  int32_t off_x, off_y, off_z, off_intensity;  /* f32 */
  int32_t off_row, off_col;                    /* u16 */
  int32_t off_label;                           /* i16 */
} bevgen_record_layout;
/* The 26-byte layout of a PointXYZIRCT binary PCD: x@0 y@4 z@8 intensity@12 row@16 col@18 (t@20) label@24. */
BEVGEN_API int bevgen_pcd_record_layout(bevgen_record_layout *out);
/* Same contract as bevgen_process_host; `records` = HOST buffer with the concatenated records of all frames
 * (frame f = records [offsets[f], offsets[f+1]), i.e. bytes from offsets[f]*stride). */
BEVGEN_API int bevgen_process_packed_host(bevgen_ctx *ctx, int n_frames, const int64_t *offsets, const void *records,
                               const bevgen_record_layout *layout, const bevgen_outputs *out);

/* Asynchronous single-frame form used by pipelined callers (one frame of the loop at :727-757):
 * submit copies the frame into the context's pinned ring and enqueues H2D + kernels + D2H; collect blocks on that
 * frame's event and copies the results out.  At most `max_frames_per_batch` frames may be in flight (ring slots are
 * built on demand).  The point arrays may be NULL only when n_in == 0; collect's output pointers may each be NULL. */
BEVGEN_API int bevgen_submit(bevgen_ctx *ctx, int frame_id, int n_in, const float *x, const float *y, const float *z,
                  const float *intensity, const uint16_t *row, const uint16_t *col, const int16_t *label);
BEVGEN_API int bevgen_collect(bevgen_ctx *ctx, int frame_id, int16_t *label_out, uint32_t *winner_bits /* (n_in+31)/32 words */,
                   uint8_t *single_bev, uint8_t *multi_bev);

/* selectMajorFrames (BatchMultiBevGen.cpp:502-566).  xyz = K*3 host floats (Pose6f x,y,z, :441-444).
 * major_idx (host, capacity K) receives the M major-frame indices; *n_major = M.  overlap_nn (host, K, may be
 * NULL): -1 major, -2 early-skipped (:528), else index into the major list of the overlapping major (:553). */
BEVGEN_API int bevgen_select_major(bevgen_ctx *ctx, int K, const float *xyz, int32_t *major_idx, int32_t *n_major,
                        int32_t *overlap_nn);
/* getKeyFrameLabel (BatchMultiBevGen.cpp:575-636) for keyframe rows [row_begin,row_end) — the per-GPU row split.
 * labels_out (host, may be NULL): dense (row_end-row_begin)*M floats.  nn_idx / nn_w (host, may be NULL): the two
 * non-zeros of each row, [rows][2]; one-hot rows have nn_idx[1] = -1, nn_w = {1,0}. */
BEVGEN_API int bevgen_labels(bevgen_ctx *ctx, int K, const float *xyz, int M, const int32_t *major_idx, int row_begin,
                  int row_end, float *labels_out, int32_t *nn_idx, float *nn_w);

/* cloud_manip (CloudManip.cpp:111-141): rigid transform of n host points (rt as in bevgen_params) and the two
 * 201x201 float max-height grids of saveAsMat (:79-95) for the input and the transformed cloud.
 * Any output pointer may be NULL, each on its own; x / y / z may be NULL when n == 0 (the reference carries on with an empty
 * cloud when loadPCDFile fails, CloudManip.cpp:117).  Works on any context (sensor params are not used). */
BEVGEN_API int bevgen_cloud_manip(bevgen_ctx *ctx, int64_t n, const float *rt, const float *x, const float *y, const float *z,
                       float *tx, float *ty, float *tz, float *bev_in, float *bev_out);

/* Device-resident form of bevgen_cloud_manip (BASELINE config #5 measured without PCIe): all pointers are device memory,
 * tx/ty/tz given together or all NULL, grids of 201*201 floats are overwritten; enqueued on the compute stream. */
BEVGEN_API int bevgen_cloud_manip_device(bevgen_ctx *ctx, int64_t n, const float *rt, const float *x, const float *y, const float *z,
                              float *tx, float *ty, float *tz, float *bev_in, float *bev_out);

/* SURVEY 8(f)-2 — the per-point projection of the keyframe extractors, i.e. the producer of the `row` / `col` fields:
 *   BEVGEN_PROJECT_MULRAN_OS1_64  (MulranPointCloudSelect.cpp:112-126): row = k % 64, col = round(azimuth / 360 * 1024)
 *                                 (col may come out equal to 1024; getOrderedCloud drops such points, :106-109)
 *   BEVGEN_PROJECT_OXFORD_HDL_32E (OxfordPointCloudSelect.cpp:201-219): x and z are NEGATED IN PLACE (sensor mounted
 *                                 upside-down), row from the elevation angle clamped to 0..31, col wrapped at 1056.
 *   BEVGEN_PROJECT_KITTI_HDL_64E  (KittiPointCloudSelect.cpp:188-243): ring detection from the azimuth sign changes
 *                                 (a new ring only after more than 2083 * 0.60f points), col = round(az / (360.0 / 2083));
 *                                 points the extractor does not place (point 0, rings outside 0..63) get row = col =
 *                                 0xFFFF, which getOrderedCloud's bounds test drops.  The caller sets intensity = -1 and
 *                                 label = -2 like :236-238.  A scan whose azimuth is NaN anywhere is undefined
 *                                 behaviour in the reference (out-of-bounds index) and yields unplaced points here.
 * Host arrays of n points in file order; z may be NULL except for OXFORD. */
#define BEVGEN_PROJECT_MULRAN_OS1_64 0
#define BEVGEN_PROJECT_OXFORD_HDL_32E 1
#define BEVGEN_PROJECT_KITTI_HDL_64E 2
BEVGEN_API int bevgen_project(bevgen_ctx *ctx, int kind, int64_t n, float *x, const float *y, float *z, uint16_t *row,
                   uint16_t *col);

/* SURVEY 8(f)-4 — extractTopAndFlatten (TopPartRegistration.cpp:79-141, BatchTopPartRegistration.cpp:90): the consumer
 * of non_ground_point_cloud/.  Points with label != 0 are binned into a 10 x 10 grid of 20 m cells; every cell holding
 * at least 20 points keeps its round(0.2f * count) highest points; output = those points cell after cell (grid_x major),
 * highest first, with z dropped (the reference sets it to 0).  out_x / out_y / out_index (optional: index of the source
 * point) need room for n entries; *n_out = number of points written.  Points of equal height keep their input order
 * (the reference's std::sort leaves that order unspecified); NaN heights are undefined behaviour in the reference. */
BEVGEN_API int bevgen_top_flatten(bevgen_ctx *ctx, int64_t n, const float *x, const float *y, const float *z, const int16_t *label,
                       float *out_x, float *out_y, uint32_t *out_index, int64_t *n_out);

/* ---- introspection for bench / tests (no reference counterpart) ---------------------------------------------- */
#define BEVGEN_N_STAGES 8
/* Stage order: 0 clear, 1 order_winners (k_order_winners, or k_order_claim for range images too large for shared memory),
 * 2 order_scatter (k_order_scatter, or k_order_fill + k_winner_bits), 3 ground_mark, 4 sector_mean (k_seg_build + k_seg_fold +
 * the k_sector_mean fallback), 5 finalize_bin_scatter (k_finalize_bin, + k_float_bev when bvm is requested), 6..7 reserved.
 * When profiling is enabled, process_device runs its waves on one stream and brackets every stage with CUDA events;
 * stage_ms returns the accumulated milliseconds and launch counts since the last reset. */
BEVGEN_API int bevgen_set_profiling(bevgen_ctx *ctx, int enabled);
BEVGEN_API int bevgen_stage_ms(bevgen_ctx *ctx, float *ms /*[BEVGEN_N_STAGES]*/, int64_t *launches /*[BEVGEN_N_STAGES]*/);
BEVGEN_API int64_t bevgen_kernel_launches(bevgen_ctx *ctx); /* total kernels launched by this context so far */
BEVGEN_API void *bevgen_compute_stream(bevgen_ctx *ctx);    /* cudaStream_t of the compute stream */
BEVGEN_API const char *bevgen_stage_name(int stage);
/* Which C++ overload set the reference's unqualified `atan2(diffZ, sqrt(...))` (BatchMultiBevGen.cpp:173) binds to is
 * decided by the reference's include tree (oracle/stub/README.md): 0 (default) = float atan2f / sqrtf, the build with
 * <math.h> visible; 1 = the C double functions.  The two only differ for pairs within ~1 float ulp of the 10-degree
 * threshold.  Applies to every later process_* / submit call of the context. */
BEVGEN_API int bevgen_set_libm(bevgen_ctx *ctx, int use_double);
/* Diagnostics of the ground criterion, accumulated over all frames processed while enabled (set_diag zeroes them):
 * out[0] = vertical pairs whose slope fell inside the +-2e-5 guard band around tan(10 deg) and were decided by the exact
 * libm evaluation, out[1] = pairs whose decision differs between the float and the double overload set,
 * out[2..3] reserved.  Enabling costs a double atan2 per borderline / degenerate pair. */
BEVGEN_API int bevgen_set_diag(bevgen_ctx *ctx, int enabled);
BEVGEN_API int bevgen_get_diag(bevgen_ctx *ctx, uint64_t *out /*[4]*/);
/* Device port of glibc's float atan2f used by the ground criterion, exposed for bit-exactness tests. */
BEVGEN_API int bevgen_debug_atan2f(bevgen_ctx *ctx, int64_t n, const float *y, const float *x, float *out);

#ifdef __cplusplus
}
#endif
#endif /* BEVGEN_H_ */

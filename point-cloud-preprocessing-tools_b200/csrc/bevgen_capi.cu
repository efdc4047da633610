// bevgen_capi.cu — C-ABI (include/bevgen.h) over the sm_100a kernels in bevgen_kernels.cuh.
// Host side of the library: context, streams, device scratch, chunked H2D / compute / D2H pipeline.
// No PyTorch, no Triton, no CPU fallback: every entry point that computes launches CUDA kernels or fails.
#include "../../include/bevgen.h"
#include "bevgen_kernels.cuh"

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace bevgen;

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));  \
      g_err = b_;                                                                                  \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

namespace {

struct DevIn { float *x = 0, *y = 0, *z = 0, *inten = 0; uint16_t *row = 0, *col = 0; int16_t* label = 0; };
struct DevOut { int16_t* label = 0; uint32_t* wbits = 0; uint8_t* single = 0; uint8_t* multi = 0; float* bvm = 0; };
constexpr size_t BVM_CELLS = (size_t)MGRID * MGRID;
static size_t wwords(size_t n_total, size_t frames) { return (n_total >> 5) + frames + 1; }

struct Scratch {            // device scratch for one wave of up to `frames` frames
  float4* rec = 0; float* gz = 0; uint32_t* cnt = 0; float* avg = 0;
  uint4* gsum = 0; uint32_t* slow = 0;   // per (row, 32-column group) summaries; per-frame "use the sweep kernel" flag
  uint32_t* gmask = 0;                   // [frames][ceil(S/32)]: ground_mat == 1 after loop 1, one bit per slot (slot-linear; k_seg_build)
  uint32_t* seg_start = 0; uint16_t* seg_len = 0;   // [frames][SEG_STRIDE] segments bucketed by sector, slot order inside a bucket
  uint32_t* kdesc = 0; uint16_t* act = 0; uint32_t* n_act = 0;   // [frames][NSECT] bucket (base<<16|count), active sectors; [frames]
  uint32_t* owner = 0;                   // [frames][S] claim table, only for range images too large for k_order_winners' shared memory
  uint32_t* occ = 0;                     // [3][frames][ceil(S/32)] slot occupancy bits, contention bits, contended-id prefix
  uint32_t* cwin = 0;                    // [frames][cw_stride] 1 + largest input index per contended slot
  uint32_t* cpt = 0; size_t cpt_words = 0;   // one bit per input point of the wave: the point's slot is contended
  size_t frames = 0;
};

struct Lane {               // host-path double buffer: device staging of inputs/outputs + scratch
  Scratch sc; DevIn in; DevOut out;
  uint8_t* raw = 0; size_t raw_cap = 0;   // packed-record staging (bevgen_process_packed_host), sized on first use
  float* bvm = 0;                         // bird-view-map staging (bevgen_outputs.bvm), allocated on first use
  cudaEvent_t ev_h2d = 0, ev_comp = 0, ev_d2h = 0;
  bool used = false;
};

struct Slot {               // submit/collect ring entry: one frame, own stream, pinned staging
  Scratch sc; DevIn in; DevOut out;
  char* pin_in = 0; char* pin_out = 0; int64_t* offs_d = 0;
  cudaStream_t st = 0; cudaEvent_t done = 0;
  int frame_id = -1; bool busy = false; int n_in = 0;
};

}  // namespace

struct bevgen_ctx {
  int device = 0;
  bevgen_params p{};
  SensorDev sp{};
  Xform xf{};
  int max_pts = 0, max_frames = 0;
  cudaStream_t s_copy = 0, s_comp = 0, s_d2h = 0;
  float* cnt_lut = 0;
  int cw_stride = 1;         // max_points_per_frame / 2 + 1
  int seg_cap = SEG_CAP;     // BEVGEN_SEG_CAP (tests): frames with more segments take the sweep kernel
  int fold_wpb = 1;          // warps per CTA of k_seg_fold (BEVGEN_FOLD_WPB: 1, 2 or 4)
  unsigned long long* diag_d = 0;   // device counters of bevgen_set_diag (borderline pairs, float/double libm disagreements)
  Scratch sc_dev;            // scratch of the device path (waves on the compute stream)
  static constexpr int MAX_AUX = 7;
  Scratch sc_aux[MAX_AUX];       // scratch sets of the auxiliary compute streams
  cudaStream_t s_aux[MAX_AUX]{}; // waves rotate over s_comp, s_aux[0], s_aux[1], ...
  cudaEvent_t ev_fork = 0, ev_join[MAX_AUX]{}; int n_dev_streams = 2;
  int64_t* offs_d = 0; size_t offs_cap = 0;
  char* tmp = 0; size_t tmp_cap = 0;   // grow-only device arena of the host-array entry points (labels, cloud_manip, project, ...)
  bool lanes_ready = false; Lane lanes[3];
  std::vector<Slot> slots;
  // profiling
  bool prof = false;
  cudaEvent_t pev[BEVGEN_N_STAGES + 1]{};
  double stage_ms[BEVGEN_N_STAGES]{};
  int64_t stage_launches[BEVGEN_N_STAGES]{};
  int64_t launches = 0;
};

static const char* kStageNames[BEVGEN_N_STAGES] = {"clear", "order_winners", "order_scatter", "ground_mark",
                                                   "sector_mean", "finalize_bin_scatter", "reserved6", "reserved7"};

// ---- params ---------------------------------------------------------------------------------------------------
extern "C" int bevgen_sensor_params(const char* s, bevgen_params* out) {
  if (!s || !out) return fail("bevgen_sensor_params: null argument");
  memset(out, 0, sizeof *out);
  out->grid_size = BEVGEN_GRID_SIZE; out->max_range = BEVGEN_MAX_RANGE; out->n_layers = BEVGEN_NUM_LAYERS;
  out->lidar_to_ground = 2.0f;
  out->rt[0] = out->rt[5] = out->rt[10] = 1.0f; out->has_transform = 0;
  // src/Utility.cpp:72-89: substring match, tested in this order; :92-124 the table
  if (strstr(s, "HDL_32E")) { out->n_scan = 32; out->horizon_scan = 1056; out->ground_upper_scan = 20; out->height_res = 0.5f; return 0; }
  if (strstr(s, "HDL_64E")) { out->n_scan = 64; out->horizon_scan = 2083; out->ground_upper_scan = 50; out->height_res = 0.25f; return 1; }
  if (strstr(s, "OS1_64")) { out->n_scan = 64; out->horizon_scan = 1024; out->ground_upper_scan = 31; out->height_res = 1.0f; return 2; }
  fail(std::string("Unknown sensor type: ") + s + "!");
  return -1;
}

extern "C" const char* bevgen_last_error(void) { return g_err.c_str(); }
extern "C" const char* bevgen_stage_name(int s) { return (s >= 0 && s < BEVGEN_N_STAGES) ? kStageNames[s] : ""; }

extern "C" void* bevgen_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
  return p;
}
extern "C" void* bevgen_host_alloc_wc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocWriteCombined) != cudaSuccess) { g_err = "cudaHostAlloc (write-combined) failed"; return nullptr; }
  return p;
}
extern "C" void bevgen_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- allocation helpers ---------------------------------------------------------------------------------------
static size_t gsum_per_frame(const SensorDev& sp) { return (size_t)(sp.G + 1) * ((sp.H + 31) / 32); }
static bool fused_order(const SensorDev& sp) { return ord_smem_bytes(sp.S) <= 200 * 1024; }
static int alloc_scratch(Scratch& s, size_t frames, const SensorDev& sp, int max_pts) {
  const size_t S = sp.S;
  s.frames = frames;
  if (!fused_order(sp)) CK(cudaMalloc(&s.owner, frames * S * sizeof(uint32_t)));
  else {
    CK(cudaMalloc(&s.occ, 3 * frames * ((S + 31) / 32) * sizeof(uint32_t)));
    CK(cudaMalloc(&s.cwin, frames * (size_t)(max_pts / 2 + 1) * sizeof(uint32_t)));   // a contended slot holds >= 2 points
    s.cpt_words = frames * (size_t)max_pts / 32 + 4;
    CK(cudaMalloc(&s.cpt, s.cpt_words * sizeof(uint32_t)));
  }
  CK(cudaMalloc(&s.gsum, frames * gsum_per_frame(sp) * sizeof(uint4)));
  CK(cudaMalloc(&s.gmask, frames * ((S + 31) / 32) * sizeof(uint32_t)));
  CK(cudaMalloc(&s.slow, frames * sizeof(uint32_t)));
  CK(cudaMalloc(&s.seg_start, frames * SEG_STRIDE * sizeof(uint32_t))); CK(cudaMalloc(&s.seg_len, frames * SEG_STRIDE * sizeof(uint16_t)));
  CK(cudaMalloc(&s.kdesc, frames * NSECT * sizeof(uint32_t))); CK(cudaMalloc(&s.act, frames * NSECT * sizeof(uint16_t)));
  CK(cudaMalloc(&s.n_act, frames * sizeof(uint32_t)));
  CK(cudaMalloc(&s.rec, frames * S * sizeof(float4)));
  CK(cudaMalloc(&s.gz, (frames * S + 64) * sizeof(float)));   // + slack: k_seg_fold's last aligned window may pass the end
  CK(cudaMalloc(&s.cnt, frames * NSECT * sizeof(uint32_t)));
  CK(cudaMalloc(&s.avg, frames * NSECT * sizeof(float)));
  return 0;
}
static void free_scratch(Scratch& s) {
  cudaFree(s.rec); cudaFree(s.gmask); cudaFree(s.gz); cudaFree(s.cnt); cudaFree(s.avg); cudaFree(s.gsum); cudaFree(s.slow); cudaFree(s.seg_start); cudaFree(s.seg_len); cudaFree(s.kdesc); cudaFree(s.act); cudaFree(s.n_act); cudaFree(s.owner); cudaFree(s.occ); cudaFree(s.cwin); cudaFree(s.cpt);
  s = Scratch();
}
static int alloc_io(DevIn& in, DevOut& out, size_t frames, size_t pts, size_t S) {
  const size_t n = frames * pts;
  CK(cudaMalloc(&in.x, n * 4)); CK(cudaMalloc(&in.y, n * 4)); CK(cudaMalloc(&in.z, n * 4)); CK(cudaMalloc(&in.inten, n * 4));
  CK(cudaMalloc(&in.row, n * 2)); CK(cudaMalloc(&in.col, n * 2)); CK(cudaMalloc(&in.label, n * 2));
  CK(cudaMalloc(&out.label, frames * S * 2)); CK(cudaMalloc(&out.wbits, (wwords(n, frames) + 2) * 4));
  CK(cudaMalloc(&out.single, frames * CELLS)); CK(cudaMalloc(&out.multi, frames * (size_t)LAYERS * CELLS));
  return 0;
}
static void free_io(DevIn& in, DevOut& out) {
  cudaFree(in.x); cudaFree(in.y); cudaFree(in.z); cudaFree(in.inten); cudaFree(in.row); cudaFree(in.col); cudaFree(in.label);
  cudaFree(out.label); cudaFree(out.wbits); cudaFree(out.single); cudaFree(out.multi);
  in = DevIn(); out = DevOut();
}

// ---- create / destroy -----------------------------------------------------------------------------------------
static int create_checks(int device, const bevgen_params* p, int max_pts, int max_frames) {
  if (p->grid_size != BEVGEN_GRID_SIZE || p->max_range != BEVGEN_MAX_RANGE || p->n_layers != BEVGEN_NUM_LAYERS ||
      p->lidar_to_ground != 2.0f)
    return fail("bevgen_create: grid_size/max_range/n_layers/lidar_to_ground must be 224/112/24/2.0 (BatchMultiBevGen.cpp:266-269)");
  if (p->n_scan <= 0 || p->horizon_scan <= 2 || p->ground_upper_scan <= 0 || p->n_scan - p->ground_upper_scan - 1 < 1 ||
      !(p->height_res > 0.0f))
    return fail("bevgen_create: invalid sensor params (need n_scan - ground_upper_scan >= 2)");
  if ((int64_t)p->n_scan * p->horizon_scan > (1 << 24)) return fail("bevgen_create: range image too large");
  if (max_pts <= 0 || max_frames <= 0) return fail("bevgen_create: max_points_per_frame / max_frames_per_batch must be > 0");
  if (max_frames > 65535) return fail("bevgen_create: max_frames_per_batch must be <= 65535 (frames are the y dimension of the launch grids)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("bevgen_create: no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("bevgen_create: bad device index");
  CK(cudaSetDevice(device));
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, k_finalize_bin<false>) != cudaSuccess)
    return fail("bevgen_create: no sm_100a kernel image usable on this device (library is built for B200 only)");
  return 0;
}

// Everything that can fail after the context object exists; on failure bevgen_create destroys the half-built context
// (bevgen_destroy copes with any prefix of this sequence).
static int create_fill(bevgen_ctx* c, int device, const bevgen_params* p, int max_pts, int max_frames) {
  c->device = device; c->p = *p; c->max_pts = max_pts; c->max_frames = max_frames; c->cw_stride = max_pts / 2 + 1;
  SensorDev& sp = c->sp;
  sp.N = p->n_scan; sp.H = p->horizon_scan; sp.G = p->ground_upper_scan; sp.S = sp.N * sp.H;
  sp.band_row0 = sp.N - sp.G - 1;
  sp.height_res = p->height_res; sp.inv_height_res = 1.0f / p->height_res;
  { int e; float m = frexpf(p->height_res, &e); sp.hr_pow2 = (m == 0.5f); }
  // t_star: largest float t with (float)((double)t * 180.0 / M_PI) <= 10.0f  (BatchMultiBevGen.cpp:173,179); bisection on
  // the bit pattern of a monotone function.  A constant of the criterion, not data-dependent compute.
  {
    uint32_t lo, hi; float a = 0.1f, b = 0.2f; memcpy(&lo, &a, 4); memcpy(&hi, &b, 4);
    while (hi - lo > 1) {
      uint32_t mid = lo + (hi - lo) / 2; float t; memcpy(&t, &mid, 4);
      float ang = (float)((double)t * 180.0 / M_PI);
      if (ang <= 10.0f) lo = mid; else hi = mid;
    }
    memcpy(&sp.t_star, &lo, 4);
    double tt = tan((double)sp.t_star);
    sp.q_lo2 = (float)(tt * tt * (1.0 - 2e-5)); sp.q_hi2 = (float)(tt * tt * (1.0 + 2e-5));
  }
  memcpy(c->xf.m, p->rt, sizeof c->xf.m); c->xf.on = p->has_transform ? 1 : 0;

  CK(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  for (auto& e : c->pev) CK(cudaEventCreate(&e));
  CK(cudaFuncSetAttribute(k_finalize_bin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BIN));
  CK(cudaFuncSetAttribute(k_finalize_bin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BIN));
  // records of up to 256 bytes (bevgen_process_packed_host): 256 of them + the alignment head per CTA
  CK(cudaFuncSetAttribute(k_unpack_records<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 256 + 32));
  CK(cudaFuncSetAttribute(k_unpack_records<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 256 + 32));
  CK(cudaFuncSetAttribute(k_seg_build<SEG_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SEG));
  CK(cudaFuncSetAttribute(k_seg_build<SEG_CAP_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_SEG_BIG));
  CK(cudaFuncSetAttribute(k_float_bev, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BVM));
  CK(cudaFuncSetAttribute(k_order_winners<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // function-wide: the largest any context may ask for
  CK(cudaFuncSetAttribute(k_order_winners<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_order_winners<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_order_winners<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_seg_build<SEG_CAP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k_seg_build<SEG_CAP_BIG>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  if (const char* e = getenv("BEVGEN_SEG_CAP")) c->seg_cap = std::max(0, std::min(SEG_CAP, atoi(e)));
  if (const char* e = getenv("BEVGEN_FOLD_WPB")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) c->fold_wpb = v; }
  CK(cudaMalloc(&c->diag_d, 4 * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(c->diag_d, 0, 4 * sizeof(unsigned long long), c->s_comp));
  // (measured: forcing the max-shared carveout on the ordering kernels makes k_order_fill 1.8x slower - it relies on L1)
  CK(cudaFuncSetAttribute(k_sector_mean, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k_seg_fold<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
  CK(cudaFuncSetAttribute(k_seg_fold<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
  CK(cudaFuncSetAttribute(k_finalize_bin<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k_finalize_bin<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaMalloc(&c->cnt_lut, ((size_t)sp.S + 1) * sizeof(float)));
  k_build_cnt_lut<<<1, 32, 0, c->s_comp>>>(sp.S, c->cnt_lut);
  CK(cudaGetLastError());
  c->launches++;
  if (alloc_scratch(c->sc_dev, max_frames, sp, max_pts)) return -1;
  if (const char* e = getenv("BEVGEN_STREAMS")) c->n_dev_streams = std::max(1, std::min(bevgen_ctx::MAX_AUX + 1, atoi(e)));
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i + 1 < c->n_dev_streams; i++) {
    if (alloc_scratch(c->sc_aux[i], max_frames, sp, max_pts)) return -1;
    CK(cudaStreamCreateWithFlags(&c->s_aux[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
  }
  CK(cudaStreamSynchronize(c->s_comp));
  return 0;
}

extern "C" int bevgen_create(bevgen_ctx** out, int device, const bevgen_params* p, int max_pts, int max_frames) {
  if (!out || !p) return fail("bevgen_create: null argument");
  *out = nullptr;
  if (create_checks(device, p, max_pts, max_frames)) return -1;
  bevgen_ctx* c = new bevgen_ctx();
  if (create_fill(c, device, p, max_pts, max_frames)) {
    const std::string why = g_err;     // bevgen_destroy may overwrite the message
    cudaGetLastError();                // clear a sticky allocation error so that the frees below run
    bevgen_destroy(c);
    g_err = why;
    return -1;
  }
  *out = c;
  return 0;
}

static void free_lanes(bevgen_ctx* c) {
  for (auto& l : c->lanes) {
    cudaFree(l.raw); cudaFree(l.bvm); free_scratch(l.sc); free_io(l.in, l.out);
    if (l.ev_h2d) cudaEventDestroy(l.ev_h2d); if (l.ev_comp) cudaEventDestroy(l.ev_comp); if (l.ev_d2h) cudaEventDestroy(l.ev_d2h);
    l = Lane();
  }
  c->lanes_ready = false;
}
static void free_slot(Slot& s) {
  free_scratch(s.sc); free_io(s.in, s.out); cudaFreeHost(s.pin_in); cudaFreeHost(s.pin_out); cudaFree(s.offs_d);
  if (s.st) cudaStreamDestroy(s.st); if (s.done) cudaEventDestroy(s.done);
  s = Slot();
}

extern "C" void bevgen_destroy(bevgen_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_scratch(c->sc_dev);
  for (int i = 0; i < bevgen_ctx::MAX_AUX; i++) {
    free_scratch(c->sc_aux[i]);
    if (c->s_aux[i]) cudaStreamDestroy(c->s_aux[i]);
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  free_lanes(c);
  for (auto& s : c->slots) free_slot(s);
  cudaFree(c->cnt_lut); cudaFree(c->offs_d); cudaFree(c->tmp); cudaFree(c->diag_d);
  for (auto& e : c->pev) if (e) cudaEventDestroy(e);
  if (c->s_copy) cudaStreamDestroy(c->s_copy); if (c->s_comp) cudaStreamDestroy(c->s_comp); if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  delete c;
}

// ---- one wave of frames ----------------------------------------------------------------------------------------
// offs_d: device offsets of the wave (nf+1 entries); `base` is subtracted so that `in` may be a staging buffer that
// starts at the wave's first point.  max_n: largest frame of the wave (host-known).
// The wave is three stages so that the device path can pipeline them across waves:
//   front  = clear + order_claim + order_fill + ground_mark      (L2 / HBM bound)
//   sweep  = sector_mean                                          (MIO / latency bound, few warps per SM)
//   back   = finalize_bin_scatter                                 (HBM + shared-memory bound)
struct WaveArgs { const Scratch* sc; int nf; const int64_t* offs_d; int64_t base; int max_n; DevIn in; DevOut out; int frame0;
                  int64_t qbase, n_pts;   // qbase: 32-aligned caller offset of the wave's first point; n_pts: points of the wave (+ alignment slack)
                  bool compact; };        // compact staging format in (in.inten = u32 meta), compact outputs (out.label = ground bits, out.multi = bit planes)

struct NvtxRange { explicit NvtxRange(const char* n) { nvtxRangePushA(n); } ~NvtxRange() { nvtxRangePop(); } };

static int wave_front(bevgen_ctx* c, cudaStream_t st, const WaveArgs& w, bool prof) {
  NvtxRange nv("bevgen front: order + ground_mark");
  const SensorDev& sp = c->sp;
  const size_t S = sp.S;
  auto mark = [&](int i) { if (prof) cudaEventRecord(c->pev[i], st); };
  // inputs are indexed with the caller's offsets: shift the pointers instead of the offsets
  const float *x = w.in.x - w.base, *y = w.in.y - w.base, *z = w.in.z - w.base, *it = w.in.inten - w.base;
  const uint16_t *row = w.in.row - w.base, *col = w.in.col - w.base;
  const int16_t* lab = w.in.label - w.base;
  mark(0);
  CK(cudaMemsetAsync(w.sc->cnt, 0, (size_t)w.nf * NSECT * sizeof(uint32_t), st));
  if (!fused_order(sp)) CK(cudaMemsetAsync(w.sc->owner, 0, (size_t)w.nf * S * sizeof(uint32_t), st));
  else CK(cudaMemsetAsync(w.sc->cpt, 0, std::min(w.sc->cpt_words, (size_t)(w.n_pts + 31) / 32 + 2) * sizeof(uint32_t), st));
  mark(1);
  if (fused_order(sp)) {
    const size_t W = (S + 31) / 32;
    uint32_t *occ = w.sc->occ, *cont = occ + (size_t)w.sc->frames * W, *cpre = cont + (size_t)w.sc->frames * W;
    dim3 g((std::max<int>(w.max_n, (int)S) + SCAT_T * SCAT_PPT - 1) / (SCAT_T * SCAT_PPT), w.nf);
    if (w.compact) {
      const uint16_t* meta = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint32_t*>(w.in.inten) - w.base);
      if (ord_needs_swizzle(sp.H)) k_order_winners<true, true><<<w.nf, ORD_T, ord_smem_bytes(sp.S), st>>>(sp, w.offs_d, c->cw_stride, meta, nullptr, occ, cont, cpre, w.sc->cwin, w.qbase, w.sc->cpt);
      else k_order_winners<true, false><<<w.nf, ORD_T, ord_smem_bytes(sp.S), st>>>(sp, w.offs_d, c->cw_stride, meta, nullptr, occ, cont, cpre, w.sc->cwin, w.qbase, w.sc->cpt);
      mark(2);
      k_order_scatter<true><<<g, SCAT_T, 0, st>>>(sp, c->xf, w.offs_d, w.frame0, c->cw_stride, x, y, z, it, nullptr, nullptr, nullptr, occ, cont, cpre, w.sc->cwin,
                                                 w.sc->rec, w.out.wbits, w.qbase, w.sc->cpt);
    } else {
      if (ord_needs_swizzle(sp.H)) k_order_winners<false, true><<<w.nf, ORD_T, ord_smem_bytes(sp.S), st>>>(sp, w.offs_d, c->cw_stride, row, col, occ, cont, cpre, w.sc->cwin, w.qbase, w.sc->cpt);
      else k_order_winners<false, false><<<w.nf, ORD_T, ord_smem_bytes(sp.S), st>>>(sp, w.offs_d, c->cw_stride, row, col, occ, cont, cpre, w.sc->cwin, w.qbase, w.sc->cpt);
      mark(2);
      k_order_scatter<false><<<g, SCAT_T, 0, st>>>(sp, c->xf, w.offs_d, w.frame0, c->cw_stride, x, y, z, it, row, col, lab, occ, cont, cpre, w.sc->cwin,
                                                  w.sc->rec, w.out.wbits, w.qbase, w.sc->cpt);
    }
    c->launches += 2;
  } else {   // range image too large for shared memory: claim table in global memory (two kernels + the winner bits)
    if (w.compact) return fail("compact staging format: range image too large (needs the shared-memory ordering kernels)");
    if (w.max_n > 0) {
      dim3 g((w.max_n + 511) / 512, w.nf);
      k_order_claim<<<g, 256, 0, st>>>(sp, w.offs_d, row, col, w.sc->owner);
    }
    mark(2);
    dim3 g((std::max<int>(w.max_n, (int)S) + 255) / 256, w.nf);
    k_order_fill<<<g, 256, 0, st>>>(sp, c->xf, w.offs_d, x, y, z, it, row, col, lab, w.sc->owner, w.sc->rec);
    if (w.max_n > 0) {
      dim3 g2((w.max_n + 255) / 256, w.nf);
      k_winner_bits<<<g2, 256, 0, st>>>(sp, w.offs_d, w.frame0, row, col, w.sc->owner, w.out.wbits);
    }
    c->launches += (w.max_n > 0 ? 3 : 1);
  }
  mark(3);
  {
    dim3 g((sp.H + GM_T - 1) / GM_T, w.nf);
    if (sp.libm_double || sp.diag) k_ground_mark<true><<<g, GM_T, 0, st>>>(sp, w.sc->rec, w.sc->gz, w.sc->cnt, w.sc->gsum);
    else k_ground_mark<false><<<g, GM_T, 0, st>>>(sp, w.sc->rec, w.sc->gz, w.sc->cnt, w.sc->gsum);
  }
  mark(4);
  CK(cudaGetLastError());
  c->launches += 1;
  return 0;
}
static int wave_sweep(bevgen_ctx* c, cudaStream_t st, const WaveArgs& w, bool prof) {
  NvtxRange nv("bevgen sweep: sector means");
  // segment form first: every frame with up to seg_cap segments, then (one CTA per SM, twice the shared memory) the frames with
  // up to SEG_CAP_BIG; what neither holds raises slow[f] = 1 and is swept by k_sector_mean
  k_seg_build<SEG_CAP><<<w.nf, SEGT, SMEM_SEG, st>>>(c->sp, c->seg_cap, w.sc->gsum, w.sc->rec, w.sc->avg, w.sc->slow, w.sc->seg_start,
                                                     w.sc->seg_len, w.sc->kdesc, w.sc->act, w.sc->n_act, w.sc->gmask, w.sc->cnt);
  k_seg_build<SEG_CAP_BIG><<<w.nf, SEGT, SMEM_SEG_BIG, st>>>(c->sp, SEG_CAP_BIG, w.sc->gsum, w.sc->rec, w.sc->avg, w.sc->slow, w.sc->seg_start,
                                                             w.sc->seg_len, w.sc->kdesc, w.sc->act, w.sc->n_act, w.sc->gmask, w.sc->cnt);
  const dim3 fg(w.nf, FOLD_PASSES / c->fold_wpb), fb(32, c->fold_wpb);
  if ((c->sp.S & 7) == 0)   // every frame of gz starts on a 32-byte boundary: 256-bit loads
    k_seg_fold<true><<<fg, fb, 0, st>>>(c->sp, w.sc->gz, w.sc->cnt, c->cnt_lut, w.sc->slow, w.sc->seg_start, w.sc->seg_len,
                                        w.sc->kdesc, w.sc->act, w.sc->n_act, w.sc->avg);
  else
    k_seg_fold<false><<<fg, fb, 0, st>>>(c->sp, w.sc->gz, w.sc->cnt, c->cnt_lut, w.sc->slow, w.sc->seg_start, w.sc->seg_len,
                                         w.sc->kdesc, w.sc->act, w.sc->n_act, w.sc->avg);
  k_sector_mean<<<w.nf, 32, 2 * NSECT * sizeof(float), st>>>(c->sp, w.sc->rec, w.sc->gz, w.sc->cnt, c->cnt_lut, w.sc->avg, w.sc->slow);
  if (prof) cudaEventRecord(c->pev[5], st);
  CK(cudaGetLastError());
  c->launches += 4;
  return 0;
}
static int wave_back(bevgen_ctx* c, cudaStream_t st, const WaveArgs& w, bool prof) {
  NvtxRange nv("bevgen back: finalize + bin + scatter");
  if (w.compact) k_finalize_bin<true><<<w.nf, 1024, SMEM_BIN, st>>>(c->sp, w.sc->rec, w.sc->gmask, w.sc->avg, w.out.label, w.out.single, w.out.multi);
  else k_finalize_bin<false><<<w.nf, 1024, SMEM_BIN, st>>>(c->sp, w.sc->rec, w.sc->gmask, w.sc->avg, w.out.label, w.out.single, w.out.multi);
  if (prof) cudaEventRecord(c->pev[6], st);
  CK(cudaGetLastError());
  c->launches += 1;
  if (w.out.bvm) {   // batch_cloud_manip's float bird-view map, from the labels just written
    k_float_bev<<<w.nf, 512, SMEM_BVM, st>>>(c->sp, w.sc->rec, w.out.label, w.out.bvm);
    CK(cudaGetLastError());
    c->launches += 1;
  }
  if (prof) {
    CK(cudaEventSynchronize(c->pev[6]));
    for (int i = 0; i < 6; i++) {
      float ms = 0; CK(cudaEventElapsedTime(&ms, c->pev[i], c->pev[i + 1]));
      c->stage_ms[i] += ms;
      c->stage_launches[i] += 1;
    }
  }
  return 0;
}
// all three stages back to back on one stream
static int run_wave(bevgen_ctx* c, cudaStream_t st, const Scratch& sc, int nf, const int64_t* offs_d, int64_t base, int max_n,
                    const DevIn& in, const DevOut& out, bool prof, int frame0, int64_t first_pt, int64_t end_pt, bool compact = false) {
  const int64_t qbase = first_pt & ~(int64_t)31;
  WaveArgs w{&sc, nf, offs_d, base, max_n, in, out, frame0, qbase, end_pt - qbase, compact};
  if (wave_front(c, st, w, prof)) return -1;
  if (wave_sweep(c, st, w, prof)) return -1;
  return wave_back(c, st, w, prof);
}

static int upload_offsets(bevgen_ctx* c, int nf, const int64_t* offsets, cudaStream_t st, int* max_n_out) {
  if ((size_t)nf + 1 > c->offs_cap) {
    CK(cudaStreamSynchronize(c->s_comp)); CK(cudaStreamSynchronize(c->s_copy));
    cudaFree(c->offs_d); c->offs_d = 0;
    c->offs_cap = (size_t)nf + 1 + 1024;
    CK(cudaMalloc(&c->offs_d, c->offs_cap * sizeof(int64_t)));
  }
  int64_t mx = 0;
  for (int f = 0; f < nf; f++) {
    int64_t n = offsets[f + 1] - offsets[f];
    if (n < 0) return fail("offsets must be non-decreasing");
    if (n > 0x7ffffff0) return fail("frame too large");
    mx = std::max(mx, n);
  }
  *max_n_out = (int)mx;
  CK(cudaMemcpyAsync(c->offs_d, offsets, ((size_t)nf + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  return 0;
}

// ---- device arena of the host-array entry points ----------------------------------------------------------------
// One grow-only allocation per context instead of cudaMalloc / cudaFree per call (both synchronise the device and cost
// far more than the kernels of these small entry points); carved into 256-byte aligned pieces.
struct Carver {
  char* p; size_t used = 0;
  static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
  template <typename T> T* take(size_t n) { T* r = reinterpret_cast<T*>(p + used); used += pad(n * sizeof(T)); return r; }
};
static int tmp_reserve(bevgen_ctx* c, size_t bytes) {
  if (bytes <= c->tmp_cap) return 0;
  CK(cudaStreamSynchronize(c->s_comp));
  cudaFree(c->tmp); c->tmp = 0; c->tmp_cap = 0;
  const size_t cap = bytes + bytes / 4 + 4096;
  CK(cudaMalloc(&c->tmp, cap));
  c->tmp_cap = cap;
  return 0;
}

// ---- device-resident path -------------------------------------------------------------------------------------
// Every array of the two structs is required (a NULL would only surface as a fault inside a kernel); bvm alone is optional,
// and a batch without a single point may come with NULL point arrays.
static bool null_arrays(const bevgen_points* in, const bevgen_outputs* out, bool any_points = true) {
  if (in && any_points && (!in->x || !in->y || !in->z || !in->intensity || !in->row || !in->col || !in->label)) return true;
  return !out->label || !out->winner_bits || !out->single_bev || !out->multi_bev;
}
extern "C" int bevgen_process_device(bevgen_ctx* c, int nf, const int64_t* offsets, const bevgen_points* in, const bevgen_outputs* out) {
  if (!c || !offsets || !in || !out) return fail("bevgen_process_device: null argument");
  if (nf <= 0) return 0;
  if (null_arrays(in, out, offsets[nf] > offsets[0])) return fail("bevgen_process_device: null array (only bevgen_outputs.bvm is optional)");
  CK(cudaSetDevice(c->device));
  int max_n_all = 0;
  if (upload_offsets(c, nf, offsets, c->s_comp, &max_n_all)) return -1;
  if (max_n_all > c->max_pts) return fail("bevgen_process_device: a frame exceeds max_points_per_frame");
  const size_t S = c->sp.S;
  DevIn di; di.x = (float*)in->x; di.y = (float*)in->y; di.z = (float*)in->z; di.inten = (float*)in->intensity;
  di.row = (uint16_t*)in->row; di.col = (uint16_t*)in->col; di.label = (int16_t*)in->label;
  // Consecutive waves alternate between the compute stream and an auxiliary stream (own scratch set), so kernels of
  // two waves interleave on the SMs (e.g. the latency-bound sweep of one wave with the ordering kernels of the other).
  // Measured alternatives (profiles/r1_notes.md): a front/sweep/back software pipeline with a high-priority sweep
  // stream was slower - sweep and ordering kernels contend for the same L1/LSU data pipe.  Profiling serialises.
  const int nw = (nf + c->max_frames - 1) / c->max_frames;
  const int ns = (c->prof || nw == 1) ? 1 : std::min(c->n_dev_streams, nw);
  if (ns > 1) { CK(cudaEventRecord(c->ev_fork, c->s_comp)); for (int i = 0; i + 1 < ns; i++) CK(cudaStreamWaitEvent(c->s_aux[i], c->ev_fork, 0)); }
  for (int w = 0; w < nw; w++) {
    const int f0 = w * c->max_frames;
    const int n = std::min(c->max_frames, nf - f0);
    int max_n = 0;
    for (int f = f0; f < f0 + n; f++) max_n = std::max<int64_t>(max_n, offsets[f + 1] - offsets[f]);
    DevOut dout; dout.label = out->label + (size_t)f0 * S; dout.wbits = out->winner_bits;   // words are indexed by absolute offsets
    dout.single = out->single_bev + (size_t)f0 * CELLS; dout.multi = out->multi_bev + (size_t)f0 * LAYERS * CELLS;
    dout.bvm = out->bvm ? out->bvm + (size_t)f0 * BVM_CELLS : nullptr;
    const int si = w % ns;   // stream 0 = s_comp, i > 0 = s_aux[i - 1]; each stream owns one scratch set, its waves serialise on it
    if (run_wave(c, si ? c->s_aux[si - 1] : c->s_comp, si ? c->sc_aux[si - 1] : c->sc_dev, n, c->offs_d + f0, 0, max_n, di, dout, c->prof, f0, offsets[f0], offsets[f0 + n])) return -1;
  }
  for (int i = 0; i + 1 < ns; i++) { CK(cudaEventRecord(c->ev_join[i], c->s_aux[i])); CK(cudaStreamWaitEvent(c->s_comp, c->ev_join[i], 0)); }
  return 0;
}

extern "C" int bevgen_sync(bevgen_ctx* c) {
  if (!c) return fail("bevgen_sync: null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->s_copy)); CK(cudaStreamSynchronize(c->s_comp)); CK(cudaStreamSynchronize(c->s_d2h));
  return 0;
}

// ---- host-buffer path: H2D (copy stream) | kernels (compute stream) | D2H (third stream), double-buffered ------
// The host path is PCIe-bound, so its chunks are kept small enough that H2D of chunk k+1, the kernels of chunk k and
// D2H of chunk k-1 overlap (the device path uses max_frames_per_batch-sized waves instead).
static int host_chunk(const bevgen_ctx* c) {
  static const int pref = [] { const char* e = getenv("BEVGEN_HOST_CHUNK"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 48; }();
  return std::min(c->max_frames, pref);
}

static int alloc_lanes(bevgen_ctx* c) {
  for (auto& l : c->lanes) {
    if (alloc_scratch(l.sc, host_chunk(c), c->sp, c->max_pts)) return -1;
    if (alloc_io(l.in, l.out, host_chunk(c), c->max_pts, c->sp.S)) return -1;
    CK(cudaEventCreateWithFlags(&l.ev_h2d, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&l.ev_comp, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&l.ev_d2h, cudaEventDisableTiming));
  }
  return 0;
}
static int ensure_lanes(bevgen_ctx* c) {
  if (c->lanes_ready) return 0;
  if (alloc_lanes(c)) {             // all or nothing: a later call starts from scratch instead of reusing half-built lanes
    const std::string why = g_err;
    cudaGetLastError();
    free_lanes(c);
    g_err = why;
    return -1;
  }
  c->lanes_ready = true;
  return 0;
}

// `in` (SoA) or `records` + `lay` (interleaved records, de-interleaved on the GPU by k_unpack_records) or `cin` (compact
// staging format) - exactly one is set; `out` (reference-layout outputs) or `cout` (compact outputs, only with `cin`).
static int process_host_impl(bevgen_ctx* c, int nf, const int64_t* offsets, const bevgen_points* in, const uint8_t* records,
                             const RecLayout* lay, const bevgen_outputs* out, const bevgen_points_compact* cin = nullptr,
                             const bevgen_outputs_compact* cout = nullptr) {
  static const bevgen_outputs no_out = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const bool compact = cin != nullptr;
  if (compact) out = &no_out;
  CK(cudaSetDevice(c->device));
  if (ensure_lanes(c)) return -1;
  int max_n_all = 0;
  if (upload_offsets(c, nf, offsets, c->s_comp, &max_n_all)) return -1;
  if (max_n_all > c->max_pts) return fail("bevgen_process_host: a frame exceeds max_points_per_frame");
  const size_t S = c->sp.S;
  const int chunk = host_chunk(c);
  if (records) {
    const size_t need = (size_t)chunk * c->max_pts * lay->stride + 64;
    for (auto& l : c->lanes)
      if (l.raw_cap < need) {
        CK(cudaStreamSynchronize(c->s_comp)); CK(cudaStreamSynchronize(c->s_copy));
        cudaFree(l.raw); l.raw = 0; l.raw_cap = 0;
        CK(cudaMalloc(&l.raw, need)); l.raw_cap = need;
      }
  }
  if (out->bvm)
    for (auto& l : c->lanes) if (!l.bvm) CK(cudaMalloc(&l.bvm, (size_t)chunk * BVM_CELLS * sizeof(float)));
  int k = 0;
  for (int f0 = 0; f0 < nf; f0 += chunk, k++) {
    Lane& l = c->lanes[k % 3];
    const int n = std::min(chunk, nf - f0);
    const int64_t base = offsets[f0];
    const size_t np = (size_t)(offsets[f0 + n] - base);
    int max_n = 0;
    for (int f = f0; f < f0 + n; f++) max_n = std::max<int64_t>(max_n, offsets[f + 1] - offsets[f]);
    // inputs of this lane may be overwritten once the kernels of its previous wave are done
    if (l.used) CK(cudaStreamWaitEvent(c->s_copy, l.ev_comp, 0));
    {
    NvtxRange nv_h2d("bevgen H2D enqueue");   // scoped: an early CK return may not leave the range open
    if (records) {
      CK(cudaMemcpyAsync(l.raw, records + (size_t)base * lay->stride, np * lay->stride, cudaMemcpyHostToDevice, c->s_copy));
    } else if (compact) {   // 16 bytes per point: x, y, z + (slot | flags); l.in.inten holds the meta words
      CK(cudaMemcpyAsync(l.in.x, cin->x + base, np * 4, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.y, cin->y + base, np * 4, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.z, cin->z + base, np * 4, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.inten, cin->meta + base, np * 4, cudaMemcpyHostToDevice, c->s_copy));
    } else {
      CK(cudaMemcpyAsync(l.in.x, in->x + base, np * 4, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.y, in->y + base, np * 4, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.z, in->z + base, np * 4, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.inten, in->intensity + base, np * 4, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.row, in->row + base, np * 2, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.col, in->col + base, np * 2, cudaMemcpyHostToDevice, c->s_copy));
      CK(cudaMemcpyAsync(l.in.label, in->label + base, np * 2, cudaMemcpyHostToDevice, c->s_copy));
    }
    }
    CK(cudaEventRecord(l.ev_h2d, c->s_copy));
    CK(cudaStreamWaitEvent(c->s_comp, l.ev_h2d, 0));
    if (l.used) CK(cudaStreamWaitEvent(c->s_comp, l.ev_d2h, 0));   // outputs of the previous wave have left
    if (records && max_n > 0) {   // interleaved records -> the lane's SoA arrays
      bool even = (lay->stride & 1) == 0;
      for (int q = 0; q < 7; q++) if (lay->off[q] >= 0 && (lay->off[q] & 1)) even = false;
      dim3 g((max_n + 255) / 256, n);
      const size_t sm = 256 * (size_t)lay->stride + 32;
      if (even) k_unpack_records<true><<<g, 256, sm, c->s_comp>>>(*lay, c->offs_d + f0, base, l.raw, l.in.x, l.in.y, l.in.z, l.in.inten, l.in.row, l.in.col, l.in.label);
      else k_unpack_records<false><<<g, 256, sm, c->s_comp>>>(*lay, c->offs_d + f0, base, l.raw, l.in.x, l.in.y, l.in.z, l.in.inten, l.in.row, l.in.col, l.in.label);
      CK(cudaGetLastError());
      c->launches++;
    }
    // winner words of this chunk: [w0, w1) of the caller's array; the kernel indexes with absolute offsets / frame ids
    const size_t w0 = (size_t)(base >> 5) + (size_t)f0, w1 = (size_t)(offsets[f0 + n] >> 5) + (size_t)(f0 + n);
    DevOut lo = l.out; lo.wbits = l.out.wbits - w0; lo.bvm = out->bvm ? l.bvm : nullptr;
    if (run_wave(c, c->s_comp, l.sc, n, c->offs_d + f0, base, max_n, l.in, lo, false, f0, offsets[f0], offsets[f0 + n], compact)) return -1;
    CK(cudaEventRecord(l.ev_comp, c->s_comp));
    CK(cudaStreamWaitEvent(c->s_d2h, l.ev_comp, 0));
    {
    NvtxRange nv_d2h("bevgen D2H enqueue");
    if (compact) {   // ground bits (in the lane's label buffer), winner bits, single, the three occupancy bit planes
      const size_t GW = (S + 31) / 32;
      CK(cudaMemcpyAsync(cout->ground_bits + (size_t)f0 * GW, l.out.label, (size_t)n * GW * 4, cudaMemcpyDeviceToHost, c->s_d2h));
      CK(cudaMemcpyAsync(cout->winner_bits + w0, l.out.wbits, (w1 - w0) * 4, cudaMemcpyDeviceToHost, c->s_d2h));
      CK(cudaMemcpyAsync(cout->single_bev + (size_t)f0 * CELLS, l.out.single, (size_t)n * CELLS, cudaMemcpyDeviceToHost, c->s_d2h));
      CK(cudaMemcpyAsync(cout->multi_planes + (size_t)f0 * 3 * CELLS, l.out.multi, (size_t)n * 3 * CELLS, cudaMemcpyDeviceToHost, c->s_d2h));
    } else {
      CK(cudaMemcpyAsync(out->label + (size_t)f0 * S, l.out.label, (size_t)n * S * 2, cudaMemcpyDeviceToHost, c->s_d2h));
      CK(cudaMemcpyAsync(out->winner_bits + w0, l.out.wbits, (w1 - w0) * 4, cudaMemcpyDeviceToHost, c->s_d2h));
      CK(cudaMemcpyAsync(out->single_bev + (size_t)f0 * CELLS, l.out.single, (size_t)n * CELLS, cudaMemcpyDeviceToHost, c->s_d2h));
      CK(cudaMemcpyAsync(out->multi_bev + (size_t)f0 * LAYERS * CELLS, l.out.multi, (size_t)n * LAYERS * CELLS, cudaMemcpyDeviceToHost, c->s_d2h));
      if (out->bvm) CK(cudaMemcpyAsync(out->bvm + (size_t)f0 * BVM_CELLS, l.bvm, (size_t)n * BVM_CELLS * sizeof(float), cudaMemcpyDeviceToHost, c->s_d2h));
    }
    }
    CK(cudaEventRecord(l.ev_d2h, c->s_d2h));
    l.used = true;
  }
  CK(cudaStreamSynchronize(c->s_d2h));
  CK(cudaStreamSynchronize(c->s_comp));
  CK(cudaStreamSynchronize(c->s_copy));
  return 0;
}

extern "C" int bevgen_process_host(bevgen_ctx* c, int nf, const int64_t* offsets, const bevgen_points* in, const bevgen_outputs* out) {
  if (!c || !offsets || !in || !out) return fail("bevgen_process_host: null argument");
  if (nf <= 0) return 0;
  if (null_arrays(in, out, offsets[nf] > offsets[0])) return fail("bevgen_process_host: null array (only bevgen_outputs.bvm is optional)");
  return process_host_impl(c, nf, offsets, in, nullptr, nullptr, out);
}

extern "C" int bevgen_process_host_compact(bevgen_ctx* c, int nf, const int64_t* offsets, const bevgen_points_compact* in,
                                           const bevgen_outputs_compact* out) {
  if (!c || !offsets || !in || !out) return fail("bevgen_process_host_compact: null argument");
  if (!in->x || !in->y || !in->z || !in->meta || !out->ground_bits || !out->winner_bits || !out->single_bev || !out->multi_planes)
    return fail("bevgen_process_host_compact: null array");
  if ((unsigned)c->sp.S >= BEVGEN_META_INVALID) return fail("bevgen_process_host_compact: range image too large for a 24-bit slot index");
  if (nf <= 0) return 0;
  return process_host_impl(c, nf, offsets, nullptr, nullptr, nullptr, nullptr, in, out);
}

// The 26-byte record savePCDFileBinary writes for pcl::PointXYZIRCT (BatchMultiBevGen.h:56-66; `t` at byte 20 is skipped).
extern "C" int bevgen_pcd_record_layout(bevgen_record_layout* out) {
  if (!out) return fail("bevgen_pcd_record_layout: null argument");
  out->stride = 26; out->off_x = 0; out->off_y = 4; out->off_z = 8; out->off_intensity = 12;
  out->off_row = 16; out->off_col = 18; out->off_label = 24;
  return 0;
}

extern "C" int bevgen_process_packed_host(bevgen_ctx* c, int nf, const int64_t* offsets, const void* records,
                                          const bevgen_record_layout* layout, const bevgen_outputs* out) {
  if (!c || !offsets || !records || !layout || !out) return fail("bevgen_process_packed_host: null argument");
  if (nf <= 0) return 0;
  if (null_arrays(nullptr, out)) return fail("bevgen_process_packed_host: null array (only bevgen_outputs.bvm is optional)");
  RecLayout L;
  L.stride = layout->stride;
  const int offs[7] = {layout->off_x, layout->off_y, layout->off_z, layout->off_intensity, layout->off_row, layout->off_col, layout->off_label};
  if (L.stride < 1 || L.stride > 256) return fail("bevgen_process_packed_host: record stride must be 1..256 bytes");
  for (int q = 0; q < 7; q++) {
    const int w = q < 4 ? 4 : 2;
    if (offs[q] < -1 || (offs[q] >= 0 && offs[q] + w > L.stride)) return fail("bevgen_process_packed_host: field offset outside the record");
    L.off[q] = offs[q];
  }
  return process_host_impl(c, nf, offsets, nullptr, (const uint8_t*)records, &L, out);
}

// ---- submit / collect (single frames in flight, one stream per ring slot) ---------------------------------------
static size_t in_bytes(size_t n) { return n * 22; }
// The ring holds up to max_frames_per_batch frames in flight (bevgen.h); a slot (one frame of scratch + staging + its own
// stream) is only built when every existing one is busy.
static int build_slot(bevgen_ctx* c, Slot& s) {
  const size_t S = c->sp.S;
  if (alloc_scratch(s.sc, 1, c->sp, c->max_pts)) return -1;
  if (alloc_io(s.in, s.out, 1, c->max_pts, S)) return -1;
  CK(cudaHostAlloc((void**)&s.pin_in, in_bytes(c->max_pts) + 16, cudaHostAllocPortable));
  CK(cudaHostAlloc((void**)&s.pin_out, ((size_t)c->max_pts / 32 + 4) * 4 + S * 2 + CELLS + (size_t)LAYERS * CELLS, cudaHostAllocPortable));
  CK(cudaMalloc(&s.offs_d, 2 * sizeof(int64_t)));
  CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  return 0;
}
static Slot* free_ring_slot(bevgen_ctx* c) {
  for (auto& t : c->slots) if (!t.busy) return &t;
  if ((int)c->slots.size() >= c->max_frames) { fail("bevgen_submit: ring full — collect a frame first"); return nullptr; }
  c->slots.emplace_back();
  if (build_slot(c, c->slots.back())) {
    const std::string why = g_err;
    cudaGetLastError();
    free_slot(c->slots.back());
    c->slots.pop_back();
    g_err = why;
    return nullptr;
  }
  return &c->slots.back();
}

extern "C" int bevgen_submit(bevgen_ctx* c, int frame_id, int n_in, const float* x, const float* y, const float* z,
                             const float* intensity, const uint16_t* row, const uint16_t* col, const int16_t* label) {
  if (!c) return fail("bevgen_submit: null ctx");
  if (n_in < 0 || n_in > c->max_pts) return fail("bevgen_submit: n_in exceeds max_points_per_frame");
  if (n_in > 0 && (!x || !y || !z || !intensity || !row || !col || !label)) return fail("bevgen_submit: null point array");
  CK(cudaSetDevice(c->device));
  for (auto& t : c->slots) { if (t.busy && t.frame_id == frame_id) return fail("bevgen_submit: frame_id already in flight"); }
  Slot* s = free_ring_slot(c);
  if (!s) return -1;
  const size_t n = (size_t)n_in, S = c->sp.S;
  // pinned staging: [x|y|z|intensity|row|col|label] SoA, then the two offsets
  char* p = s->pin_in;
  float* px = (float*)p; float* py = px + n; float* pz = py + n; float* pi = pz + n;
  uint16_t* pr = (uint16_t*)(pi + n); uint16_t* pc = pr + n; int16_t* pl = (int16_t*)(pc + n);
  if (n) {   // an empty frame may come with NULL arrays
    memcpy(px, x, n * 4); memcpy(py, y, n * 4); memcpy(pz, z, n * 4); memcpy(pi, intensity, n * 4);
    memcpy(pr, row, n * 2); memcpy(pc, col, n * 2); memcpy(pl, label, n * 2);
  }
  int64_t* po = (int64_t*)(s->pin_out);   // offsets live at the head of pin_out until the kernels ran
  po[0] = 0; po[1] = n_in;
  CK(cudaMemcpyAsync(s->offs_d, po, 16, cudaMemcpyHostToDevice, s->st));
  CK(cudaMemcpyAsync(s->in.x, px, n * 4, cudaMemcpyHostToDevice, s->st));
  CK(cudaMemcpyAsync(s->in.y, py, n * 4, cudaMemcpyHostToDevice, s->st));
  CK(cudaMemcpyAsync(s->in.z, pz, n * 4, cudaMemcpyHostToDevice, s->st));
  CK(cudaMemcpyAsync(s->in.inten, pi, n * 4, cudaMemcpyHostToDevice, s->st));
  CK(cudaMemcpyAsync(s->in.row, pr, n * 2, cudaMemcpyHostToDevice, s->st));
  CK(cudaMemcpyAsync(s->in.col, pc, n * 2, cudaMemcpyHostToDevice, s->st));
  CK(cudaMemcpyAsync(s->in.label, pl, n * 2, cudaMemcpyHostToDevice, s->st));
  if (run_wave(c, s->st, s->sc, 1, s->offs_d, 0, n_in, s->in, s->out, false, 0, 0, n_in)) return -1;
  char* q = s->pin_out;
  const size_t ww = ((size_t)c->max_pts / 32 + 4) * 4;
  CK(cudaMemcpyAsync(q, s->out.wbits, ((n + 31) / 32) * 4, cudaMemcpyDeviceToHost, s->st)); q += ww;
  CK(cudaMemcpyAsync(q, s->out.label, S * 2, cudaMemcpyDeviceToHost, s->st)); q += S * 2;
  CK(cudaMemcpyAsync(q, s->out.single, CELLS, cudaMemcpyDeviceToHost, s->st)); q += CELLS;
  CK(cudaMemcpyAsync(q, s->out.multi, (size_t)LAYERS * CELLS, cudaMemcpyDeviceToHost, s->st));
  CK(cudaEventRecord(s->done, s->st));
  s->busy = true; s->frame_id = frame_id; s->n_in = n_in;
  return 0;
}

extern "C" int bevgen_collect(bevgen_ctx* c, int frame_id, int16_t* label_out, uint32_t* winner_bits, uint8_t* single_bev, uint8_t* multi_bev) {
  if (!c) return fail("bevgen_collect: null ctx");
  CK(cudaSetDevice(c->device));
  for (auto& s : c->slots) {
    if (!s.busy || s.frame_id != frame_id) continue;
    CK(cudaEventSynchronize(s.done));
    const size_t S = c->sp.S;
    const char* q = s.pin_out;
    if (winner_bits) memcpy(winner_bits, q, (((size_t)s.n_in + 31) / 32) * 4);
    q += ((size_t)c->max_pts / 32 + 4) * 4;
    if (label_out) memcpy(label_out, q, S * 2); q += S * 2;
    if (single_bev) memcpy(single_bev, q, CELLS); q += CELLS;
    if (multi_bev) memcpy(multi_bev, q, (size_t)LAYERS * CELLS);
    s.busy = false; s.frame_id = -1;
    return 0;
  }
  return fail("bevgen_collect: frame_id not in flight");
}

// ---- labels ---------------------------------------------------------------------------------------------------
namespace bevgen {
__global__ void k_gather_mpos(int M, const float* __restrict__ xyz, const int32_t* __restrict__ major_idx, float* __restrict__ mpos) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < M) { int i = major_idx[j]; mpos[3 * j] = xyz[3 * i]; mpos[3 * j + 1] = xyz[3 * i + 1]; mpos[3 * j + 2] = xyz[3 * i + 2]; }
}
}  // namespace bevgen

extern "C" int bevgen_select_major(bevgen_ctx* c, int K, const float* xyz, int32_t* major_idx, int32_t* n_major, int32_t* overlap_nn) {
  if (!c || !xyz || !major_idx || !n_major) return fail("bevgen_select_major: null argument");
  if (K <= 0) { *n_major = 0; return 0; }
  CK(cudaSetDevice(c->device));
  if (tmp_reserve(c, 2 * Carver::pad((size_t)K * 12) + 2 * Carver::pad((size_t)K * 4) + 256)) return -1;
  Carver cv{c->tmp};
  float* d_xyz = cv.take<float>((size_t)K * 3); float* d_mpos = cv.take<float>((size_t)K * 3);
  int32_t* d_mi = cv.take<int32_t>(K); int32_t* d_ov = cv.take<int32_t>(K); int32_t* d_n = cv.take<int32_t>(1);
  CK(cudaMemcpyAsync(d_xyz, xyz, (size_t)K * 12, cudaMemcpyHostToDevice, c->s_comp));
  k_select_major<<<1, 32, 0, c->s_comp>>>(K, d_xyz, d_mpos, d_mi, d_ov, d_n);
  CK(cudaGetLastError()); c->launches++;
  int32_t M = 0;
  CK(cudaMemcpyAsync(&M, d_n, 4, cudaMemcpyDeviceToHost, c->s_comp));
  CK(cudaStreamSynchronize(c->s_comp));
  CK(cudaMemcpy(major_idx, d_mi, (size_t)M * 4, cudaMemcpyDeviceToHost));
  if (overlap_nn) CK(cudaMemcpy(overlap_nn, d_ov, (size_t)K * 4, cudaMemcpyDeviceToHost));
  *n_major = M;
  return 0;
}

extern "C" int bevgen_labels(bevgen_ctx* c, int K, const float* xyz, int M, const int32_t* major_idx, int row_begin, int row_end,
                             float* labels_out, int32_t* nn_idx, float* nn_w) {
  if (!c || !xyz || !major_idx) return fail("bevgen_labels: null argument");
  if (K <= 0 || M <= 0) return fail("bevgen_labels: K and M must be > 0");
  if (row_begin < 0 || row_end > K || row_begin > row_end) return fail("bevgen_labels: bad row range");
  for (int j = 0; j < M; j++) if (major_idx[j] < 0 || major_idx[j] >= K) return fail("bevgen_labels: major index out of range");
  const int rows = row_end - row_begin;
  if (rows == 0) return 0;
  CK(cudaSetDevice(c->device));
  const size_t dense_n = labels_out ? (size_t)rows * M : 0;
  if (tmp_reserve(c, Carver::pad((size_t)K * 12) + Carver::pad((size_t)M * 12) + Carver::pad((size_t)M * 4) + 2 * Carver::pad((size_t)rows * 8) +
                         Carver::pad(dense_n * 4))) return -1;
  Carver cv{c->tmp};
  float* d_xyz = cv.take<float>((size_t)K * 3); float* d_mpos = cv.take<float>((size_t)M * 3); int32_t* d_mi = cv.take<int32_t>(M);
  int32_t* d_nn = cv.take<int32_t>((size_t)rows * 2); float* d_w = cv.take<float>((size_t)rows * 2);
  float* d_dense = labels_out ? cv.take<float>(dense_n) : nullptr;
  if (labels_out) CK(cudaMemsetAsync(d_dense, 0, dense_n * 4, c->s_comp));
  CK(cudaMemcpyAsync(d_xyz, xyz, (size_t)K * 12, cudaMemcpyHostToDevice, c->s_comp));
  CK(cudaMemcpyAsync(d_mi, major_idx, (size_t)M * 4, cudaMemcpyHostToDevice, c->s_comp));
  k_gather_mpos<<<(M + 127) / 128, 128, 0, c->s_comp>>>(M, d_xyz, d_mi, d_mpos);
  k_labels<<<(rows + 127) / 128, 128, 0, c->s_comp>>>(K, d_xyz, M, d_mi, d_mpos, row_begin, row_end, d_nn, d_w, d_dense);
  CK(cudaGetLastError()); c->launches += 2;
  CK(cudaStreamSynchronize(c->s_comp));
  if (labels_out) CK(cudaMemcpy(labels_out, d_dense, (size_t)rows * M * 4, cudaMemcpyDeviceToHost));
  if (nn_idx) CK(cudaMemcpy(nn_idx, d_nn, (size_t)rows * 8, cudaMemcpyDeviceToHost));
  if (nn_w) CK(cudaMemcpy(nn_w, d_w, (size_t)rows * 8, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- cloud_manip ------------------------------------------------------------------------------------------------
// Kernel + merge on the compute stream.  x .. tz, g_in, g_out are device pointers (g_in / g_out: the final 201 x 201 grids,
// either may be NULL); `rep` = scratch for 2 * n_rep replicated grids.
static int manip_replicas(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(MANIP_MAX_REP, n / 65536)); }
static int launch_cloud_manip(bevgen_ctx* c, int64_t n, const Xform& xf, const float* x, const float* y, const float* z, float* tx, float* ty,
                              float* tz, float* g_in, float* g_out, int* rep, int n_rep) {
  const size_t cells = (size_t)MGRID * MGRID;
  int* rep_in = g_in ? rep : nullptr;
  int* rep_out = g_out ? rep + (size_t)n_rep * cells : nullptr;
  CK(cudaMemsetAsync(rep, 0, 2 * (size_t)n_rep * cells * sizeof(int), c->s_comp));
  if (n > 0) {
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    k_cloud_manip<<<blocks, 256, 0, c->s_comp>>>(n, xf, x, y, z, tx, ty, tz, rep_in, rep_out, n_rep);
    c->launches++;
  }
  k_manip_merge<<<dim3((unsigned)((cells + 255) / 256), 2), 256, 0, c->s_comp>>>(rep_in, rep_out, n_rep, g_in, g_out);
  CK(cudaGetLastError()); c->launches++;
  return 0;
}

extern "C" int bevgen_cloud_manip(bevgen_ctx* c, int64_t n, const float* rt, const float* x, const float* y, const float* z,
                                  float* tx, float* ty, float* tz, float* bev_in, float* bev_out) {
  if (!c || !rt) return fail("bevgen_cloud_manip: null argument");
  if (n < 0) return fail("bevgen_cloud_manip: n < 0");
  if (n > 0 && (!x || !y || !z)) return fail("bevgen_cloud_manip: null point array");   // an empty cloud may come with NULL arrays
  CK(cudaSetDevice(c->device));
  const size_t np = (size_t)std::max<int64_t>(n, 1);
  const int n_rep = manip_replicas(n);
  if (tmp_reserve(c, 6 * Carver::pad(np * 4) + 2 * Carver::pad(MGRID * MGRID * 4) + Carver::pad(2 * (size_t)n_rep * MGRID * MGRID * 4))) return -1;
  Carver cv{c->tmp};
  float* d[6]; float* g[2];
  for (int i = 0; i < 6; i++) d[i] = cv.take<float>(np);
  for (int i = 0; i < 2; i++) g[i] = cv.take<float>(MGRID * MGRID);
  int* rep = cv.take<int>(2 * (size_t)n_rep * MGRID * MGRID);
  Xform xf; memcpy(xf.m, rt, sizeof xf.m); xf.on = 1;
  if (n > 0) {
    CK(cudaMemcpyAsync(d[0], x, (size_t)n * 4, cudaMemcpyHostToDevice, c->s_comp));
    CK(cudaMemcpyAsync(d[1], y, (size_t)n * 4, cudaMemcpyHostToDevice, c->s_comp));
    CK(cudaMemcpyAsync(d[2], z, (size_t)n * 4, cudaMemcpyHostToDevice, c->s_comp));
  }
  if (launch_cloud_manip(c, n, xf, d[0], d[1], d[2], d[3], d[4], d[5], bev_in ? g[0] : nullptr, bev_out ? g[1] : nullptr, rep, n_rep)) return -1;
  CK(cudaStreamSynchronize(c->s_comp));
  // every output is optional on its own
  if (tx && n > 0) CK(cudaMemcpy(tx, d[3], (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (ty && n > 0) CK(cudaMemcpy(ty, d[4], (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (tz && n > 0) CK(cudaMemcpy(tz, d[5], (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (bev_in) CK(cudaMemcpy(bev_in, g[0], MGRID * MGRID * 4, cudaMemcpyDeviceToHost));
  if (bev_out) CK(cudaMemcpy(bev_out, g[1], MGRID * MGRID * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// Device-resident form: every pointer is device memory on the context's device; the grids must hold 201*201 floats and
// are overwritten.  Enqueued on the compute stream (bevgen_sync before reading).
extern "C" int bevgen_cloud_manip_device(bevgen_ctx* c, int64_t n, const float* rt, const float* x, const float* y, const float* z,
                                         float* tx, float* ty, float* tz, float* bev_in, float* bev_out) {
  if (!c || !rt) return fail("bevgen_cloud_manip_device: null argument");
  if (n < 0 || (n > 0 && (!x || !y || !z))) return fail("bevgen_cloud_manip_device: bad point arrays");
  if ((tx || ty || tz) && !(tx && ty && tz)) return fail("bevgen_cloud_manip_device: tx, ty, tz must be given together");
  CK(cudaSetDevice(c->device));
  const int n_rep = manip_replicas(n);
  if (tmp_reserve(c, Carver::pad(2 * (size_t)n_rep * MGRID * MGRID * 4))) return -1;   // grow-only: no allocation after the first call of this size
  Xform xf; memcpy(xf.m, rt, sizeof xf.m); xf.on = 1;
  return launch_cloud_manip(c, n, xf, x, y, z, tx, ty, tz, bev_in, bev_out, reinterpret_cast<int*>(c->tmp), n_rep);
}

// ---- projection step of the keyframe extractors ---------------------------------------------------
extern "C" int bevgen_project(bevgen_ctx* c, int kind, int64_t n, float* x, const float* y, float* z, uint16_t* row, uint16_t* col) {
  if (!c || !x || !y || !row || !col) return fail("bevgen_project: null argument");
  if (kind != BEVGEN_PROJECT_MULRAN_OS1_64 && kind != BEVGEN_PROJECT_OXFORD_HDL_32E && kind != BEVGEN_PROJECT_KITTI_HDL_64E)
    return fail("bevgen_project: unknown kind");
  if (kind == BEVGEN_PROJECT_KITTI_HDL_64E && n > 0x7ffffff0) return fail("bevgen_project: scan too large");
  if (kind == BEVGEN_PROJECT_OXFORD_HDL_32E && !z) return fail("bevgen_project: the Oxford projection needs z");
  if (n < 0) return fail("bevgen_project: n < 0");
  if (n == 0) return 0;
  CK(cudaSetDevice(c->device));
  if (tmp_reserve(c, 3 * Carver::pad((size_t)n * 4) + 2 * Carver::pad((size_t)n * 2) + Carver::pad(((size_t)n / 2 + 2) * 4) +
                         Carver::pad((KITTI_MAX_RINGS + 1) * 4) + 256)) return -1;
  Carver cv{c->tmp};
  float* d[3]; uint16_t* r[2];
  for (int i = 0; i < 3; i++) d[i] = cv.take<float>((size_t)n);
  for (int i = 0; i < 2; i++) r[i] = cv.take<uint16_t>((size_t)n);
  if (kind == BEVGEN_PROJECT_KITTI_HDL_64E) {   // d[2] holds the azimuths instead of z
    int* ev = cv.take<int>((size_t)n / 2 + 2); int* acc = cv.take<int>(KITTI_MAX_RINGS + 1); int* base = cv.take<int>(1);
    CK(cudaMemcpyAsync(d[0], x, (size_t)n * 4, cudaMemcpyHostToDevice, c->s_comp));
    CK(cudaMemcpyAsync(d[1], y, (size_t)n * 4, cudaMemcpyHostToDevice, c->s_comp));
    const unsigned nb = (unsigned)((n + 255) / 256);
    k_kitti_azimuth<<<nb, 256, 0, c->s_comp>>>(n, d[0], d[1], d[2]);
    k_kitti_rings<<<1, 1024, 0, c->s_comp>>>((int)n, d[2], ev, acc, base);
    k_kitti_assign<<<nb, 256, 0, c->s_comp>>>((int)n, d[2], acc, base, r[0], r[1]);
    CK(cudaGetLastError()); c->launches += 3;
    CK(cudaMemcpyAsync(row, r[0], (size_t)n * 2, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaMemcpyAsync(col, r[1], (size_t)n * 2, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaStreamSynchronize(c->s_comp));
    return 0;
  }
  CK(cudaMemcpyAsync(d[0], x, (size_t)n * 4, cudaMemcpyHostToDevice, c->s_comp));
  CK(cudaMemcpyAsync(d[1], y, (size_t)n * 4, cudaMemcpyHostToDevice, c->s_comp));
  if (z) CK(cudaMemcpyAsync(d[2], z, (size_t)n * 4, cudaMemcpyHostToDevice, c->s_comp));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (kind == BEVGEN_PROJECT_MULRAN_OS1_64) k_project<PROJECT_MULRAN><<<blocks, 256, 0, c->s_comp>>>(n, d[0], d[1], d[2], r[0], r[1]);
  else k_project<PROJECT_OXFORD><<<blocks, 256, 0, c->s_comp>>>(n, d[0], d[1], d[2], r[0], r[1]);
  CK(cudaGetLastError()); c->launches++;
  if (kind == BEVGEN_PROJECT_OXFORD_HDL_32E) {   // x and z come back negated (OxfordPointCloudSelect.cpp:203-204)
    CK(cudaMemcpyAsync(x, d[0], (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_comp));
    CK(cudaMemcpyAsync(z, d[2], (size_t)n * 4, cudaMemcpyDeviceToHost, c->s_comp));
  }
  CK(cudaMemcpyAsync(row, r[0], (size_t)n * 2, cudaMemcpyDeviceToHost, c->s_comp));
  CK(cudaMemcpyAsync(col, r[1], (size_t)n * 2, cudaMemcpyDeviceToHost, c->s_comp));
  CK(cudaStreamSynchronize(c->s_comp));
  return 0;
}

// ---- extractTopAndFlatten ---------------------------------------------------------------------------
extern "C" int bevgen_top_flatten(bevgen_ctx* c, int64_t n64, const float* x, const float* y, const float* z, const int16_t* label,
                                  float* out_x, float* out_y, uint32_t* out_index, int64_t* n_out) {
  if (!c || !x || !y || !z || !label || !out_x || !out_y || !n_out) return fail("bevgen_top_flatten: null argument");
  if (n64 < 0 || n64 > (1 << 28)) return fail("bevgen_top_flatten: n out of range");
  *n_out = 0;
  if (n64 == 0) return 0;
  const int n = (int)n64;
  CK(cudaSetDevice(c->device));
  const int n_tiles = (n + RS_TILE - 1) / RS_TILE;
  const size_t np = (size_t)n;
  if (tmp_reserve(c, 3 * Carver::pad(np * 4) + Carver::pad(np * 2) + 2 * Carver::pad(np * 8) + 2 * Carver::pad(np * 4) +
                         Carver::pad((size_t)256 * n_tiles * 4) + 3 * Carver::pad(np * 4) + 4 * Carver::pad(TOP_CELLS * 4) + 256)) return -1;
  Carver cv{c->tmp};
  float* dx = cv.take<float>(np); float* dy = cv.take<float>(np); float* dz = cv.take<float>(np); int16_t* dl = cv.take<int16_t>(np);
  uint64_t* k0 = cv.take<uint64_t>(np); uint64_t* k1 = cv.take<uint64_t>(np);
  uint32_t* v0 = cv.take<uint32_t>(np); uint32_t* v1 = cv.take<uint32_t>(np);
  uint32_t* hist = cv.take<uint32_t>((size_t)256 * n_tiles);
  float* ox = cv.take<float>(np); float* oy = cv.take<float>(np); uint32_t* oi = cv.take<uint32_t>(np);
  int* cstart = cv.take<int>(TOP_CELLS); int* cquota = cv.take<int>(TOP_CELLS); int* cout_ = cv.take<int>(TOP_CELLS); int* dn = cv.take<int>(1);
  cudaStream_t st = c->s_comp;
  CK(cudaMemcpyAsync(dx, x, np * 4, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(dy, y, np * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dz, z, np * 4, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(dl, label, np * 2, cudaMemcpyHostToDevice, st));
  const unsigned nb = (unsigned)((n + 255) / 256);
  k_top_keys<<<nb, 256, 0, st>>>(n, dx, dy, dz, dl, k0, v0);
  c->launches++;
  for (int pass = 0; pass < 5; pass++) {       // 32 bits of height, then the 8 bits of the cell: stable, so ties keep input order
    k_rs_hist<<<n_tiles, RS_T, 0, st>>>(n, pass * 8, k0, hist, n_tiles);
    k_rs_scan<<<1, 1024, 0, st>>>(256 * n_tiles, hist);
    k_rs_scatter<<<n_tiles, RS_T, 0, st>>>(n, pass * 8, k0, v0, hist, n_tiles, k1, v1);
    std::swap(k0, k1); std::swap(v0, v1);
    c->launches += 3;
  }
  k_top_cells<<<1, 128, 0, st>>>(n, k0, cstart, cquota, cout_, dn);
  k_top_gather<<<nb, 256, 0, st>>>(n, k0, v0, cstart, cquota, cout_, dx, dy, ox, oy, oi);
  CK(cudaGetLastError()); c->launches += 2;
  int m = 0;
  CK(cudaMemcpyAsync(&m, dn, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (m > 0) {
    CK(cudaMemcpy(out_x, ox, (size_t)m * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(out_y, oy, (size_t)m * 4, cudaMemcpyDeviceToHost));
    if (out_index) CK(cudaMemcpy(out_index, oi, (size_t)m * 4, cudaMemcpyDeviceToHost));
  }
  *n_out = m;
  return 0;
}

// ---- introspection ----------------------------------------------------------------------------------------------
extern "C" int bevgen_set_profiling(bevgen_ctx* c, int on) {
  if (!c) return fail("null ctx");
  c->prof = on != 0;
  for (auto& m : c->stage_ms) m = 0; for (auto& l : c->stage_launches) l = 0;
  return 0;
}
extern "C" int bevgen_stage_ms(bevgen_ctx* c, float* ms, int64_t* launches) {
  if (!c) return fail("null ctx");
  for (int i = 0; i < BEVGEN_N_STAGES; i++) { if (ms) ms[i] = (float)c->stage_ms[i]; if (launches) launches[i] = c->stage_launches[i]; }
  return 0;
}
extern "C" int64_t bevgen_kernel_launches(bevgen_ctx* c) { return c ? c->launches : 0; }
extern "C" void* bevgen_compute_stream(bevgen_ctx* c) { return c ? (void*)c->s_comp : nullptr; }

extern "C" int bevgen_debug_atan2f(bevgen_ctx* c, int64_t n, const float* y, const float* x, float* out) {
  if (!c || !y || !x || !out) return fail("bevgen_debug_atan2f: null argument");
  if (n <= 0) return 0;
  CK(cudaSetDevice(c->device));
  if (tmp_reserve(c, 3 * Carver::pad((size_t)n * 4))) return -1;
  Carver cv{c->tmp};
  float* dy = cv.take<float>((size_t)n); float* dx = cv.take<float>((size_t)n); float* dout = cv.take<float>((size_t)n);
  CK(cudaMemcpyAsync(dy, y, n * 4, cudaMemcpyHostToDevice, c->s_comp)); CK(cudaMemcpyAsync(dx, x, n * 4, cudaMemcpyHostToDevice, c->s_comp));
  k_debug_atan2f<<<(unsigned)((n + 255) / 256), 256, 0, c->s_comp>>>(n, dy, dx, dout);
  CK(cudaGetLastError()); c->launches++;
  CK(cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, c->s_comp));
  CK(cudaStreamSynchronize(c->s_comp));
  return 0;
}

// ---- libm overload set of the ground criterion + diagnostics ------------------------------------------------------
extern "C" int bevgen_set_libm(bevgen_ctx* c, int use_double) {
  if (!c) return fail("bevgen_set_libm: null ctx");
  c->sp.libm_double = use_double ? 1 : 0;
  return 0;
}
extern "C" int bevgen_set_diag(bevgen_ctx* c, int enabled) {
  if (!c) return fail("bevgen_set_diag: null ctx");
  CK(cudaSetDevice(c->device));
  if (bevgen_sync(c)) return -1;
  CK(cudaMemset(c->diag_d, 0, 4 * sizeof(unsigned long long)));
  c->sp.diag = enabled ? c->diag_d : nullptr;
  return 0;
}
extern "C" int bevgen_get_diag(bevgen_ctx* c, uint64_t* out) {
  if (!c || !out) return fail("bevgen_get_diag: null argument");
  CK(cudaSetDevice(c->device));
  if (bevgen_sync(c)) return -1;
  for (auto& s : c->slots) if (s.st) CK(cudaStreamSynchronize(s.st));
  unsigned long long h[4];
  CK(cudaMemcpy(h, c->diag_d, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; i++) out[i] = h[i];
  return 0;
}

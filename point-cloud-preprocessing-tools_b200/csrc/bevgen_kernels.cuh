// bevgen_kernels.cuh — sm_100a kernels of the batch_multi_bev_gen hot path (hand-written; no libraries).
//
// Reference being replaced: soytony/Point-Cloud-Preprocessing-Tools @ d94040e, BatchMultiBevGen.cpp.
// All float arithmetic that feeds a rounding/threshold decision is written with explicit round-to-nearest
// intrinsics (and the TU is compiled --fmad=false) because the reference is x86-64 SSE2 code without FMA
// (CMakeLists.txt:10): every float op there is an individually rounded IEEE op.
//
// HBM layout per frame f of a batch (S = N_SCAN * Horizon_SCAN slots):
//   winner_bits   1 bit per INPUT point: the point is the last writer of its slot (see bevgen.h)        — output
//   rec   [f][S]  f32x4 ordered cloud: x, y, z, w = {label:16 | I==-1:1 | owned:1} — scratch, written once
//   gsum  [f][G+1][ceil(H/32)] uint4  per (band row, 32-column group): participating lanes, sector-change lanes, first | last
//                 sector, ground_mat == 1 bits of the group (k_ground_mark)                             — scratch
//   gmask [f][ceil(S/32)] u32  the same ground bits in slot order, bit (s & 31) of word (s >> 5) (k_seg_build)  — scratch
//   gz    [f][S]  f32   z of the ground slots (0 elsewhere)                  — scratch
//   cnt   [f][3750] u32 ground slots per sector (zero heights: k_ground_mark, the others: k_seg_build);  avg [f][3750] f32 sector mean heights — scratch
//   label [f][S] i16, single [f][224*224] u8, multi [f][24][224*224] u8      — outputs
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bevgen {

constexpr int GRID = 224;
constexpr int CELLS = GRID * GRID;        // 50176
constexpr int CELL_WORDS = CELLS / 4;     // 12544 u32 words of 4 cell-bytes
constexpr int LAYERS = 24;
constexpr int SECT_R = 75, SECT_C = 50, NSECT = SECT_R * SECT_C;
constexpr unsigned NO_KEY = 0xFFFFu;
constexpr unsigned W_NEG1 = 1u << 16;     // rec.w flag: intensity == -1
constexpr unsigned W_OWNED = 1u << 17;    // rec.w flag: slot written by an input point

struct SensorDev {
  int N, H, G, S;
  int band_row0;        // N - G - 1: first row that can carry ground_mat == 1
  float height_res;
  float inv_height_res; // exact when height_res is a power of two (all three sensors)
  int hr_pow2;
  // ground criterion constants (computed on the host at context creation, see bevgen_capi.cu)
  float t_star;         // largest float t with (float)((double)t*180.0/M_PI) <= 10.0f   (BatchMultiBevGen.cpp:173,179)
  float q_lo2, q_hi2;   // (tan(t_star)*(1 -/+ 1e-5))^2: outside this band of dz^2/(dx^2+dy^2) no atan2f is needed
  int libm_double;      // 1: decide borderline pairs with the C double atan2 / sqrt (the other overload set)
  unsigned long long* diag;   // NULL, or device counters: [0] pairs inside the guard band, [1] pairs where the float- and
                              // the double-libm decision differ
};

struct Xform { float m[12]; int on; };

// ------------------------------------------------------------------------------------------------------------
// glibc 2.39 float atan2f / atanf (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c — the fdlibm algorithm),
// restated with individually rounded float ops.  tests/test_gpu_parity.py::test_atan2f_bit_exact checks it
// against the host libm bit for bit; oracle/ probes showed 0 mismatches in 4e8 samples for the C mirror.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float atanf_glibc(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  int hx = __float_as_int(x);
  int ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return __fadd_rn(x, x);
    float r = __fadd_rn(atanhi[3], atanlo[3]);
    return hx > 0 ? r : -r;
  }
  if (ix < 0x3ee00000) {   // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x)); }
      else                 { id = 1; x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f)); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x))); }
      else                 { id = 3; x = __fdiv_rn(-1.0f, x); }
    }
  }
  float z = __fmul_rn(x, x);
  float w = __fmul_rn(z, z);
  float s1 = __fmul_rn(z, __fadd_rn(aT0, __fmul_rn(w, __fadd_rn(aT2, __fmul_rn(w, __fadd_rn(aT4, __fmul_rn(w,
             __fadd_rn(aT6, __fmul_rn(w, __fadd_rn(aT8, __fmul_rn(w, aT10)))))))))));
  float s2 = __fmul_rn(w, __fadd_rn(aT1, __fmul_rn(w, __fadd_rn(aT3, __fmul_rn(w, __fadd_rn(aT5, __fmul_rn(w,
             __fadd_rn(aT7, __fmul_rn(w, aT9)))))))));
  float s = __fadd_rn(s1, s2);
  if (id < 0) return __fsub_rn(x, __fmul_rn(x, s));
  float hi = id == 0 ? atanhi[0] : id == 1 ? atanhi[1] : id == 2 ? atanhi[2] : atanhi[3];
  float lo = id == 0 ? atanlo[0] : id == 1 ? atanlo[1] : id == 2 ? atanlo[2] : atanlo[3];
  z = __fsub_rn(hi, __fsub_rn(__fsub_rn(__fmul_rn(x, s), lo), x));
  return hx < 0 ? -z : z;
}

__device__ __forceinline__ float atan2f_glibc(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  int hx = __float_as_int(x), hy = __float_as_int(y);
  int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return __fadd_rn(x, y);
  if (hx == 0x3f800000) return atanf_glibc(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    if (m < 2) return y;
    return m == 2 ? __fadd_rn(pi, tiny) : __fsub_rn(-pi, tiny);
  }
  if (ix == 0) return hy < 0 ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return __fadd_rn(pi_o_4, tiny);
        case 1: return __fsub_rn(-pi_o_4, tiny);
        case 2: return __fadd_rn(__fmul_rn(3.0f, pi_o_4), tiny);
        default: return __fsub_rn(__fmul_rn(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return __fadd_rn(pi, tiny);
        default: return __fsub_rn(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return hy < 0 ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
  int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = __fadd_rn(pi_o_2, __fmul_rn(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_glibc(fabsf(__fdiv_rn(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return __fsub_rn(pi, __fsub_rn(z, pi_lo));
    default: return __fsub_rn(__fsub_rn(z, pi_lo), pi);
  }
}

__global__ void k_debug_atan2f(int64_t n, const float* __restrict__ y, const float* __restrict__ x, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = atan2f_glibc(y[i], x[i]);
}

// x86 cvttsd2si: what the reference binary executes for (int)double — INT_MIN for NaN / out of range.
__device__ __forceinline__ int cvtt_x86(double v) {
  return (v > -2147483649.0 && v < 2147483648.0) ? __double2int_rz(v) : INT32_MIN;
}

// getBelongingGrid, BatchMultiBevGen.h:73-99 -> sector id row*50+col, in float-only arithmetic:
//  * normalized = float(double(p) + 75.0) equals the single-rounded float sum fl(p + 75.0f): the double sum is exact
//    unless |p| < 2^-23, and then both roundings give exactly 75.0f (p is far below half a float ulp of 75).
//  * floor(double(n) / 2.0) == floorf(n * 0.5f) (scaling by 1/2 is exact; a denormal that rounds to -0 clamps to 0 either way).
//  * static_cast<int> is cvttsd2si: NaN and |v| >= 2^31 give INT_MIN, which the `< 0` clamp turns into 0.
__device__ __forceinline__ int sector_axis(float p, float off, int n) {
  const float t = __fmul_rn(__fadd_rn(p, off), 0.5f);
  int i = __float2int_rd(t);                 // floor; saturates, NaN -> 0
  if (t >= 2147483648.0f) i = 0;             // cvttsd2si gives INT_MIN there, which the reference clamps to 0
  return min(max(i, 0), n - 1);
}
__device__ __forceinline__ unsigned sector_of(float px, float py) {
  return (unsigned)(sector_axis(px, 75.0f, SECT_R) * SECT_C + sector_axis(py, 50.0f, SECT_C));
}

// ------------------------------------------------------------------------------------------------------------
// Kr project — SURVEY 8(f)-2: the keyframe extractors' per-point projection (row / col of the range image), the step
// that produces the `row` / `col` fields batch_multi_bev_gen consumes.
//   MULRAN (MulranPointCloudSelect.cpp:112-126): row = k % 64; az = float(double(atan2(y, x)) / M_PI * 180.0f), wrapped
//     into [0, 360]; col = uint16(round(az / 360.0f * 1024))  (col may equal 1024).
//   OXFORD (OxfordPointCloudSelect.cpp:201-219): x, z negated (sensor mounted upside-down); elevation =
//     float(double(atan2(z, sqrt(x*x + y*y))) / M_PI * 180.0f); row = clamp(int(round((-elevation + 10.67) / 1.3335)), 0, 31);
//     col as above with 1056 columns, then `if (col >= 1056) col -= 1056`.
// atan2 / sqrt are the float overloads (normative choice of SURVEY 8a.1-G3): atan2f_glibc is bit-exact to glibc.
// One thread per point.  grid ceil(n/256), block 256.
// ------------------------------------------------------------------------------------------------------------
constexpr int PROJECT_MULRAN = 0, PROJECT_OXFORD = 1;

__device__ __forceinline__ float rad2deg_ref(float t) {              // (float)((double)t / M_PI * 180.0f)
  return __double2float_rn(__dmul_rn(__ddiv_rn((double)t, 3.14159265358979323846), 180.0));
}
__device__ __forceinline__ uint16_t u16_cast_x86(float v) {          // static_cast<uint16_t>(float): cvttss2si, low 16 bits
  const int i = (v > -2147483904.0f && v < 2147483648.0f) ? __float2int_rz(v) : INT32_MIN;   // NaN / out of range -> INT_MIN
  return (uint16_t)(i & 0xFFFF);
}

template <int KIND>
__global__ void __launch_bounds__(256) k_project(int64_t n, float* __restrict__ x, const float* __restrict__ y, float* __restrict__ z,
                                                  uint16_t* __restrict__ row, uint16_t* __restrict__ col) {
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= n) return;
  float px = x[k], py = y[k];
  constexpr float COLS = KIND == PROJECT_MULRAN ? 1024.0f : 1056.0f;
  if (KIND == PROJECT_MULRAN) {
    row[k] = (uint16_t)(k % 64);                                     // :120
  } else {
    float pz = z[k];
    px = -px; pz = -pz;                                              // :203-204
    x[k] = px; z[k] = pz;
    const float hyp = __fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)));
    const float elev = rad2deg_ref(atan2f_glibc(pz, hyp));           // :208
    const double rr = round(__ddiv_rn(__dadd_rn((double)(-elev), 10.67), 1.3335));   // :209
    int ri = cvtt_x86(rr);
    ri = min(31, max(0, ri));                                        // :210
    row[k] = (uint16_t)ri;
  }
  float az = rad2deg_ref(atan2f_glibc(py, px));                      // :121 / :213
  if (az > 360.0f) az = __fsub_rn(az, 360.0f);
  else if (az < 0.0f) az = __fadd_rn(az, 360.0f);
  uint16_t c = u16_cast_x86(roundf(__fmul_rn(__fdiv_rn(az, 360.0f), COLS)));   // :125 / :216
  if (KIND == PROJECT_OXFORD && c >= 1056) c -= 1056;                // :217
  col[k] = c;
}

// ------------------------------------------------------------------------------------------------------------
// Kr' KITTI ring detection — KittiPointCloudSelect.cpp:188-243.  The scan file has no ring index; the extractor walks
// the points in file order and starts a new ring when the azimuth crosses from <= 0 to > 0, but only if the current
// ring already holds more than Horizon_SCAN * 0.60f points (:213-221).  That acceptance rule is a serial chain over
// the CROSSINGS only (a few hundred per scan), so:
//   k_kitti_azimuth : azimuth per point (parallel)
//   k_kitti_rings   : one CTA compacts the crossing indices in file order, thread 0 runs the greedy acceptance chain
//   k_kitti_assign  : ring(i) = base + #accepted crossings <= i (binary search), col = round(az' / (360.0 / 2083))
// Point 0 is never placed (the loop starts at 1, :211); points of rings outside 0..63 are dropped (:228): both come
// back as row = col = 0xFFFF, which getOrderedCloud's bounds test discards.
// ------------------------------------------------------------------------------------------------------------
constexpr int KITTI_N = 64, KITTI_H = 2083;
constexpr int KITTI_MAX_RINGS = 4096;        // accepted crossings kept (they are >= 1250 points apart)

__global__ void __launch_bounds__(256) k_kitti_azimuth(int64_t n, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ az) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) az[i] = rad2deg_ref(atan2f_glibc(y[i], x[i]));          // :191-194
}

// grid 1, block 1024.  ev: scratch for the crossing indices (capacity n/2 + 1); acc[0] = number of accepted crossings,
// acc[1..] their indices; acc_base = ring index before the first accepted crossing (0 or -1, :199-204).
__global__ void __launch_bounds__(1024) k_kitti_rings(int n, const float* __restrict__ az, int* __restrict__ ev, int* __restrict__ acc,
                                                       int* __restrict__ acc_base) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int n_ev = 0;
  for (int b = 1; b < n; b += 1024) {                                 // crossings in file order (stable compaction)
    const int i = b + tid;
    const bool e = i < n && az[i - 1] <= 0.0f && az[i] > 0.0f;        // :213
    const unsigned m = __ballot_sync(0xffffffffu, e);
    if (lane == 0) s_warp[wid] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; w++) { const int c = s_warp[w]; if (w < wid) before += c; total += c; }
    if (e) ev[n_ev + before + __popc(m & ((1u << lane) - 1u))] = i;
    n_ev += total;
    __syncthreads();
  }
  if (tid == 0) s_total = n_ev;
  __threadfence_block();
  __syncthreads();
  if (tid == 0) {                                                     // the serial acceptance chain (:214-220)
    int ring = az[0] > 0.0f ? 0 : -1;                                 // :199-204
    *acc_base = ring;
    int last = 1;                                                     // num_points_on_this_ring == i - last at index i
    int na = 0;
    const float need = __fmul_rn((float)KITTI_H, 0.60f);
    for (int k = 0; k < s_total; k++) {
      const int i = ev[k];
      bool take;
      if (ring == -1) take = true;                                    // :214-216
      else take = (float)(i - last) > need;                           // :217
      if (take) { ring++; last = i; if (na < KITTI_MAX_RINGS) acc[1 + na] = i; na++; }
    }
    acc[0] = min(na, KITTI_MAX_RINGS);
  }
}

__global__ void __launch_bounds__(256) k_kitti_assign(int n, const float* __restrict__ az, const int* __restrict__ acc,
                                                       const int* __restrict__ acc_base, uint16_t* __restrict__ row, uint16_t* __restrict__ col) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int na = acc[0];
  int lo = 0, hi = na;                                               // number of accepted crossings with index <= i
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (acc[1 + mid] <= i) lo = mid + 1; else hi = mid; }
  const int ring = *acc_base + lo;
  float a = az[i];
  if (a >= 360.0f) a = __fsub_rn(a, 360.0f); else if (a < 0.0f) a = __fadd_rn(a, 360.0f);            // makeAngleSemiPositive :137-146
  int c = cvtt_x86(round(__ddiv_rn((double)a, __ddiv_rn(360.0, (double)KITTI_H))));                  // :226
  uint16_t r16 = 0xFFFFu, c16 = 0xFFFFu;
  if (i >= 1 && ring >= 0 && ring < KITTI_N) {                        // :211, :228
    if (c >= KITTI_H) c -= KITTI_H; else if (c < 0) c += KITTI_H;    // :229-233
    if (c >= 0 && c < KITTI_H) { r16 = (uint16_t)ring; c16 = (uint16_t)c; }   // (outside: the reference indexes out of bounds)
  }
  row[i] = r16; col[i] = c16;
}

// ------------------------------------------------------------------------------------------------------------
// Kp unpack_records — SURVEY §8(f)-1: the de-interleave of pcl::io::loadPCDFile (BatchMultiBevGen.cpp:730) moved to the
// GPU.  A binary PCD payload is an array of interleaved records (26 packed bytes for PointXYZIRCT as written by
// savePCDFileBinary: x y z intensity f32 | row col u16 | t u32 | label i16, BatchMultiBevGen.h:56-66); the host only
// copies the payload into pinned memory, this kernel turns it into the SoA arrays the ordering kernels read.
// Fields are addressed by byte offset inside a record (any order / padding; -1 = field absent => 0, like
// pcl::fromPCLPointCloud2 leaves a missing field value-initialised).  A CTA stages its 256 records through shared
// memory with 16-byte coalesced loads (records are only 2-byte aligned), then every thread assembles its record.
// grid (ceil(max_n/256), F), block 256, dynamic smem 256*stride + 32.
// ------------------------------------------------------------------------------------------------------------
struct RecLayout { int stride; int off[7]; };   // x, y, z, intensity (f32), row, col (u16), label (i16)

template <bool EVEN>   // EVEN: stride and every offset are even => 16-bit shared-memory reads instead of bytes
__global__ void __launch_bounds__(256) k_unpack_records(RecLayout L, const int64_t* __restrict__ offs, int64_t base,
                                                         const uint8_t* __restrict__ raw, float* __restrict__ x,
                                                         float* __restrict__ y, float* __restrict__ z, float* __restrict__ inten,
                                                         uint16_t* __restrict__ row, uint16_t* __restrict__ col,
                                                         int16_t* __restrict__ label) {
  extern __shared__ __align__(16) unsigned char urec[];
  const int f = blockIdx.y;
  const int64_t o = offs[f] - base;                    // first point of the frame inside the staged chunk
  const int n = (int)(offs[f + 1] - offs[f]);
  const int i0 = blockIdx.x * 256;
  if (i0 >= n) return;
  const int cnt = min(256, n - i0);
  const int64_t b0 = (o + i0) * (int64_t)L.stride;     // first byte of the tile; `raw` is 16-byte aligned and padded
  const int64_t a0 = b0 & ~(int64_t)15;
  const int head = (int)(b0 - a0);
  const int n16 = (head + cnt * L.stride + 15) >> 4;
  const uint4* src = reinterpret_cast<const uint4*>(raw + a0);
  for (int v = threadIdx.x; v < n16; v += 256) reinterpret_cast<uint4*>(urec)[v] = __ldcs(src + v);
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= cnt) return;
  const unsigned char* r = urec + head + t * L.stride;
  auto u16 = [&](int off) -> unsigned {
    if (EVEN) return *reinterpret_cast<const uint16_t*>(r + off);
    return (unsigned)r[off] | ((unsigned)r[off + 1] << 8);
  };
  auto u32 = [&](int off) -> unsigned { return u16(off) | (u16(off + 2) << 16); };
  const int64_t q = o + i0 + t;
  x[q] = L.off[0] >= 0 ? __uint_as_float(u32(L.off[0])) : 0.0f;
  y[q] = L.off[1] >= 0 ? __uint_as_float(u32(L.off[1])) : 0.0f;
  z[q] = L.off[2] >= 0 ? __uint_as_float(u32(L.off[2])) : 0.0f;
  inten[q] = L.off[3] >= 0 ? __uint_as_float(u32(L.off[3])) : 0.0f;
  row[q] = L.off[4] >= 0 ? (uint16_t)u16(L.off[4]) : (uint16_t)0;
  col[q] = L.off[5] >= 0 ? (uint16_t)u16(L.off[5]) : (uint16_t)0;
  label[q] = L.off[6] >= 0 ? (int16_t)u16(L.off[6]) : (int16_t)0;
}

// ------------------------------------------------------------------------------------------------------------
// K0a order_claim — getOrderedCloud (BatchMultiBevGen.cpp:94-117), pass 1: the serial loop makes the LAST input
// point of a slot win; atomicMax over (input index + 1) reproduces that deterministically.
// grid (ceil(max_n/256), F), block 256.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_order_claim(SensorDev sp, const int64_t* __restrict__ offs,
                                                      const uint16_t* __restrict__ row, const uint16_t* __restrict__ col,
                                                      uint32_t* __restrict__ owner) {
  const int f = blockIdx.y;
  const int64_t o = offs[f];
  const int n = (int)(offs[f + 1] - o);
  const int i0 = blockIdx.x * 512 + threadIdx.x, i1 = i0 + 256;   // two points per thread: four loads in flight
  unsigned r0 = 0xFFFF, c0 = 0xFFFF, r1 = 0xFFFF, c1 = 0xFFFF;
  if (i0 < n) { r0 = row[o + i0]; c0 = col[o + i0]; }
  if (i1 < n) { r1 = row[o + i1]; c1 = col[o + i1]; }
  uint32_t* own = owner + (size_t)f * sp.S;
  if (i0 < n && r0 < (unsigned)sp.N && c0 < (unsigned)sp.H) atomicMax(&own[r0 * sp.H + c0], (uint32_t)(i0 + 1));   // :106-113
  if (i1 < n && r1 < (unsigned)sp.N && c1 < (unsigned)sp.H) atomicMax(&own[r1 * sp.H + c1], (uint32_t)(i1 + 1));
}

// K0b order_fill — pass 2: winners write their record, unowned slots get the value-initialised record (:98).
// Optional rigid transform (pcl::transformPointCloud, PCL>=1.9 SSE order p0 + (p1 + (p2 + t)), CloudManip.cpp:128).
// grid (ceil(max(max_n,S)/256), F), block 256.
__global__ void __launch_bounds__(256) k_order_fill(SensorDev sp, Xform xf, const int64_t* __restrict__ offs,
                                                     const float* __restrict__ x, const float* __restrict__ y,
                                                     const float* __restrict__ z, const float* __restrict__ inten,
                                                     const uint16_t* __restrict__ row, const uint16_t* __restrict__ col,
                                                     const int16_t* __restrict__ label, const uint32_t* __restrict__ owner,
                                                     float4* __restrict__ rec) {
  const int f = blockIdx.y;
  const int64_t o = offs[f];
  const int n = (int)(offs[f + 1] - o);
  const int i = blockIdx.x * 256 + threadIdx.x;
  const size_t fb = (size_t)f * sp.S;
  // All loads that do not depend on the owner probe are issued up front (the kernel is bound by the latency of its
  // dependent chain row/col -> owner[slot] -> record store, not by bandwidth); 99.5 % of the points are winners.
  const bool pv = i < n;
  unsigned r = 0xFFFF, c = 0xFFFF;
  float px = 0.f, py = 0.f, pz = 0.f, pi = 0.f; int16_t lb = 0;
  if (pv) { r = row[o + i]; c = col[o + i]; px = x[o + i]; py = y[o + i]; pz = z[o + i]; pi = inten[o + i]; lb = label[o + i]; }
  const uint32_t own_s = i < sp.S ? owner[fb + i] : 1u;
  const bool valid = pv && r < (unsigned)sp.N && c < (unsigned)sp.H;
  const size_t slot = fb + (valid ? r * sp.H + c : 0u);
  const uint32_t own_p = valid ? owner[slot] : 0u;
  if (i < sp.S && own_s == 0) rec[fb + i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid && own_p == (uint32_t)(i + 1)) {
    if (xf.on) {
      float ox = __fadd_rn(__fmul_rn(px, xf.m[0]), __fadd_rn(__fmul_rn(py, xf.m[1]), __fadd_rn(__fmul_rn(pz, xf.m[2]), xf.m[3])));
      float oy = __fadd_rn(__fmul_rn(px, xf.m[4]), __fadd_rn(__fmul_rn(py, xf.m[5]), __fadd_rn(__fmul_rn(pz, xf.m[6]), xf.m[7])));
      float oz = __fadd_rn(__fmul_rn(px, xf.m[8]), __fadd_rn(__fmul_rn(py, xf.m[9]), __fadd_rn(__fmul_rn(pz, xf.m[10]), xf.m[11])));
      px = ox; py = oy; pz = oz;
    }
    const unsigned w = (unsigned)(uint16_t)lb | W_OWNED | (pi == -1.0f ? W_NEG1 : 0u);
    rec[slot] = make_float4(px, py, pz, __uint_as_float(w));
  }
}

// Winner bits for the global-memory claim path: point i survives iff owner[slot(i)] == i + 1.
// grid (ceil(max_n/256), F), block 256.
__global__ void __launch_bounds__(256) k_winner_bits(SensorDev sp, const int64_t* __restrict__ offs, int frame0,
                                                      const uint16_t* __restrict__ row, const uint16_t* __restrict__ col,
                                                      const uint32_t* __restrict__ owner, uint32_t* __restrict__ winner_bits) {
  const int f = blockIdx.y;
  const int64_t o = offs[f];
  const int n = (int)(offs[f + 1] - o);
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i == 0)
    for (int64_t w = (n + 31) >> 5; w < ((o + n) >> 5) + 1 - (o >> 5); w++) winner_bits[(o >> 5) + (frame0 + f) + w] = 0u;
  if (i - (int)(threadIdx.x & 31) >= n) return;
  bool win = false;
  if (i < n) {
    const unsigned r = row[o + i], c = col[o + i];
    if (r < (unsigned)sp.N && c < (unsigned)sp.H) win = owner[(size_t)f * sp.S + r * sp.H + c] == (uint32_t)(i + 1);
  }
  const unsigned m = __ballot_sync(0xffffffffu, win);
  if ((threadIdx.x & 31) == 0) winner_bits[(o >> 5) + (frame0 + f) + (i >> 5)] = m;
}

// ------------------------------------------------------------------------------------------------------------
// K0 order, shared-memory form — getOrderedCloud (BatchMultiBevGen.cpp:94-117) as two kernels:
//
// k_order_winners (one CTA per frame): which input point survives in each slot, decided entirely in shared memory.
//   occ  : 1 bit per slot, set by every valid point (atomicOr); a point that finds its bit already set marks the
//   cont : "contended" bit of the slot (two or more input points map to it; 0.5 % of the slots on sensor data).
//   A point whose slot is not contended is the slot's only writer and wins without further communication.  Contended
//   slots get a dense id (prefix popcount over `cont`); the serial loop's last-writer-wins is max(input index) per id
//   (one global atomicMax per contended point into cwin[f][id]).  Reads only row/col (4 B per point, 128-bit loads);
//   writes the occupancy / contention bits and the id prefix of the frame (3 x S/8 bytes).
// k_order_scatter (grid over points x frames): a point wins iff its slot is uncontended or cwin says so; it emits
//   the WINNER bit per input point, winners write their record into the ordered cloud `rec` (one scattered 16-byte
//   store per point; all CTAs of a frame run together, so the two halves of a 32-byte sector meet in L2), unowned
//   slots get the value-initialised record (:98).
//
// The previous form (k_order_claim / k_order_fill, kept for range images too large for shared memory) paid an L2
// atomic and an L2 probe per point on top of the store.  A single fused CTA-per-frame kernel was measured and
// rejected: with ~300 frames in flight the half-written sectors of `rec` are evicted before their second half
// arrives (DRAM traffic 13.5 MB per frame).
// ------------------------------------------------------------------------------------------------------------
#ifndef ORD_THREADS
#define ORD_THREADS 512   // measured: 0.275 / 0.251 / 0.261 us per frame for 256 / 512 / 1024 threads
#endif
constexpr int ORD_T = ORD_THREADS;
#ifndef ORD_PREFETCH
#define ORD_PREFETCH 0   // measured (profiles/r2_notes.md): 0.245 us per frame without, 0.252 with (40 registers, 3 CTAs per SM), 0.262 at 4 CTAs (spills)
#endif
#ifdef ORD_MIN_CTAS      // CTAs of 512 threads per SM the register allocation must leave room for (only with ORD_PREFETCH)
#define ORD_BOUNDS __launch_bounds__(ORD_T, ORD_MIN_CTAS)
#else
#define ORD_BOUNDS __launch_bounds__(ORD_T)
#endif
// occupancy + contention words, each padded to a multiple of 32 words (the swizzle below permutes inside 32-word blocks)
__host__ __device__ inline size_t ord_smem_bytes(int S) { return ((((size_t)S + 31) / 32 + 31) & ~(size_t)31) * 4 * 2 + 256; }
// Shared-memory word of slot-word w.  Scans in sensor order (MulRan: row = k % 64, so the 32 lanes of a warp hold 32 rows of ONE
// column) would hit words w = row * (H / 32) + const: with H = 1024 all in one bank, a 32-way conflict on every atomic (OS1_64:
// 0.25 us per frame for 65 k points, as much as HDL_64E's two scans of 118 k).  XOR-ing the bank bits with the next five
// bits spreads such a column over the banks and leaves scattered input as it was.
// Only rows that are a multiple of 4 words long need it (H a multiple of 128; OS1_64: 1024); elsewhere consecutive rows already
// start in different banks and the kernel is instantiated without the two extra instructions per point (ord_needs_swizzle).
template <bool SWZ> __device__ __forceinline__ unsigned ord_swz(unsigned w) { return SWZ ? (w ^ ((w >> 5) & 31u)) : w; }
__host__ __device__ inline bool ord_needs_swizzle(int H) { return (H & 127) == 0; }   // rows of a multiple of 4 words: 8-way conflicts or worse

// dst = *p if i < n (dst keeps its value otherwise): a predicated load instead of a branch around it.  volatile: the loads stay
// in program order, none is dropped; .cs = streaming (read once), .nc = read-only path.
#define BEVGEN_LD_IF(NAME, CTYPE, CONS, PTXLD)                                                                          \
  __device__ __forceinline__ void NAME(CTYPE& dst, const void* p, int i, int n) {                                       \
    asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.s32 q, %2, %3;\n\t@q " PTXLD " %0, [%1];\n\t}" : "+" CONS(dst) : "l"(p), "r"(i), "r"(n)); \
  }
BEVGEN_LD_IF(ld_f32_cs_if, float, "f", "ld.global.cs.f32")
BEVGEN_LD_IF(ld_u32_cs_if, unsigned, "r", "ld.global.cs.u32")
BEVGEN_LD_IF(ld_u32_nc_if, unsigned, "r", "ld.global.nc.u32")
BEVGEN_LD_IF(ld_u16_nc_if, unsigned, "r", "ld.global.nc.u16")
BEVGEN_LD_IF(ld_s16_cs_if, int, "r", "ld.global.cs.s16")
#undef BEVGEN_LD_IF

// 8 consecutive u16 as one 128-bit load: block v8 of the 16-byte aligned pointer.
__device__ __forceinline__ uint4 ld8_u16(const uint16_t* aligned_base, int v8) {
  return __ldg(reinterpret_cast<const uint4*>(aligned_base) + v8);
}

// PACKED (bevgen_process_host_compact): `row` points at the u32 meta array instead (slot | flags, see bevgen.h), `col` is unused.
constexpr unsigned META_SLOT = 0x00FFFFFFu, META_NEG1 = 1u << 24, META_LABELED = 1u << 25;
template <bool PACKED, bool SWZ>
__global__ void ORD_BOUNDS k_order_winners(SensorDev sp, const int64_t* __restrict__ offs, int cw_stride,
                                                          const uint16_t* __restrict__ row, const uint16_t* __restrict__ col,
                                                          uint32_t* __restrict__ occ_bits, uint32_t* __restrict__ cont_bits,
                                                          uint32_t* __restrict__ cont_pre, uint32_t* __restrict__ cwin,
                                                          int64_t qbase, uint32_t* __restrict__ cpt_bits) {
  extern __shared__ __align__(16) unsigned char ord_smem[];
  const int W = (sp.S + 31) >> 5;
  uint32_t* occ = reinterpret_cast<uint32_t*>(ord_smem);           // [W]
  const int WP = (W + 31) & ~31;                                   // padded: ord_swz stays inside a 32-word block
  uint32_t* cont = occ + WP;                                       // [WP]
  uint32_t* misc = cont + WP;                                      // [64] scan carries
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t o = offs[f];
  const int n = (int)(offs[f + 1] - o);
  const unsigned N = (unsigned)sp.N, H = (unsigned)sp.H;
  // row and col start at the same element offset, so they share the misalignment (both arrays are 16-byte aligned)
  const uint16_t* R = row + o; const uint16_t* C = col + o;
  const int mis = (int)((reinterpret_cast<uintptr_t>(R) >> 1) & 7);
  const bool vec = (((reinterpret_cast<uintptr_t>(R) ^ reinterpret_cast<uintptr_t>(C)) & 15) == 0);
  const uint16_t* Ra = R - mis; const uint16_t* Ca = C - mis;
  const int n8 = (n + mis + 7) >> 3;                                // 16-byte blocks that hold the frame's row/col

  for (int i = tid; i < WP; i += ORD_T) { occ[i] = 0u; cont[i] = 0u; }
  __syncthreads();
  // ---- scan 1: occupancy + contention bits ----
  auto visit1 = [&](unsigned slot, bool ok) {
    if (ok) {                                                       // :106-109
      const unsigned bit = 1u << (slot & 31);
      const unsigned w = ord_swz<SWZ>(slot >> 5);
      if (atomicOr(&occ[w], bit) & bit) atomicOr(&cont[w], bit);
    }
  };
  // cpt_bits: one bit per INPUT point (bit index = o + i - qbase, qbase = 32-aligned first point of the wave), set iff the
  // point's slot is contended.  k_order_scatter reads it coalesced, so only the 0.5 % contended points pay the random
  // cont_bits / cont_pre / cwin lookups (ncu: those lookups were 0.76 L2 sector reads per point, a fifth of its L2 traffic).
  const int64_t q0 = o - qbase;
  auto visit2 = [&](unsigned slot, bool ok, int i) {
    if (ok) {
      const unsigned w = ord_swz<SWZ>(slot >> 5);
      const unsigned bit = 1u << (slot & 31), cw = cont[w];
      if (cw & bit) {
        atomicMax(&cwin[(size_t)f * cw_stride + occ[w] + __popc(cw & (bit - 1u))], (uint32_t)i + 1u);
        const int64_t q = q0 + i;
        atomicOr(&cpt_bits[q >> 5], 1u << (q & 31));
      }
    }
  };
  // visit(slot, valid, input index) over the frame's points, 16 bytes per load where the alignment allows
  auto scan = [&](auto visit) {
    if (PACKED) {
      const uint32_t* M = reinterpret_cast<const uint32_t*>(row) + o;
      const int mis4 = (int)((reinterpret_cast<uintptr_t>(M) >> 2) & 3);
      const uint32_t* Ma = M - mis4;
      const int n4 = (n + mis4 + 3) >> 2;
      for (int v = tid; v < n4; v += ORD_T) {
        const int i = v * 4 - mis4;
        if (i < 0 || i + 4 > n) {
          for (int k = max(i, 0); k < min(i + 4, n); k++) { const unsigned sl = M[k] & META_SLOT; visit(sl, sl < (unsigned)sp.S, k); }
          continue;
        }
        const uint4 mm = __ldg(reinterpret_cast<const uint4*>(Ma) + v);
        visit(mm.x & META_SLOT, (mm.x & META_SLOT) < (unsigned)sp.S, i);     visit(mm.y & META_SLOT, (mm.y & META_SLOT) < (unsigned)sp.S, i + 1);
        visit(mm.z & META_SLOT, (mm.z & META_SLOT) < (unsigned)sp.S, i + 2); visit(mm.w & META_SLOT, (mm.w & META_SLOT) < (unsigned)sp.S, i + 3);
      }
      return;
    }
    auto rc = [&](unsigned r, unsigned c, int i) { visit(r * H + c, r < N && c < H, i); };   // callers only pass 0 <= i < n
    if (vec) {
#if ORD_PREFETCH
      // Experiment, off by default: register double buffer - the next block's row / col words are in flight while this block's
      // eight points go through the shared-memory atomics (ncu source view: a third of the kernel's stall samples sit on the
      // first use of these loads).  Measured slower: the eight extra registers cost a CTA per SM (or spill), and the other
      // warps already covered that wait (issue slots 71 % busy).
      auto whole = [&](int v) { const int i = v * 8 - mis; return v < n8 && i >= 0 && i + 8 <= n; };
      uint4 nr = make_uint4(0, 0, 0, 0), nc = nr;
      if (whole(tid)) { nr = ld8_u16(Ra, tid); nc = ld8_u16(Ca, tid); }
#endif
      for (int v = tid; v < n8; v += ORD_T) {
        const int i = v * 8 - mis;
#if ORD_PREFETCH
        const uint4 rr = nr, cc = nc;
        if (whole(v + ORD_T)) { nr = ld8_u16(Ra, v + ORD_T); nc = ld8_u16(Ca, v + ORD_T); }
#endif
        if (i < 0 || i + 8 > n) {                                   // head / tail block: stay inside the frame's elements
          for (int k = max(i, 0); k < min(i + 8, n); k++) rc(R[k], C[k], k);
          continue;
        }
#if !ORD_PREFETCH
        const uint4 rr = ld8_u16(Ra, v), cc = ld8_u16(Ca, v);
#endif
        rc(rr.x & 0xFFFFu, cc.x & 0xFFFFu, i);     rc(rr.x >> 16, cc.x >> 16, i + 1);
        rc(rr.y & 0xFFFFu, cc.y & 0xFFFFu, i + 2); rc(rr.y >> 16, cc.y >> 16, i + 3);
        rc(rr.z & 0xFFFFu, cc.z & 0xFFFFu, i + 4); rc(rr.z >> 16, cc.z >> 16, i + 5);
        rc(rr.w & 0xFFFFu, cc.w & 0xFFFFu, i + 6); rc(rr.w >> 16, cc.w >> 16, i + 7);
      }
    } else {
      for (int i = tid; i < n; i += ORD_T) rc(R[i], C[i], i);
    }
  };
  scan([&](unsigned slot, bool ok, int) { visit1(slot, ok); });
  __syncthreads();
  // ---- occupancy / contention bits out; dense ids of the contended slots (prefix popcount, kept in `occ`) ----
  for (int w = tid; w < W; w += ORD_T) { occ_bits[(size_t)f * W + w] = occ[ord_swz<SWZ>(w)]; cont_bits[(size_t)f * W + w] = cont[ord_swz<SWZ>(w)]; }
  {
    const int wpt = (W + ORD_T - 1) / ORD_T;                        // consecutive words per thread
    const int w0 = min(tid * wpt, W), w1 = min(w0 + wpt, W);
    unsigned loc = 0;
    for (int w = w0; w < w1; w++) loc += __popc(cont[ord_swz<SWZ>(w)]);
    unsigned incl = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) misc[wid] = incl;
    __syncthreads();                                                // also: every thread is done reading occ
    unsigned wbase = 0, total = 0;
    for (int w = 0; w < ORD_T / 32; w++) { const unsigned t = misc[w]; if (w < wid) wbase += t; total += t; }
    unsigned run = wbase + incl - loc;
    for (int w = w0; w < w1; w++) { occ[ord_swz<SWZ>(w)] = run; cont_pre[(size_t)f * W + w] = run; run += __popc(cont[ord_swz<SWZ>(w)]); }
    // the serial loop's last writer (:102-116) = largest input index: one global atomicMax per contended point
    for (unsigned j = tid; j < total; j += ORD_T) cwin[(size_t)f * cw_stride + j] = 0u;
    if (total == 0) return;                                         // uniform
  }
  __syncthreads();
  scan(visit2);
}

// grid (ceil(max(max_n, S) / (SCAT_T * SCAT_PPT)), F), block SCAT_T.  A thread handles SCAT_PPT points, SCAT_T apart (every warp
// access stays coalesced), and issues every load - the point's fields, its occupancy word, its "contended" bit - before it uses
// the first.  One point per thread is the measured optimum: 1.16 us per frame, against 1.59 / 1.45 / 1.91 for 2 / 4 / 8 points
// per thread (more scattered 16-byte stores in flight per SM make the kernel slower, not faster: it is bound by what the L2
// does with them, not by the latency of its own chain offsets -> point -> store; on slot-ordered input, where the stores
// coalesce, it takes 1.08).
#ifndef SCAT_T
#define SCAT_T 128   // measured: 1.21 / 1.205 / 1.264 / 1.35 us per frame for 64 / 128 / 256 / 512 threads
#endif
#ifndef SCAT_PPT
#define SCAT_PPT 1
#endif
#ifndef SCAT_PRED_LOADS
#define SCAT_PRED_LOADS 0   // measured (profiles/r2_notes.md): 1.139 us per frame with the branch, 1.160 with predicated loads
#endif

template <bool PACKED>   // PACKED: `inten` points at the u32 meta array (slot | flags); row / col / label are unused
__global__ void __launch_bounds__(SCAT_T) k_order_scatter(SensorDev sp, Xform xf, const int64_t* __restrict__ offs, int frame0, int cw_stride,
                                                        const float* __restrict__ x, const float* __restrict__ y,
                                                        const float* __restrict__ z, const float* __restrict__ inten,
                                                        const uint16_t* __restrict__ row, const uint16_t* __restrict__ col,
                                                        const int16_t* __restrict__ label, const uint32_t* __restrict__ occ_bits,
                                                        const uint32_t* __restrict__ cont_bits, const uint32_t* __restrict__ cont_pre,
                                                        const uint32_t* __restrict__ cwin, float4* __restrict__ rec,
                                                        uint32_t* __restrict__ winner_bits, int64_t qbase,
                                                        const uint32_t* __restrict__ cpt_bits) {
  constexpr int P = SCAT_PPT;
  const int f = blockIdx.y;
  const int64_t o = offs[f];
  const int n = (int)(offs[f + 1] - o);
  const int i0 = blockIdx.x * (SCAT_T * P) + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const size_t fb = (size_t)f * sp.S;
  const int W = (sp.S + 31) >> 5;
  const size_t fw = (size_t)f * W;
  uint32_t* wb = winner_bits + (o >> 5) + (frame0 + f);            // this frame's winner words (see bevgen.h)
  if (i0 == 0)    // words between this frame's bits and the next frame's first word are defined as zero
    for (int64_t w = (n + 31) >> 5; w < ((o + n) >> 5) + 1 - (o >> 5); w++) wb[w] = 0u;
  // ---- every load of the thread's points (99.5 % of the points win, so nothing waits for the winner test) ----
  float px[P], py[P], pz[P]; unsigned meta[P], rr[P], cc[P], occw[P], cptw[P]; float pin[P]; int lbl[P];
#pragma unroll
  for (int u = 0; u < P; u++) {
    const int i = i0 + u * SCAT_T;
    px[u] = py[u] = pz[u] = pin[u] = 0.f; meta[u] = META_SLOT; rr[u] = cc[u] = 0xFFFFu; lbl[u] = 0; cptw[u] = 0u;
#if SCAT_PRED_LOADS
    // Experiment, off by default: predicated loads in ONE basic block.  With `if (i < n) { loads }` the compiler evaluates the
    // occupancy test in front of the branch, so a thread waits for its occupancy word before it issues the point's loads (ncu
    // source view: 23 % of the kernel's stall samples on that shift, two memory round trips in series per thread).  Taking
    // that round trip out changes nothing (1.160 against 1.139 us per frame): the kernel is bound by the rate at which the L2
    // takes scattered 16-byte stores, not by the latency of a thread's own chain.
    const int64_t q = o + i - qbase;
    ld_f32_cs_if(px[u], x + o + i, i, n); ld_f32_cs_if(py[u], y + o + i, i, n); ld_f32_cs_if(pz[u], z + o + i, i, n);
    if (PACKED) ld_u32_cs_if(meta[u], reinterpret_cast<const uint32_t*>(inten) + o + i, i, n);
    else {
      ld_u16_nc_if(rr[u], row + o + i, i, n); ld_u16_nc_if(cc[u], col + o + i, i, n);
      ld_f32_cs_if(pin[u], inten + o + i, i, n); ld_s16_cs_if(lbl[u], label + o + i, i, n);
    }
    ld_u32_nc_if(cptw[u], cpt_bits + (q >> 5), i, n);
    cptw[u] >>= (q & 31);
    occw[u] = 0xFFFFFFFFu;
    ld_u32_nc_if(occw[u], occ_bits + fw + (i >> 5), i, sp.S);
#else
    occw[u] = i < sp.S ? __ldg(occ_bits + fw + (i >> 5)) : 0xFFFFFFFFu;
    if (i < n) {
      px[u] = __ldcs(x + o + i); py[u] = __ldcs(y + o + i); pz[u] = __ldcs(z + o + i);
      if (PACKED) meta[u] = __ldcs(reinterpret_cast<const uint32_t*>(inten) + o + i);
      else { rr[u] = row[o + i]; cc[u] = col[o + i]; pin[u] = __ldcs(inten + o + i); lbl[u] = __ldcs(label + o + i); }
      const int64_t q = o + i - qbase;
      cptw[u] = __ldg(cpt_bits + (q >> 5)) >> (q & 31);
    }
#endif
  }
#pragma unroll
  for (int u = 0; u < P; u++) {
    const int i = i0 + u * SCAT_T;
    if (i < sp.S && !((occw[u] >> (i & 31)) & 1u)) rec[fb + i] = make_float4(0.f, 0.f, 0.f, 0.f);   // :98
    if (i - lane >= n) continue;                                    // whole warp past the end (warp-uniform: the ballot needs all lanes)
    unsigned lb16, slot; bool neg1, valid;
    if (PACKED) {
      slot = meta[u] & META_SLOT; valid = slot < (unsigned)sp.S;
      neg1 = (meta[u] & META_NEG1) != 0u; lb16 = (meta[u] & META_LABELED) ? 1u : 0u;    // the device only ever tests label != 0
      if (!valid) slot = 0u;
    } else {
      valid = rr[u] < (unsigned)sp.N && cc[u] < (unsigned)sp.H;      // :106-109
      slot = valid ? rr[u] * sp.H + cc[u] : 0u;
      neg1 = pin[u] == -1.0f; lb16 = (unsigned)(uint16_t)(int16_t)lbl[u];
    }
    bool win = valid;
    if (valid && (cptw[u] & 1u)) {   // contended slot (rare): the serial loop's last writer = largest index
      const unsigned bit = 1u << (slot & 31);
      const uint32_t cw = cont_bits[fw + (slot >> 5)];
      win = cwin[(size_t)f * cw_stride + cont_pre[fw + (slot >> 5)] + __popc(cw & (bit - 1u))] == (uint32_t)i + 1u;
    }
    const unsigned wm = __ballot_sync(0xffffffffu, win);
    if (lane == 0) wb[i >> 5] = wm;
    if (!win) continue;
    float vx = px[u], vy = py[u], vz = pz[u];
    if (xf.on) {   // pcl::transformPointCloud, PCL >= 1.9 SSE order (CloudManip.cpp:128)
      const float ox = __fadd_rn(__fmul_rn(vx, xf.m[0]), __fadd_rn(__fmul_rn(vy, xf.m[1]), __fadd_rn(__fmul_rn(vz, xf.m[2]), xf.m[3])));
      const float oy = __fadd_rn(__fmul_rn(vx, xf.m[4]), __fadd_rn(__fmul_rn(vy, xf.m[5]), __fadd_rn(__fmul_rn(vz, xf.m[6]), xf.m[7])));
      const float oz = __fadd_rn(__fmul_rn(vx, xf.m[8]), __fadd_rn(__fmul_rn(vy, xf.m[9]), __fadd_rn(__fmul_rn(vz, xf.m[10]), xf.m[11])));
      vx = ox; vy = oy; vz = oz;
    }
    const unsigned w = lb16 | W_OWNED | (neg1 ? W_NEG1 : 0u);
    rec[fb + slot] = make_float4(vx, vy, vz, __uint_as_float(w));
  }
}

// ------------------------------------------------------------------------------------------------------------
// K1 ground_mark — markGroundPoints loop 1 (BatchMultiBevGen.cpp:139-184).  One thread per range-image column,
// walking rows N-1 .. N-G like the reference; consecutive lanes = consecutive columns, so every record load is a
// coalesced 512-byte warp access and the "upper" record is reused as the next "lower".
// Closed form of the row-descending overwrite order:  gm[r] = -1 if invalid(r) else (ground(r) | ground(r+1)).
// Emits gz and the ground_mat bits for rows [N-G-1, N), the per-group segment summaries and the count of zero-height
// ground slots per sector (loop 2's `num`, :205; the fold counts the others).
// grid (ceil(H/GM_T), F), block GM_T.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_neg1(const float4& p) { return (__float_as_uint(p.w) & W_NEG1) != 0; }

// DBL: also evaluate the double overload set (bevgen_set_libm / bevgen_set_diag); the default instantiation carries none of it.
template <bool DBL>
__device__ __forceinline__ bool ground_decision(const SensorDev& sp, const float4& up, const float4& lo) {
  const float dx = __fsub_rn(up.x, lo.x), dy = __fsub_rn(up.y, lo.y), dz = __fsub_rn(up.z, lo.z);      // :169-171
  const float hh = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  const float zz = __fmul_rn(dz, dz);
  // Pre-filter on squares (no sqrt, no division): tan^2 thresholds carry a 2e-5 relative guard band, the three
  // roundings involved are ~2e-7, and glibc's atan2f is accurate to < 1 ulp, so outside the band the sign of
  // |atan2f(dz, sqrtf(hh))| - t_star is decided.  Only taken when nothing can overflow / underflow.
  // Two vertically adjacent empty slots (all-zero records, :98) are the common degenerate pair - one in a hundred pairs,
  // i.e. a quarter of all warps hold one: sqrtf(0) = 0 and atan2f(dz, +0) is +-0 for dz = +-0 (ground, :173 atan2(0,0) = 0)
  // and +-pi/2 for any other dz, NaN excluded by the comparison.  (hh is exactly 0 only if both products are +-0.)
  if (hh == 0.0f) return dz == 0.0f;
  const bool in_range = hh > 1e-30f && hh < 1e30f && zz < 1e30f;
  if (in_range) {
    if (zz <= __fmul_rn(sp.q_lo2, hh)) return true;
    if (zz >= __fmul_rn(sp.q_hi2, hh)) return false;
  }
  const float hyp = __fsqrt_rn(hh);                                                                  // :173 sqrtf
  const float t = atan2f_glibc(dz, hyp);   // borderline / special values: the exact libm value decides
  const bool gf = fabsf(t) <= sp.t_star;   // <=> fabsf((float)((double)t*180.0/M_PI)) <= 10.0f  (:173,:179)
  if (!DBL) return gf;
  // The other overload set (no <math.h> in the include tree): atan2(double, double), sqrt(double) of the float sum, the
  // product rounded to float on assignment.  CUDA's double atan2 is within 2 ulp of glibc's; the float rounding of the
  // angle absorbs that except within ~1e-15 relative of a rounding boundary.
  const double ad = __ddiv_rn(__dmul_rn(atan2((double)dz, sqrt((double)hh)), 180.0), 3.14159265358979323846);
  const bool gd = fabsf(__double2float_rn(ad)) <= 10.0f;
  if (sp.diag != nullptr) {
    if (in_range) atomicAdd(&sp.diag[0], 1ull);
    if (gf != gd) atomicAdd(&sp.diag[1], 1ull);
  }
  return sp.libm_double ? gd : gf;
}

constexpr int GM_T = 64;   // columns per CTA: H = 2083 columns fill 33 CTAs of 64 to 98.6 % (17 of 128: 95.7 %)
template <bool DBL>
__global__ void __launch_bounds__(GM_T) k_ground_mark(SensorDev sp, const float4* __restrict__ rec, float* __restrict__ gz,
                                                      uint32_t* __restrict__ cnt, uint4* __restrict__ gsum) {
  const int f = blockIdx.y;
  const int c0 = blockIdx.x * GM_T + threadIdx.x;
  const bool act = c0 < sp.H;
  const int c = act ? c0 : sp.H - 1;
  const int lane = threadIdx.x & 31;
  const int H = sp.H, N = sp.N;
  const size_t fb = (size_t)f * sp.S;
  uint32_t* const cntf = cnt + (size_t)f * NSECT;
  // everything is addressed relative to the current row and walked upwards by -H per iteration
  const float4* pr = rec + fb + (size_t)(N - 1) * H + c;        // (row r, col c)
  float* gzp = gz + fb + (size_t)(N - 1) * H + c;
  const int dplus = (c + 2 >= H ? c + 2 - H : c + 2) - c;       // (col+2) % H, relative to c           (:147)
  const int dminus = -2;                                        // (col-2) % H stays negative in C++ for col < 2 (:152)

  // per (row, 32-column group): summary for the segment form of loop 2 (k_seg_build) and the ground_mat == 1 bits
  // loop 3 needs (k_finalize_bin)
  const int NG = (H + 31) >> 5;
  const size_t g0 = ((size_t)f * (sp.G + 1) + sp.G) * NG + (c0 >> 5);     // row N-1 first, walked upwards by -NG
  uint4* gs = gsum + g0;
  const unsigned lt = (1u << lane) - 1u;

  auto emit = [&](const float4& p, bool gm1) {
    const bool g1 = act && gm1;
    if (act) *gzp = gm1 ? p.z : 0.0f;
    // Loop 2's float sums (:198) only change when a non-zero height is added, so the "participating" slots are the
    // ground slots with z != 0 (NaN participates).  pm = participating lanes, hm = lanes whose sector differs from the
    // previous participating lane of this group, fk / lk = sector of the first / last participating lane.
    const bool part = g1 && p.z != 0.0f;
    const unsigned k2 = g1 ? sector_of(p.x, p.y) : NO_KEY;
    // loop 2's count (:205) is order-free and split in two: ground slots with a non-zero height are counted by the fold
    // (they are exactly the non-zero entries of gz inside the sector's segments); the rare ground slots of height 0
    // (pairs of empty slots, :173 atan2(0,0) = 0) take part in no segment and are counted here.
    if (g1 && !part) atomicAdd(&cntf[k2], 1u);
    const unsigned gmm = __ballot_sync(0xffffffffu, g1);
    const unsigned pm = __ballot_sync(0xffffffffu, part);
    const unsigned below = pm & lt;
    const unsigned pk = __shfl_sync(0xffffffffu, k2, (31 - __clz(below)) & 31);
    const unsigned hm = __ballot_sync(0xffffffffu, part && below != 0u && pk != k2);
    // summary of the (row, group): x = participating lanes, y = sector-change lanes, z = first | last sector << 16 (stored by
    // those two lanes themselves; measured: fetching them to lane 0 with two shuffles for one 16-byte store is slower),
    // w = the ground_mat == 1 bits of the group (k_seg_build re-packs those in slot order for k_finalize_bin)
    if (c0 - lane < H) {                                         // the group exists (warp-uniform)
      if (lane == 0) { gs->x = pm; gs->y = hm; gs->w = gmm; }
      uint16_t* fl = reinterpret_cast<uint16_t*>(&gs->z);
      if (part && below == 0u) fl[0] = (uint16_t)k2;
      if (part && (pm >> lane) == 1u) fl[1] = (uint16_t)k2;
    }
  };

  float4 lower = pr[0];
  float4 nxt = pr[-H];
  bool ground_prev = false;
  for (int r = N - 1; r > N - sp.G - 1; --r) {
    const float4 direct = nxt;
    if (r >= 2) nxt = pr[-2 * H];                                 // next iteration's upper: in flight during this row's math
    float4 up = direct;
    if (is_neg1(up)) up = pr[-H + dplus];                         // :146-149  (predicated loads; a warp-uniform skip of the
    if (is_neg1(up)) up = pr[-H + dminus];                        // :151-154   chain behind __any_sync measured slower)
    if (is_neg1(up) && r >= 2) up = pr[-2 * H];                   // :157-160
    const bool invalid = is_neg1(lower) || is_neg1(up);           // :162
    const bool ground = !invalid && ground_decision<DBL>(sp, up, lower);
    emit(lower, !invalid && (ground || ground_prev));
    ground_prev = ground;
    lower = direct;
    pr -= H; gzp -= H; gs -= NG;
  }
  emit(lower, ground_prev);   // row above the band only receives gm[row-1] = 1 (:181)
}

// ------------------------------------------------------------------------------------------------------------
// K2 sector_mean — markGroundPoints loop 2 + divide (:187-210).  The reference accumulates
// ground_grid_avg_heights[sector] += z serially in row-major slot order: float addition is not associative, so the
// per-sector order must be kept.  One warp sweeps one frame in slot order, 32 slots per step; inside a step each
// sector group (match_any) is folded sequentially by its lowest lane from shared-memory accumulators.  Adding
// +-0 never changes a (never -0) sum, so empty/zero-height ground slots are skipped; the count is order-free: K1
// counted the zero-height ground slots, the others are counted here.  num = 0.01f + 1 + 1 ... is a pure function of the count: cnt_lut[n].
// grid F, block 32, dynamic smem 2*NSECT*4.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_sector_mean(SensorDev sp, const float4* __restrict__ rec,
                                                     const float* __restrict__ gz, const uint32_t* __restrict__ cnt,
                                                     const float* __restrict__ cnt_lut, float* __restrict__ avg,
                                                     const uint32_t* __restrict__ slow_flag) {
  extern __shared__ float ssum[];                 // [NSECT] running sums of this frame, then [NSECT] counts of non-zero heights
  uint32_t* scnt = reinterpret_cast<uint32_t*>(ssum + NSECT);
  if (slow_flag && !slow_flag[blockIdx.x]) return; // the segment form (k_seg_build + k_seg_fold) already did this frame
  __shared__ __align__(16) float zb[2][32];       // the current step's 32 heights (double-buffered)
  const int f = blockIdx.x, lane = threadIdx.x;
  for (int i = lane; i < NSECT; i += 32) { ssum[i] = 0.0f; scnt[i] = 0u; }
  __syncwarp();
  const size_t fb = (size_t)f * sp.S;
  const float4* R = rec + fb;
  const float* Z = gz + fb;
  constexpr int U = 4;
  unsigned k[U], kn[U]; float zz[U], zn[U];
  // gz is the slot's height where ground_mat == 1 and 0 elsewhere, so "participates in a sum" <=> gz != 0 (NaN included);
  // the sector is recomputed from the record (this sweep is the rare fallback; the segment form never reads it)
  auto load = [&](int base, unsigned* kk, float* zv) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int idx = base + u * 32 + lane;
      const bool in = idx < sp.S;
      zv[u] = in ? Z[idx] : 0.0f;
      kk[u] = NO_KEY;
      if (in && zv[u] != 0.0f) { const float4 p = R[idx]; kk[u] = sector_of(p.x, p.y); }
    }
  };
  int base = sp.band_row0 * sp.H;
  int buf = 0;
  load(base, kn, zn);
  for (; base < sp.S; base += 32 * U) {
#pragma unroll
    for (int u = 0; u < U; u++) { k[u] = kn[u]; zz[u] = zn[u]; }
    if (base + 32 * U < sp.S) load(base + 32 * U, kn, zn);     // next group's loads fly during this group's chains
#pragma unroll
    for (int u = 0; u < U; u++) {
      const bool valid = k[u] != NO_KEY;
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      if (vm == 0) continue;
      // Group the lanes by sector (exact, also for A B A patterns).  The group's lowest lane leads; that test is plain
      // mask logic - bit-scan ops (FLO: ffs/clz/popc) run on the quarter-rate XU pipe and measurably stall this
      // latency-bound single-warp kernel, so the sweep contains none.
      const unsigned peers = __match_any_sync(0xffffffffu, valid ? k[u] : NO_KEY);
      const bool leader = valid && (peers & ((1u << lane) - 1u)) == 0u;
      const unsigned mine = leader ? peers : 0u;       // lanes whose z this lane folds, in lane (= slot) order
      // Stage the 32 heights once; every leader then reads them back as broadcast 128-bit words (all leaders read the
      // same address => one shared-memory wavefront per read) instead of 32 warp shuffles through the same MIO pipe.
      zb[buf][lane] = zz[u];                           // gz is 0 for non-ground slots; adding +-0 is a no-op
      __syncwarp();
      float acc = leader ? ssum[k[u]] : 0.0f;
      const float4* z4 = reinterpret_cast<const float4*>(zb[buf]);
      float4 v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = z4[q];        // 8 independent broadcast reads, issued back to back
      // the only serial dependence is the leader's chain acc = fl(acc + z) in slot order (:198); selects are off it
#pragma unroll
      for (int q = 0; q < 8; q++) {
        acc = __fadd_rn(acc, (mine & (1u << (4 * q + 0))) ? v[q].x : 0.0f);
        acc = __fadd_rn(acc, (mine & (1u << (4 * q + 1))) ? v[q].y : 0.0f);
        acc = __fadd_rn(acc, (mine & (1u << (4 * q + 2))) ? v[q].z : 0.0f);
        acc = __fadd_rn(acc, (mine & (1u << (4 * q + 3))) ? v[q].w : 0.0f);
      }
      const unsigned nzm = __ballot_sync(0xffffffffu, valid && zz[u] != 0.0f);
      if (leader) { ssum[k[u]] = acc; scnt[k[u]] += (unsigned)__popc(peers & nzm); }
      buf ^= 1;
      __syncwarp();
    }
  }
  for (int i = lane; i < NSECT; i += 32)
    avg[(size_t)f * NSECT + i] = __fdiv_rn(ssum[i], cnt_lut[cnt[(size_t)f * NSECT + i] + scnt[i]]);   // :210 IEEE divide
}

// ------------------------------------------------------------------------------------------------------------
// K2' sector_mean, segment form — the same order-exact sums as k_sector_mean, parallel across sectors instead of
// along the slot order.  A sector's chain  sum = fl(sum + z)  (:198) only has to see ITS ground slots in row-major
// order; chains of different sectors are independent.  k_ground_mark leaves one summary per (row, 32-column group);
// from those this kernel cuts the band into segments = maximal stretches whose participating slots (ground, z != 0)
// all lie in one sector.  gz is 0 on every other slot, and adding +-0 never changes a sum that starts at +0, so a
// segment is folded by simply adding gz[start..end].  k_seg_build (one CTA per frame) buckets the segments per
// sector in slot order (ordered list + one warp assigning ranks with match_any) and leaves the lists in global
// memory; k_seg_fold then runs one LANE per sector over its segments.
// HDL_64E synthetic frames: ~2.7 k segments, ~350 active sectors, longest chain ~3 k additions.
// Frames with more than `cap` segments (or a segment longer than 65535 slots) raise slow_flag[f] and are handled by
// k_sector_mean (the sweep form).
// k_seg_build: grid F, block SEGT, dynamic smem SMEM_SEG.
// ------------------------------------------------------------------------------------------------------------
#ifndef SEG_MIN_CTAS
#define SEG_MIN_CTAS 2
#endif
constexpr int SEG_CAP = 4096;          // segments a frame may have for the two-CTAs-per-SM build
constexpr int SEG_CAP_BIG = 12288;     // ... for the one-CTA-per-SM build that takes the frames above it (one synthetic HDL_64E frame in
                                       // two hundred has more than 4096 segments, noisier scenes have 6 k - 12 k; without this they fell
                                       // to the one-warp sweep, which made the sector-mean stage of a whole wave 4.6 times slower)
constexpr int SEG_STRIDE = SEG_CAP_BIG;   // entries per frame of the segment lists in global memory
constexpr int SEGT = 512;
// the big build keeps its bucket positions (s_order) in the dead half of s_endtmp: 14 instead of 16 bytes per segment
__host__ __device__ constexpr int seg_smem_bytes(int cap) {
  return cap * 4 * 2 + NSECT * 4 + SEGT * 8 + 320 + cap * 2 * (cap == SEG_CAP ? 3 : 2) + (NSECT + 2) * 2 + 8 + cap * 2;
}
constexpr int SMEM_SEG = seg_smem_bytes(SEG_CAP);          // 92,464 B (two CTAs per SM)
constexpr int SMEM_SEG_BIG = seg_smem_bytes(SEG_CAP_BIG);  // 198,960 B
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifndef FOLD_STEP_N
#define FOLD_STEP_N 32   // 16 is 10 % slower
#endif
constexpr int FOLD_STEP = FOLD_STEP_N;           // heights per step of a chain in k_seg_fold
constexpr int FOLD_PASSES = 12;         // warps per frame in k_seg_fold (32 sectors each; more sectors wrap around)

template <int CAP>   // SEG_CAP: every frame; SEG_CAP_BIG: only the frames the first pass flagged (slow_flag == 2)
__global__ void __launch_bounds__(SEGT, CAP == SEG_CAP ? SEG_MIN_CTAS : 1) k_seg_build(SensorDev sp, int cap, const uint4* __restrict__ gsum,
                                                     const float4* __restrict__ rec, float* __restrict__ avg,
                                                     uint32_t* __restrict__ slow_flag, uint32_t* __restrict__ seg_start,
                                                     uint16_t* __restrict__ seg_len, uint32_t* __restrict__ kdesc,
                                                     uint16_t* __restrict__ act, uint32_t* __restrict__ n_act_out,
                                                     uint32_t* __restrict__ gmask, uint32_t* __restrict__ cnt) {
  extern __shared__ __align__(16) unsigned char seg_smem[];
  uint32_t* s_start = reinterpret_cast<uint32_t*>(seg_smem);   // [CAP] first slot of the segment
  uint32_t* s_endtmp = s_start + CAP;                       // [CAP] last participating slot of the segment
  uint32_t* s_kcnt = s_endtmp + CAP;                        // [NSECT] segments per sector
  uint32_t* s_lk = s_kcnt + NSECT;                              // [SEGT] sector of the last participating slot of the thread's groups
  int* s_lp = reinterpret_cast<int*>(s_lk + SEGT);              // [SEGT] its slot (relative to the frame), -1 if none
  uint32_t* s_scan = reinterpret_cast<uint32_t*>(s_lp + SEGT);  // [32]
  uint32_t* s_warp = s_scan + 32;                               // [32]
  uint32_t* s_misc = s_warp + 32;                               // [16]
  uint16_t* s_len = reinterpret_cast<uint16_t*>(s_misc + 16);   // [CAP] last participating slot - first slot
  uint16_t* s_key = s_len + CAP;                            // [CAP] sector of the segment
  // [CAP] segment ids bucketed by sector, slot order kept.  The big build places it behind the first NSECT words of s_endtmp,
  // which is dead once the segment lengths exist (those NSECT words later hold s_span)
  constexpr bool ALIAS = CAP != SEG_CAP;
  static_assert(!ALIAS || CAP * 4 >= NSECT * 4 + CAP * 2, "s_order does not fit behind s_span inside s_endtmp");
  uint16_t* s_order = ALIAS ? reinterpret_cast<uint16_t*>(s_endtmp + NSECT) : s_key + CAP;
  uint16_t* s_kbase = ALIAS ? s_key + CAP : s_order + CAP;  // [NSECT + 2] bucket base, later bucket end
  // [CAP / 2] participating slots (ground, z != 0) of every segment, two 16-bit counts per word (smem atomics are 32-bit)
  uint32_t* s_np = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(s_kbase + NSECT + 2 + 3) & ~(uintptr_t)7);

  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (CAP != SEG_CAP && slow_flag[f] != 2u) return;             // second pass: only the frames the first one could not hold (uniform)
  const int H = sp.H, NG = (H + 31) >> 5;
  const int n_groups = (sp.G + 1) * NG;
  const int gpt = (n_groups + SEGT - 1) / SEGT;                 // consecutive groups per thread
  const uint4* GS = gsum + (size_t)f * n_groups;
  const size_t fb = (size_t)f * sp.S;
  const float4* R = rec + fb;
  const int g0 = min(tid * gpt, n_groups), g1 = min(g0 + gpt, n_groups);
  auto slot_of = [&](int g, int l) { const int rb = g / NG, cg = g - rb * NG; return (sp.band_row0 + rb) * H + cg * 32 + l; };
  // exclusive block scan of one value per thread; returns the exclusive prefix, *total = sum over the block
  auto block_scan = [&](unsigned v, unsigned* total) {
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    __syncthreads();
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    unsigned wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SEGT / 32; w++) { const unsigned c = s_warp[w]; if (w < wid) wbase += c; tot += c; }
    *total = tot;
    return wbase + incl - v;
  };

  for (int i = tid; i < NSECT; i += SEGT) s_kcnt[i] = 0;
  for (int i = tid; i < CAP / 2; i += SEGT) s_np[i] = 0u;
  if (tid == 0) s_misc[0] = 0;
  // ---- the ground_mat == 1 bits, re-packed from k_ground_mark's (row, 32-column group) words (summary field w) into one bit per slot in
  // slot order (bit s & 31 of word s >> 5): k_finalize_bin walks the frame 32 consecutive slots per warp and then needs a
  // single broadcast word.  Rows are H columns wide and H is not a multiple of 32, so a word is pieced together from up to
  // four source words.  Independent of everything below; its loads overlap the summary loads of pass 1.
  if (CAP == SEG_CAP) {                                         // (the first pass does it for every frame)
    const int W = (sp.S + 31) >> 5;
    uint32_t* GB = gmask + (size_t)f * W;
    for (int w = tid; w < W; w += SEGT) {
      uint32_t out = 0u;
      int b = 0, slot = w * 32;
      int r = slot / H, c = slot - r * H;
      while (b < 32 && slot < sp.S) {
        const int n = min(32 - b, min(32 - (c & 31), H - c));       // bits this source word supplies
        const int rb = r - sp.band_row0;
        if (rb >= 0) {
          const uint32_t src = GS[rb * NG + (c >> 5)].w >> (c & 31);
          out |= (n == 32 ? src : (src & ((1u << n) - 1u))) << b;
        }
        b += n; slot += n; c += n;
        if (c >= H) { c = 0; r++; }
      }
      GB[w] = out;
    }
  }
  // ---- pass 1: last participating slot / sector of this thread's groups ----
  // the thread's group summaries: the first 8 live in registers (one round trip), further ones are re-read
  uint4 gv[8];
#pragma unroll
  for (int i = 0; i < 8; i++) gv[i] = g0 + i < g1 ? GS[g0 + i] : make_uint4(0u, 0u, 0u, 0u);
  auto summary = [&](int g) -> uint4 {
    const int i = g - g0;
    uint4 r = GS[min(g, n_groups - 1)];                           // only used (and only then waited for) when i >= 8
#pragma unroll
    for (int q = 0; q < 8; q++) if (i == q) r = gv[q];
    return r;
  };
  {
    unsigned lk = NO_KEY; int lp = -1;
    for (int g = g0; g < g1; g++) {
      const uint4 v = summary(g);
      if (v.x) { lk = v.z >> 16; lp = slot_of(g, 31 - __clz(v.x)); }
    }
    s_lk[tid] = lk; s_lp[tid] = lp;
  }
  __syncthreads();
  // carry-in: last participating slot before this thread's first group
  unsigned ck = NO_KEY; int cp = -1;
  for (int t = tid - 1; t >= 0; t--) if (s_lp[t] >= 0) { ck = s_lk[t]; cp = s_lp[t]; break; }
  // ---- pass 2: count heads ----
  unsigned nh = 0;
  {
    unsigned k = ck;
    for (int g = g0; g < g1; g++) {
      const uint4 v = summary(g);
      if (!v.x) continue;
      nh += __popc(v.y) + ((v.z & 0xFFFFu) != k ? 1u : 0u);
      k = v.z >> 16;
    }
  }
  unsigned total = 0;
  const unsigned ebase = block_scan(nh, &total);
  const int nseg = (int)total;
  if (nseg > cap) {                                               // uniform: the frame goes to the second pass if that can hold it
    if (tid == 0) slow_flag[f] = (CAP == SEG_CAP && cap == SEG_CAP && nseg <= SEG_CAP_BIG) ? 2u : 1u;   // (2), else to the sweep kernel (1)
    return;
  }
  // ---- pass 3: emit segments in slot order (lengths follow once every start is known) ----
  {
    unsigned e = ebase;
    unsigned k = ck; int lastp = cp;
    for (int g = g0; g < g1; g++) {
      const uint4 v = summary(g);
      const unsigned pm = v.x;
      if (!pm) continue;
      unsigned hm = v.y;
      if ((v.z & 0xFFFFu) != k) hm |= pm & (0u - pm);             // the first participating lane opens a segment
      const int base = slot_of(g, 0);
      unsigned rest = pm;                                           // participating lanes not yet credited to a segment
      auto credit = [&](unsigned seg, unsigned lanes) { if (lanes) atomicAdd(&s_np[seg >> 1], (unsigned)__popc(lanes) << ((seg & 1u) * 16u)); };
      while (hm) {
        const int l = __ffs(hm) - 1; hm &= hm - 1;
        const unsigned below = pm & ((1u << l) - 1u);
        const int prev_end = below ? base + 31 - __clz(below) : lastp;
        credit(e - 1u, rest & ((1u << l) - 1u));                    // lanes below a head belong to the segment before it
        rest &= ~((1u << l) - 1u);
        s_start[e] = (uint32_t)(base + l);
        if (e > 0) s_endtmp[e - 1] = (uint32_t)prev_end;
        e++;
      }
      credit(e - 1u, rest);
      k = v.z >> 16; lastp = base + 31 - __clz(pm);
    }
  }
  if (tid == 0 && nseg > 0) {
    int lp = -1;
    for (int t = SEGT - 1; t >= 0; t--) if (s_lp[t] >= 0) { lp = s_lp[t]; break; }
    s_endtmp[nseg - 1] = (uint32_t)lp;
  }
  __syncthreads();
  for (int e = tid; e < nseg; e += SEGT) {
    const unsigned st = s_start[e], d = s_endtmp[e] - st;
    const float4 p0 = R[st];                                      // a segment starts on a participating slot: its sector
    const unsigned k = sector_of(p0.x, p0.y);                     // is the segment's (independent loads, all in flight)
    if (d >= 0xFFFFu) s_misc[0] = 1u;                             // a segment of 65535 slots or more: sweep kernel (16-bit length and count)
    s_len[e] = (uint16_t)d; s_key[e] = (uint16_t)k;
    atomicAdd(&s_kcnt[k], 1u);
  }
  __syncthreads();
  if (s_misc[0]) { if (tid == 0) slow_flag[f] = 1u; return; }     // uniform
  if (tid == 0) slow_flag[f] = 0u;
  // ---- exclusive scan of the segment counts over sectors; active sectors compacted in ascending order ----
  constexpr int KPT = (NSECT + SEGT - 1) / SEGT;                  // 8 sectors per thread
  unsigned n_act = 0;
  {
    unsigned loc = 0, actn = 0;
#pragma unroll
    for (int j = 0; j < KPT; j++) { const int k = tid * KPT + j; if (k < NSECT) { const unsigned c = s_kcnt[k]; loc += c; actn += c ? 1u : 0u; } }
    unsigned dummy;
    unsigned run = block_scan(loc, &dummy);
    unsigned arun = block_scan(actn, &n_act);
#pragma unroll
    for (int j = 0; j < KPT; j++) {
      const int k = tid * KPT + j;
      if (k < NSECT) {
        const unsigned c = s_kcnt[k];
        s_kbase[k] = (uint16_t)run;
        if (c) { act[(size_t)f * NSECT + arun] = (uint16_t)k; kdesc[(size_t)f * NSECT + k] = (run << 16) | c; arun++; }
        else avg[(size_t)f * NSECT + k] = 0.0f;                     // 0 / num, num > 0
        run += c;
      }
    }
    if (tid == 0) n_act_out[f] = n_act;
  }
  __syncthreads();
  // ---- one warp walks the ordered segment list and hands out positions inside the sector buckets ----
  if (wid == 0) {
    // The only serial dependence between two batches of 32 segments is the fill level of a sector both touch; the keys, the
    // match and the rank inside the batch do not depend on it, so they are computed one batch ahead (the walk was a quarter
    // of the kernel's samples: every step waited for its own shared-memory load, match and two warp syncs in a row).
    const unsigned ltm = (1u << lane) - 1u;
    unsigned k = lane < nseg ? (unsigned)s_key[lane] : NO_KEY;
    unsigned peers = __match_any_sync(0xffffffffu, k);
    for (int b = 0; b < nseg; b += 32) {
      const int e = b + lane;
      const bool valid = e < nseg;
      const unsigned kc = k, pc = peers;
      const int en = e + 32;                                        // next batch: key and match in flight during this one
      k = en < nseg ? (unsigned)s_key[en] : NO_KEY;
      peers = __match_any_sync(0xffffffffu, k);
      const unsigned pos = valid ? (unsigned)s_kbase[kc] + __popc(pc & ltm) : 0u;
      __syncwarp();
      if (valid && (pc >> lane) == 1u) s_kbase[kc] = (uint16_t)(pos + 1u);  // the group's highest lane publishes the new fill level
      if (valid) s_order[e] = (uint16_t)pos;
      __syncwarp();
    }
  }
  __syncthreads();
  // ---- the bucketed list goes to global memory: entry `pos` of the frame = (first slot, length - 1) ----
  uint32_t* s_span = s_endtmp;                                    // [NSECT] slots a sector's chain walks over (its cost)
  uint32_t* s_knp = s_kcnt;                                       // [NSECT] participating slots of the sector (s_kcnt is dead: kdesc holds the counts)
  for (int k = tid; k < NSECT; k += SEGT) { s_span[k] = 0u; s_knp[k] = 0u; }
  __syncthreads();
  for (int e = tid; e < nseg; e += SEGT) {
    const unsigned pos = s_order[e];
    seg_start[(size_t)f * SEG_STRIDE + pos] = s_start[e];
    seg_len[(size_t)f * SEG_STRIDE + pos] = s_len[e];
    atomicAdd(&s_span[s_key[e]], (unsigned)s_len[e] + 1u);
    atomicAdd(&s_knp[s_key[e]], (s_np[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu);
  }
  __syncthreads();
  // loop 2's count (:205) is order-free: ground slots of height 0 were counted by k_ground_mark, the participating ones
  // (exactly the non-zero heights the fold adds up) are known from the summaries - the fold itself counts nothing
  for (int k = tid; k < NSECT; k += SEGT) { const unsigned c = s_knp[k]; if (c) cnt[(size_t)f * NSECT + k] += c; }
  __syncthreads();
  // ---- active sectors sorted by chain cost, longest first: k_seg_fold gives 32 consecutive entries to one warp, and a
  // warp runs as long as its longest chain ----
  if (n_act > 1 && n_act <= 1024) {
    uint32_t* c_span = s_kcnt;                                    // s_kcnt is dead (kdesc holds the counts)
    uint16_t* c_key = reinterpret_cast<uint16_t*>(s_kcnt + 1024);
    for (unsigned a = tid; a < n_act; a += SEGT) { const unsigned k = act[(size_t)f * NSECT + a]; c_key[a] = (uint16_t)k; c_span[a] = s_span[k]; }
    __syncthreads();
    for (unsigned a = tid; a < n_act; a += SEGT) {
      const unsigned mine = c_span[a];
      unsigned r = 0;
      for (unsigned b = 0; b < n_act; b++) { const unsigned o = c_span[b]; r += (o > mine || (o == mine && b < a)) ? 1u : 0u; }
      act[(size_t)f * NSECT + r] = c_key[a];
    }
  }
}

// k_seg_fold — one LANE per active sector runs the sector's serial chain sum = fl(sum + z) (:198) over its segments in
// slot order, straight from gz.  The lane walks a flat sequence of windows of FOLD_STEP heights (32-byte aligned; heights
// outside the segment are skipped by a window mask), so the 32 chains of a warp stay in lockstep whatever their segment
// boundaries are.  Then the IEEE divide (:210).
//
// The chain is latency bound (ncu: 15 % warps active; a frame's longest chain alone is ~165 steps of one warp), so the
// address stream is decoupled from the adds: a position `pa` runs FOLD_DIST windows ahead of the chain and only issues
// prefetches (the walk over segment descriptors does not depend on the sums), the chain's own four 256-bit loads per window
// find their lines in L1, and every position holds the descriptor of its NEXT segment from the moment it enters the current
// one.  One window buffer (61 registers; the double-buffered form needed 122 and kept the other wave's kernels off the SM).
// grid (F, FOLD_PASSES / wpb), block (32, wpb): warp p of a frame owns the active sectors p*32 + lane (+ 32*FOLD_PASSES ...);
// the list is sorted by cost, so a frame's pass 0 holds its 32 longest chains.
#ifndef FOLD_DIST_N
#define FOLD_DIST_N 2
#endif
#ifndef FOLD_SINGLE_BUFFER
#define FOLD_SINGLE_BUFFER 1
#endif
constexpr int FOLD_DIST = FOLD_DIST_N;
constexpr int FOLD_MAX_WPB = 4;

// 8 consecutive floats as one 256-bit load (sm_100: LDG.E.256; the address must be 32-byte aligned).  The fold is bound by
// L1 wavefronts - every lane reads its own line - so the wider the load, the fewer wavefronts per height.
struct __align__(32) F8 { float v[8]; };
__device__ __forceinline__ F8 ldg_f8(const float* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}

#ifndef FOLD_RAW_DESC
#define FOLD_RAW_DESC 0   // measured (profiles/r2_notes.md): sector_mean 0.479 us per frame without, 0.476 with - the wait moves, the chain stays
#endif
struct FoldPos {            // a lane's position in its sector's window sequence
  unsigned cur, end;        // current / one-past-last segment (indices into the frame's bucketed segment list)
  unsigned j, hi;           // remaining slots [j, hi] of the current segment
  unsigned nj, nlen;        // the FOLLOWING segment's start and length, fetched when the current one was entered.  FOLD_RAW_DESC
                            // (experiment, off by default) leaves them untouched until the position moves on: computing
                            // `nj + len` at fetch time makes every segment entry wait for the two loads it has just issued
                            // (ncu source view: 27 % of the kernel's stall samples).  Measured: no change (0.476 against 0.479 us
                            // per frame) - the wait moves to the window loads; the fold is bound by the bytes an SM keeps in
                            // flight (32 one-warp CTAs x one 128-byte window per lane), not by where a warp waits.
};
__device__ __forceinline__ bool fold_valid(const FoldPos& p) { return p.cur < p.end; }
__device__ __forceinline__ void fold_fetch_next(FoldPos& p, const uint32_t* __restrict__ SS, const uint16_t* __restrict__ SL) {
#if FOLD_RAW_DESC
  // predicated loads straight into the loop-carried registers: no copy that would wait for them
  ld_u32_nc_if(p.nj, SS + p.cur + 1, (int)(p.cur + 1), (int)p.end);
  ld_u16_nc_if(p.nlen, SL + p.cur + 1, (int)(p.cur + 1), (int)p.end);
#else
  if (p.cur + 1 < p.end) { p.nj = SS[p.cur + 1]; p.nlen = SL[p.cur + 1]; }
#endif
}
__device__ __forceinline__ void fold_init(FoldPos& p, unsigned first, unsigned count, const uint32_t* __restrict__ SS, const uint16_t* __restrict__ SL) {
  p.cur = first; p.end = first + count; p.j = SS[first]; p.hi = p.j + SL[first]; p.nj = 0u; p.nlen = 0u;
  fold_fetch_next(p, SS, SL);
}
__device__ __forceinline__ void fold_advance(FoldPos& p, const uint32_t* __restrict__ SS, const uint16_t* __restrict__ SL) {
  const unsigned nj = (p.j & ~7u) + FOLD_STEP;
  if (nj <= p.hi) { p.j = nj; return; }
  p.cur++;
  if (p.cur < p.end) {
#if FOLD_RAW_DESC
    // the old descriptor is consumed by (volatile, hence ordered) moves in front of the loads that overwrite its registers:
    // left to itself the compiler loaded into a temporary and copied it over afterwards - a copy that waits for the load
    unsigned len;
    asm volatile("mov.u32 %0, %2;\n\tmov.u32 %1, %3;" : "=r"(p.j), "=r"(len) : "r"(p.nj), "r"(p.nlen));
    p.hi = p.j + len;
#else
    p.j = p.nj; p.hi = p.nj + p.nlen;
#endif
    fold_fetch_next(p, SS, SL);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(32 * FOLD_MAX_WPB) k_seg_fold(SensorDev sp, const float* __restrict__ gz, const uint32_t* __restrict__ cnt,
                                                  const float* __restrict__ cnt_lut, const uint32_t* __restrict__ slow_flag,
                                                  const uint32_t* __restrict__ seg_start, const uint16_t* __restrict__ seg_len,
                                                  const uint32_t* __restrict__ kdesc, const uint16_t* __restrict__ act,
                                                  const uint32_t* __restrict__ n_act_in, float* __restrict__ avg) {
  const int f = blockIdx.x;
  if (slow_flag[f]) return;                                       // the sweep kernel takes this frame
  const unsigned n_act = n_act_in[f];
  const float* Z = gz + (size_t)f * sp.S;
  const uint32_t* SS = seg_start + (size_t)f * SEG_STRIDE;
  const uint16_t* SL = seg_len + (size_t)f * SEG_STRIDE;
  const unsigned pass = blockIdx.y * blockDim.y + threadIdx.y;
  for (unsigned a = pass * 32 + threadIdx.x; a < n_act; a += 32 * FOLD_PASSES) {
    const unsigned k = act[(size_t)f * NSECT + a];
    const unsigned d = kdesc[(size_t)f * NSECT + k];
    float acc = 0.0f;
    if (VEC) {
      static_assert(FOLD_STEP == 32, "the window mask of the 256-bit form is one 32-bit word");
      FoldPos pc; fold_init(pc, d >> 16, d & 0xFFFFu, SS, SL);
      FoldPos pa = pc;
#pragma unroll
      for (int q = 0; q < FOLD_DIST; q++) {
        if (fold_valid(pa)) { prefetch_l1(Z + (pa.j & ~7u)); prefetch_l1(Z + (pa.j & ~7u) + FOLD_STEP - 1); fold_advance(pa, SS, SL); }
      }
#if FOLD_SINGLE_BUFFER
      // one window buffer: the loads find their lines in L1 (the prefetch position runs FOLD_DIST windows ahead), half the
      // registers of the double-buffered form, no window copies
      while (true) {
        F8 w[FOLD_STEP / 8];
        const unsigned jb = pc.j & ~7u;
        const float* src = Z + jb;                                  // gz is padded: the window may pass the frame's end
#pragma unroll
        for (int u = 0; u < FOLD_STEP / 8; u++) w[u] = ldg_f8(src + 8 * u);
        if (fold_valid(pa)) { prefetch_l1(Z + (pa.j & ~7u)); prefetch_l1(Z + (pa.j & ~7u) + FOLD_STEP - 1); fold_advance(pa, SS, SL); }
        const unsigned m = (0xFFFFFFFFu << (pc.j - jb)) & (0xFFFFFFFFu >> (31u - min(pc.hi - jb, 31u)));
#pragma unroll
        for (int u = 0; u < FOLD_STEP / 8; u++) {
#pragma unroll
          for (int e = 0; e < 8; e++)
            if (m & (1u << (8 * u + e))) acc = __fadd_rn(acc, w[u].v[e]);
        }
        fold_advance(pc, SS, SL);
        if (!fold_valid(pc)) break;
      }
#else
      F8 w[FOLD_STEP / 8];
      unsigned wj = pc.j, whi = pc.hi;                              // the window held in w
      {
        const float* src = Z + (pc.j & ~7u);                        // gz is padded: the window may pass the frame's end
#pragma unroll
        for (int u = 0; u < FOLD_STEP / 8; u++) w[u] = ldg_f8(src + 8 * u);
      }
      while (true) {
        FoldPos pn = pc; fold_advance(pn, SS, SL);
        const bool more = fold_valid(pn);
        F8 wn[FOLD_STEP / 8];
        if (more) {
          const float* src = Z + (pn.j & ~7u);
#pragma unroll
          for (int u = 0; u < FOLD_STEP / 8; u++) wn[u] = ldg_f8(src + 8 * u);
        }
        if (fold_valid(pa)) { prefetch_l1(Z + (pa.j & ~7u)); prefetch_l1(Z + (pa.j & ~7u) + FOLD_STEP - 1); fold_advance(pa, SS, SL); }
        // heights outside [wj, whi] belong to other sectors (or to nobody) and are skipped: one predicated add per height
        // (skipping equals adding +0: the sum starts at +0 and can never become -0)
        const unsigned jb = wj & ~7u;
        const unsigned m = (0xFFFFFFFFu << (wj - jb)) & (0xFFFFFFFFu >> (31u - min(whi - jb, 31u)));
#pragma unroll
        for (int u = 0; u < FOLD_STEP / 8; u++) {
#pragma unroll
          for (int e = 0; e < 8; e++)
            if (m & (1u << (8 * u + e))) acc = __fadd_rn(acc, w[u].v[e]);
        }
        if (!more) break;
        pc = pn; wj = pn.j; whi = pn.hi;
#pragma unroll
        for (int u = 0; u < FOLD_STEP / 8; u++) w[u] = wn[u];
      }
#endif
    } else {
      unsigned cur = d >> 16; const unsigned endseg = cur + (d & 0xFFFFu);
      unsigned j = SS[cur], hi = j + SL[cur];
      while (true) {
        float v[FOLD_STEP];
#pragma unroll
        for (int u = 0; u < FOLD_STEP; u++) { const unsigned q = j + u; v[u] = q <= hi ? Z[q] : 0.0f; }
#pragma unroll
        for (int u = 0; u < FOLD_STEP; u++) acc = __fadd_rn(acc, v[u]);
        j += FOLD_STEP;
        if (j <= hi) continue;
        cur++;
        if (cur >= endseg) break;
        j = SS[cur]; hi = j + SL[cur];
      }
    }
    avg[(size_t)f * NSECT + k] = __fdiv_rn(acc, cnt_lut[cnt[(size_t)f * NSECT + k]]);   // :210; cnt = all ground slots of the sector (k_ground_mark + k_seg_build)
  }
}

// num_ground_grid_points replay (:135, :205): lut[n] = fl(...fl(fl(0.01f + 1) + 1)... + 1), n additions.
__global__ void k_build_cnt_lut(int n_max, float* __restrict__ lut) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float c = (float)0.01;
    lut[0] = c;
    for (int i = 1; i <= n_max; i++) { c = __fadd_rn(c, 1.0f); lut[i] = c; }
  }
}

// ------------------------------------------------------------------------------------------------------------
// K3+K4 finalize_bin_scatter — markGroundPoints loop 3 (:216-250) fused with the binning of both BEVs
// (:278-292, :342-356), the scatter and the expansion to the reference's byte layout.
// One CTA per frame owns the frame's whole 224x224 grid in shared memory:
//   occ  3 planes x 50176 bytes: bit (layer&7) of plane (layer>>3) = cell occupied in that layer
//   hgt  50176 bytes: running max height
// updates are check-first (plain read, atomic only if it would change the word), OR / byte-max are order-free so
// the result is deterministic.  Flush = coalesced 16-byte stores: single <- hgt, multi[layer] <- bit ? 255 : 0.
// grid F, block 1024, dynamic smem SMEM_BIN.
// ------------------------------------------------------------------------------------------------------------
constexpr int SAVG_BYTES = 15008;                                   // 3750 floats, padded to 16
constexpr int SMEM_BIN = SAVG_BYTES + 4 * CELLS;                    // 215,712 B

// COMPACT (bevgen_process_host_compact): the outputs leave in the form that crosses PCIe cheapest and the host expands
//   ground_bits  [F][ceil(S/32)] u32  bit = the slot is ground after loop 3 (its label becomes 0, :244-245) instead of label_out
//   multi_planes [F][3][224*224] u8   the three occupancy bit planes as they sit in shared memory (bit l&7 of plane l>>3 =
//                                     layer l occupied) instead of the 24 expanded 0/255 layers
template <bool COMPACT>
__global__ void __launch_bounds__(1024, 1) k_finalize_bin(SensorDev sp, const float4* __restrict__ rec,
                                                           const uint32_t* __restrict__ gmask, const float* __restrict__ avg,
                                                           int16_t* __restrict__ label_out, uint8_t* __restrict__ single,
                                                           uint8_t* __restrict__ multi) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* savg = reinterpret_cast<float*>(smem);
  uint32_t* occ = reinterpret_cast<uint32_t*>(smem + SAVG_BYTES);   // [3][CELL_WORDS]
  uint32_t* hgt = occ + 3 * CELL_WORDS;                             // [CELL_WORDS]
  const int f = blockIdx.x, tid = threadIdx.x;
  const size_t fb = (size_t)f * sp.S;

  for (int i = tid; i < CELL_WORDS; i += 1024) reinterpret_cast<uint4*>(occ)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < NSECT; i += 1024) savg[i] = avg[(size_t)f * NSECT + i];
  __syncthreads();

  const float4* R = rec + fb;
  const int W = (sp.S + 31) >> 5;
  const uint32_t* GM = gmask + (size_t)f * W;                       // ground_mat == 1 after loop 1, one bit per slot (k_ground_mark)
  uint32_t* gbits = reinterpret_cast<uint32_t*>(label_out) + (size_t)f * W;   // COMPACT only
  // Register double buffer: the loads of batch k+1 are in flight while batch k goes through the shared-memory atomics
  // (ncu: the loop was load-batch -> wait -> process, the memory pipe idled during every process phase).
  constexpr int UB = 2;   // measured: 0.87 / 0.81 / 0.86 / 0.92 us per frame for 1 / 2 / 3 / 4 records per thread and batch
  float4 nv[UB]; unsigned ng[UB];
  const int lane = tid & 31;
  auto fetch = [&](int s0) {
#pragma unroll
    for (int u = 0; u < UB; u++) {
      const int sl = s0 + u * 1024;
      nv[u] = sl < sp.S ? __ldcs(R + sl) : make_float4(0.f, 0.f, 0.f, 0.f);       // read once: streaming (evict-first) loads
      // the warp's 32 slots are one aligned word of the ground bits: a single broadcast load
      ng[u] = sl - lane < sp.S ? __ldg(GM + (sl >> 5)) : 0u;
    }
  };
  fetch(tid);
  for (int slot0 = tid; slot0 - (tid & 31) < sp.S; slot0 += 1024 * UB) {   // warp-uniform trip count (the ballot below)
   float4 pv[UB]; unsigned gv[UB];
#pragma unroll
   for (int u = 0; u < UB; u++) { pv[u] = nv[u]; gv[u] = ng[u]; }
   if (slot0 + 1024 * UB < sp.S) fetch(slot0 + 1024 * UB);
#pragma unroll
   for (int u = 0; u < UB; u++) {
    const int slot = slot0 + u * 1024;
    if (slot - (tid & 31) >= sp.S) break;        // whole warp past the end (warp-uniform: the ballot below needs all lanes)
    const bool in = slot < sp.S;
    const float4 p = pv[u];
    int16_t lab = (int16_t)(__float_as_uint(p.w) & 0xFFFFu);
    bool ground = false;
    if (in && ((gv[u] >> lane) & 1u)) {          // ground_mat == 1 after loop 1
      const int sr = sector_axis(p.x, 75.0f, SECT_R), sc = sector_axis(p.y, 50.0f, SECT_C);   // getBelongingGrid, BatchMultiBevGen.h:73-99
      const int key = sr * SECT_C + sc;
      bool cleared = false;
      // neighbour order (-1,0),(0,1),(0,-1),(1,0) (:73-84); (double)(z - avg) > 0.30  <=>  (z - avg) >= 0.3f
      // because 0.3f is the smallest float above the double 0.30 (:236-237)
      if (sr - 1 >= 0)      cleared = __fsub_rn(p.z, savg[key - SECT_C]) >= 0.3f;
      if (!cleared && sc + 1 < SECT_C) cleared = __fsub_rn(p.z, savg[key + 1]) >= 0.3f;
      if (!cleared && sc - 1 >= 0)     cleared = __fsub_rn(p.z, savg[key - 1]) >= 0.3f;
      if (!cleared && sr + 1 < SECT_R) cleared = __fsub_rn(p.z, savg[key + SECT_C]) >= 0.3f;
      if (!cleared) { lab = 0; ground = true; }  // :244-245
    }
    if (COMPACT) {
      const unsigned gb = __ballot_sync(0xffffffffu, ground);   // lanes = 32 consecutive slots (slot & 31 == lane)
      if ((tid & 31) == 0) __stcs(gbits + (slot >> 5), gb);
    } else if (in) {
      __stcs(label_out + fb + slot, lab);     // outputs are never re-read on the device: streaming stores
    }
    if (!in || lab == 0) continue;               // :285 / :349
    const float vx = __fadd_rn(p.x, 112.0f), vy = __fadd_rn(p.y, 112.0f);   // (pi.x + MAX_RANGE) / 1.0f
    // x = round(v + 0.5) in double, valid 0..223  <=>  -1 < v < 223 and then x = floor(v) + 1   (:279-284)
    if (!(vx > -1.0f && vx < 223.0f && vy > -1.0f && vy < 223.0f)) continue;
    const int cell = (__float2int_rd(vx) + 1) * GRID + (__float2int_rd(vy) + 1);
    const int sh = (cell & 3) * 8;
    // single: height = clamp(int((z + 2.0f) * 4.0), 0, 255); the double product of a float by 4 is exact (:345-346)
    const float b = __fmul_rn(__fadd_rn(p.z, 2.0f), 4.0f);
    int h = 0;
    if (b < 2147483648.0f && b > 0.0f) { h = __float2int_rz(b); h = h > 255 ? 255 : h; }
    if (h > 0) {
      uint32_t* wp = &hgt[cell >> 2];
      uint32_t old = *reinterpret_cast<volatile uint32_t*>(wp);
      while (((old >> sh) & 0xFFu) < (uint32_t)h) {
        const uint32_t nw = (old & ~(0xFFu << sh)) | ((uint32_t)h << sh);
        const uint32_t prev = atomicCAS(wp, old, nw);
        if (prev == old) break;
        old = prev;
      }
    }
    // multi: layer = round(z / HEIGHT_RES + 2.0f), half away from zero, valid 0..23 <=> -0.5 < w < 23.5   (:281-284)
    const float w = __fadd_rn(sp.hr_pow2 ? __fmul_rn(p.z, sp.inv_height_res) : __fdiv_rn(p.z, sp.height_res), 2.0f);
    if (w > -0.5f && w < 23.5f) {
      const float t = truncf(w);
      const int layer = (int)t + (__fsub_rn(w, t) >= 0.5f ? 1 : 0);
      uint32_t* wp = &occ[(layer >> 3) * CELL_WORDS + (cell >> 2)];
      const uint32_t bit = (1u << (layer & 7)) << sh;
      if (!(*reinterpret_cast<volatile uint32_t*>(wp) & bit)) atomicOr(wp, bit);
    }
  }
  }
  __syncthreads();

  uint4* so = reinterpret_cast<uint4*>(single + (size_t)f * CELLS);
  for (int i = tid; i < CELL_WORDS / 4; i += 1024) __stcs(so + i, reinterpret_cast<const uint4*>(hgt)[i]);
  if (COMPACT) {   // the three bit planes as they are: 150 528 bytes instead of 1 204 224
    uint4* mo = reinterpret_cast<uint4*>(multi + (size_t)f * 3 * CELLS);
    for (int i = tid; i < 3 * CELL_WORDS / 4; i += 1024) __stcs(mo + i, reinterpret_cast<const uint4*>(occ)[i]);
    return;
  }
  uint4* mo = reinterpret_cast<uint4*>(multi + (size_t)f * LAYERS * CELLS);
  constexpr int Q = CELL_WORDS / 4;   // uint4 per layer = 3136
  // a 16-byte piece of a bit plane is read once and expanded into its eight layers (eight coalesced streaming stores)
  for (int i = tid; i < 3 * Q; i += 1024) {
    const int plane = i >= 2 * Q ? 2 : (i >= Q ? 1 : 0), q = i - plane * Q;
    const uint4 v = reinterpret_cast<const uint4*>(occ + plane * CELL_WORDS)[q];
    uint4* dst = mo + (size_t)plane * 8 * Q + q;
#pragma unroll
    for (int l = 0; l < 8; l++) {
      uint4 o;
      o.x = ((v.x >> l) & 0x01010101u) * 255u; o.y = ((v.y >> l) & 0x01010101u) * 255u;
      o.z = ((v.z >> l) & 0x01010101u) * 255u; o.w = ((v.w >> l) & 0x01010101u) * 255u;
      __stcs(dst + (size_t)l * Q, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// K5 pose labels.  d2 = ((dx*dx) + dy*dy) + dz*dz with d = query - major, all float, exactly
// nanoflann L2_Adaptor::evalMetric for dim 3 (include/nanoflann.hpp:383-407) and getDistance (src/Utility.cpp:43-49).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float d2_pose(float qx, float qy, float qz, const float* m) {
  float dx = __fsub_rn(qx, m[0]), dy = __fsub_rn(qy, m[1]), dz = __fsub_rn(qz, m[2]);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// selectMajorFrames (BatchMultiBevGen.cpp:502-566): an inherently serial greedy scan over keyframes; one warp walks
// it, the 1-NN over the current majors is split across lanes (lowest index wins ties).  grid 1, block 32.
__global__ void __launch_bounds__(32) k_select_major(int K, const float* __restrict__ xyz, float* __restrict__ mpos,
                                                      int32_t* __restrict__ major_idx, int32_t* __restrict__ overlap,
                                                      int32_t* __restrict__ n_major) {
  const int lane = threadIdx.x;
  int M = 0;
  if (K <= 0) { if (lane == 0) *n_major = 0; return; }
  if (lane == 0) { major_idx[0] = 0; mpos[0] = xyz[0]; mpos[1] = xyz[1]; mpos[2] = xyz[2]; overlap[0] = -1; }
  M = 1;
  __syncwarp();
  for (int i = 1; i < K; i++) {
    const float qx = xyz[3 * i], qy = xyz[3 * i + 1], qz = xyz[3 * i + 2];
    const float dl = __fsqrt_rn(d2_pose(qx, qy, qz, mpos + 3 * (M - 1)));   // getDistance to the last major (:527)
    if (dl < 20.0f) { if (lane == 0) overlap[i] = -2; continue; }           // :528
    float best = 3.402823466e+38f; int bj = 0x7fffffff;
    for (int j = lane; j < M; j += 32) {
      float d = d2_pose(qx, qy, qz, mpos + 3 * j);
      if (d < best) { best = d; bj = j; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, s); int oj = __shfl_xor_sync(0xffffffffu, bj, s);
      if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    if (best < 400.0f) { if (lane == 0) overlap[i] = bj; continue; }        // :552
    if (lane == 0) { major_idx[M] = i; mpos[3 * M] = qx; mpos[3 * M + 1] = qy; mpos[3 * M + 2] = qz; overlap[i] = -1; }
    M++;
    __syncwarp();
  }
  if (lane == 0) *n_major = M;
}

// getKeyFrameLabel (BatchMultiBevGen.cpp:575-636): one thread per keyframe row, majors tiled through smem,
// 2-NN with KNNResultSet semantics (strict '>' insertion => lowest index first on ties), weights in the
// reference's float/double mix.  Writes the two non-zeros (sparse) and, if dense != NULL, scatters them into the
// zero-filled dense rows.  grid ceil(rows/128), block 128.
__global__ void __launch_bounds__(128) k_labels(int K, const float* __restrict__ xyz, int M, const int32_t* __restrict__ major_idx,
                                                 const float* __restrict__ mpos, int row_begin, int row_end,
                                                 int32_t* __restrict__ nn_idx, float* __restrict__ nn_w, float* __restrict__ dense) {
  __shared__ float sm[128 * 3];
  const int r = row_begin + blockIdx.x * 128 + threadIdx.x;
  const bool act = r < row_end;
  float qx = 0, qy = 0, qz = 0;
  if (act) { qx = xyz[3 * r]; qy = xyz[3 * r + 1]; qz = xyz[3 * r + 2]; }
  float d0 = 0.0f, d1 = 3.402823466e+38f;  // dists[capacity-1] = max (nanoflann.hpp:164-165); vectors value-init to 0
  int c0 = 0, c1 = 0, count = 0;
  for (int base = 0; base < M; base += 128) {
    const int nt = min(128, M - base);
    __syncthreads();
    for (int t = threadIdx.x; t < nt * 3; t += 128) sm[t] = mpos[3 * base + t];
    __syncthreads();
    if (act) {
      for (int j = 0; j < nt; j++) {
        const float d = d2_pose(qx, qy, qz, sm + 3 * j);
        if (!(d < d1)) continue;                       // `dist < worst_dist` gate (nanoflann.hpp:1358)
        if (count == 0) { d0 = d; c0 = base + j; count = 1; }
        else if (d0 > d) { d1 = d0; c1 = c0; d0 = d; c0 = base + j; count = 2; }
        else { d1 = d; c1 = base + j; count = 2; }
      }
    }
  }
  if (!act) return;
  const int o = r - row_begin;
  float w0, w1; int i1 = c1;
  if (r == major_idx[c0]) { w0 = 1.0f; w1 = 0.0f; i1 = -1; }                     // :616-618
  else {
    w0 = __double2float_rn(1.0 / ((double)d0 + 1e-5));                          // :623
    w1 = __double2float_rn(1.0 / ((double)d1 + 1e-5));                          // :624
    const float s = __fadd_rn(w0, w1);
    w0 = __fdiv_rn(w0, s); w1 = __fdiv_rn(w1, s);
  }
  if (nn_idx) { nn_idx[2 * o] = c0; nn_idx[2 * o + 1] = i1; }
  if (nn_w) { nn_w[2 * o] = w0; nn_w[2 * o + 1] = w1; }
  if (dense) {
    dense[(size_t)o * M + c0] = w0;                                             // :618 / :629
    if (i1 >= 0) dense[(size_t)o * M + i1] = w1;                                // :630 (M == 1: overwrites index 0)
  }
}

// ------------------------------------------------------------------------------------------------------------
// cloud_manip (BASELINE config #5): rigid transform (CloudManip.cpp:128) + saveAsMat max grid (:79-95) of the input
// and of the transformed cloud.  Cells start at 0 and only strictly larger values are stored, so stored values are
// positive floats and an int atomicMax on the bit pattern is an exact, order-free float max.  Lanes that hit the
// same cell are folded first (match_any + redux max) so hot cells cost at most one atomic per warp.
// ------------------------------------------------------------------------------------------------------------
constexpr int MGRID = 201;

// One point's update of one max grid, split in two so that the kernel can have the reads of both grids in flight together:
//   manip_probe : cell + value; when every lane of the warp sits in ONE cell (a heavily contended cell, BASELINE config #5's
//                 stress) the warp's max is taken first (one REDUX) and only lane 0 goes on; then the check-first read - a
//                 cell's max settles after a few arrivals, later points only read it (lanes of one cell read one address)
//   manip_commit: lanes that still have to raise their cell.  A few of them: plain atomics (duplicates among them are rare).
//                 Many: fold the lanes of a cell first (match_any); a reduce with a per-group mask is executed once per
//                 DISTINCT mask of the warp (ncu, round 2: 32 serialised CREDUX per warp on scattered points, 40 % of the
//                 kernel's samples), so lanes alone in their cell skip it.
struct ManipProbe { int cell, vb; bool need; };
__device__ __forceinline__ ManipProbe manip_probe(float px, float py, float pz, bool valid, const int* __restrict__ grid) {
  const float vx = __fadd_rn(px, 100.0f), vy = __fadd_rn(py, 100.0f);
  const float v = __fadd_rn(pz, 2.0f);
  bool in = valid && vx > -1.0f && vx < 200.0f && vy > -1.0f && vy < 200.0f && v > 0.0f;
  const int lane = threadIdx.x & 31;
  ManipProbe r;
  r.cell = in ? (__float2int_rd(vx) + 1) * MGRID + (__float2int_rd(vy) + 1) : -1 - lane;
  r.vb = __float_as_int(v);                          // positive floats order like their int patterns
  const int c0 = __shfl_sync(0xffffffffu, r.cell, 0);
  if (__all_sync(0xffffffffu, r.cell == c0)) { r.vb = __reduce_max_sync(0xffffffffu, r.vb); in = in && lane == 0; }
  r.need = in && *reinterpret_cast<const volatile int*>(&grid[in ? r.cell : 0]) < r.vb;
  return r;
}
__device__ __forceinline__ void manip_commit(const ManipProbe& r, int* __restrict__ grid) {
  const unsigned nm = __ballot_sync(0xffffffffu, r.need);
  if (nm == 0u) return;
  if (__popc(nm) <= 8) { if (r.need) atomicMax(&grid[r.cell], r.vb); return; }
  if (r.need) {
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(nm, r.cell);
    int best = r.vb;
    if (peers != (1u << lane)) best = __reduce_max_sync(peers, best);
    if (lane == __ffs(peers) - 1) atomicMax(&grid[r.cell], best);
  }
}

// ------------------------------------------------------------------------------------------------------------
// batch_cloud_manip's bird-view map - saveAsMat of BatchCloudManip.cpp:201-226 on the ground-removed
// ordered cloud: 201x201 floats, cell <- z + 2.0f when strictly greater (cells start at 0), points with label == 0
// (ground, or empty slot) skipped (:214).  Same order-free int max on positive float patterns as manip_scatter, with
// the frame's whole grid in shared memory.  Reads the labels k_finalize_bin wrote and the ordered cloud `rec`.
// grid F, block 512, dynamic smem SMEM_BVM.
// ------------------------------------------------------------------------------------------------------------
constexpr int SMEM_BVM = MGRID * MGRID * 4;   // 161,604 B

__global__ void __launch_bounds__(512) k_float_bev(SensorDev sp, const float4* __restrict__ rec, const int16_t* __restrict__ label,
                                                    float* __restrict__ bvm) {
  extern __shared__ __align__(16) int bvm_s[];
  const int f = blockIdx.x, tid = threadIdx.x;
  const size_t fb = (size_t)f * sp.S;
  for (int i = tid; i < MGRID * MGRID; i += 512) bvm_s[i] = 0;
  __syncthreads();
  for (int slot = tid; slot < sp.S; slot += 512) {
    if (label[fb + slot] == 0) continue;                                     // :214
    const float4 p = rec[fb + slot];
    const float vx = __fadd_rn(p.x, 100.0f), vy = __fadd_rn(p.y, 100.0f);   // (pi.x + MAX_RANGE) / interval, interval = 1.0f (:211-212)
    const float v = __fadd_rn(p.z, 2.0f);                                    // :218
    // x = round(vx + 0.5) in double, valid 0..200  <=>  -1 < vx < 200 and then x = floor(vx) + 1
    if (!(vx > -1.0f && vx < 200.0f && vy > -1.0f && vy < 200.0f && v > 0.0f)) continue;
    const int cell = (__float2int_rd(vx) + 1) * MGRID + (__float2int_rd(vy) + 1);
    const int m = __float_as_int(v);
    if (*reinterpret_cast<volatile int*>(&bvm_s[cell]) < m) atomicMax(&bvm_s[cell], m);
  }
  __syncthreads();
  float* o = bvm + (size_t)f * MGRID * MGRID;
  for (int i = tid; i < MGRID * MGRID; i += 512) o[i] = __int_as_float(bvm_s[i]);
}

// The max grids are REPLICATED: CTA b scatters into replica b % n_rep of each grid and k_manip_merge takes the cell-wise
// max of the replicas.  A 201 x 201 grid is only 1263 lines of 128 bytes and the L2 serialises atomics that hit one line, so
// 2 M points into ONE copy of the grid ran at the L2's same-line atomic rate (ncu / bench_extra: 133 us per call whatever
// the kernel did around the atomics); sixteen copies spread the same atomics over sixteen times as many lines.
constexpr int MANIP_MAX_REP = 16;
__global__ void __launch_bounds__(256) k_cloud_manip(int64_t n, Xform xf, const float* __restrict__ x, const float* __restrict__ y,
                                                      const float* __restrict__ z, float* __restrict__ tx, float* __restrict__ ty,
                                                      float* __restrict__ tz, int* __restrict__ rep_in, int* __restrict__ rep_out, int n_rep) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  const int64_t n_pad = (n + 31) & ~(int64_t)31;
  const size_t roff = (size_t)(blockIdx.x % n_rep) * (MGRID * MGRID);
  int* const bev_in = rep_in ? rep_in + roff : nullptr;
  int* const bev_out = rep_out ? rep_out + roff : nullptr;
  // the next point's loads are in flight while this one goes through the check-first reads (the loop is a chain of
  // dependent memory round trips: point -> probe reads -> atomics; the probes of the two grids are issued together)
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  float nx = 0, ny = 0, nz = 0;
  if (i < n) { nx = __ldcs(x + i); ny = __ldcs(y + i); nz = __ldcs(z + i); }
  for (; i < n_pad; i += stride) {
    const bool valid = i < n;
    const float px = nx, py = ny, pz = nz;
    const int64_t j = i + stride;
    if (j < n) { nx = __ldcs(x + j); ny = __ldcs(y + j); nz = __ldcs(z + j); }
    const float ox = __fadd_rn(__fmul_rn(px, xf.m[0]), __fadd_rn(__fmul_rn(py, xf.m[1]), __fadd_rn(__fmul_rn(pz, xf.m[2]), xf.m[3])));
    const float oy = __fadd_rn(__fmul_rn(px, xf.m[4]), __fadd_rn(__fmul_rn(py, xf.m[5]), __fadd_rn(__fmul_rn(pz, xf.m[6]), xf.m[7])));
    const float oz = __fadd_rn(__fmul_rn(px, xf.m[8]), __fadd_rn(__fmul_rn(py, xf.m[9]), __fadd_rn(__fmul_rn(pz, xf.m[10]), xf.m[11])));
    if (valid && tx) { __stcs(tx + i, ox); __stcs(ty + i, oy); __stcs(tz + i, oz); }
    ManipProbe a, b;
    if (bev_in) a = manip_probe(px, py, pz, valid, bev_in);
    if (bev_out) b = manip_probe(ox, oy, oz, valid, bev_out);
    if (bev_in) manip_commit(a, bev_in);
    if (bev_out) manip_commit(b, bev_out);
  }
}

// cell-wise max of the replicas (non-negative float patterns order like ints) -> the two 201 x 201 float grids.
__global__ void __launch_bounds__(256) k_manip_merge(const int* __restrict__ rep_in, const int* __restrict__ rep_out, int n_rep,
                                                      float* __restrict__ bev_in, float* __restrict__ bev_out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= MGRID * MGRID) return;
  const int* rep = blockIdx.y == 0 ? rep_in : rep_out;
  float* out = blockIdx.y == 0 ? bev_in : bev_out;
  if (!rep || !out) return;
  int m = 0;
  for (int r = 0; r < n_rep; r++) m = max(m, rep[(size_t)r * (MGRID * MGRID) + i]);
  out[i] = __int_as_float(m);
}

// ------------------------------------------------------------------------------------------------------------
// extractTopAndFlatten — TopPartRegistration.cpp:79-141 (same code in BatchTopPartRegistration.cpp:90):
// non-ground points are binned into a 10 x 10 grid of 20 m cells (grid = round((p + 100.0f) / 20.0f), valid 0..9);
// every cell with at least 20 points keeps its round(0.2f * count) highest points; the output lists them cell after
// cell (x major), highest first, flattened to z = 0.  std::sort's order among equal heights is unspecified; here equal
// heights keep their input order.
// A segmented selection = one stable LSD radix sort of 40-bit keys (cell : 8 | ~orderable(z) : 32), values = point
// indices in input order, 5 passes of 8 bits; each pass = per-tile digit histogram, scan of the [digit][tile] table,
// stable scatter (rank inside a tile = rounds of 256 elements in thread order: match_any inside the warp + per-warp
// digit counts).  Then per cell: bounds by binary search, quota, prefix over the 100 cells, gather.
// ------------------------------------------------------------------------------------------------------------
constexpr int TOP_GRID = 10, TOP_CELLS = TOP_GRID * TOP_GRID, TOP_MIN_PTS = 20;
constexpr int RS_T = 256, RS_ITEMS = 8, RS_TILE = RS_T * RS_ITEMS;   // 2048 keys per tile
constexpr uint64_t TOP_INVALID = 0xFFull << 32;                       // cell 255: sorts after every real cell

__device__ __forceinline__ int top_axis(float p) {                    // round((p + 100.0f) / 20.0f) -> int (:104-105)
  const float g = roundf(__fdiv_rn(__fadd_rn(p, 100.0f), 20.0f));
  return (g > -2147483904.0f && g < 2147483648.0f) ? __float2int_rz(g) : INT32_MIN;
}

__global__ void __launch_bounds__(256) k_top_keys(int n, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                                   const int16_t* __restrict__ label, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  uint64_t k = TOP_INVALID;
  if (label[i] != 0) {                                                // :99-101
    const int gx = top_axis(x[i]), gy = top_axis(y[i]);
    if (gx >= 0 && gx < TOP_GRID && gy >= 0 && gy < TOP_GRID) {       // :107-109
      float zz = z[i];
      if (zz == 0.0f) zz = 0.0f;                                      // -0 and +0 compare equal in the reference's comparator
      uint32_t u = __float_as_uint(zz);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);                 // order-preserving: larger float -> larger unsigned
      k = ((uint64_t)(gx * TOP_GRID + gy) << 32) | (uint32_t)~u;      // descending height = ascending key
    }
  }
  keys[i] = k; vals[i] = (uint32_t)i;
}

// digit histogram of every tile: hist[digit * n_tiles + tile].  grid n_tiles, block RS_T.
__global__ void __launch_bounds__(RS_T) k_rs_hist(int n, int shift, const uint64_t* __restrict__ keys, uint32_t* __restrict__ hist, int n_tiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_TILE;
  for (int j = 0; j < RS_ITEMS; j++) {
    const int i = base + j * RS_T + threadIdx.x;
    if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of the whole table in place (digit-major, tile-minor = the order of a stable sort).  grid 1, block 1024.
__global__ void __launch_bounds__(1024) k_rs_scan(int m, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b = 0; b < m; b += 1024) {
    const int i = b + tid;
    const uint32_t v = i < m ? hist[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    for (int w = 0; w < 32; w++) { const uint32_t c = s_warp[w]; if (w < wid) wbase += c; total += c; }
    const uint32_t carry = s_carry;
    if (i < m) hist[i] = carry + wbase + incl - v;
    __syncthreads();
    if (tid == 0) s_carry = carry + total;
    __syncthreads();
  }
}

// stable scatter of one tile.  grid n_tiles, block RS_T.
__global__ void __launch_bounds__(RS_T) k_rs_scatter(int n, int shift, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                      const uint32_t* __restrict__ hist, int n_tiles, uint64_t* __restrict__ keys_out,
                                                      uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t s_run[256];              // elements of each digit placed by earlier rounds of this tile
  __shared__ uint16_t s_wc[RS_T / 32][256];    // this round: per-warp count of each digit
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  s_run[tid] = hist[(size_t)tid * n_tiles + blockIdx.x];               // global position of the tile's first element of digit `tid`
  const int base = blockIdx.x * RS_TILE;
  for (int j = 0; j < RS_ITEMS; j++) {
    for (int q = tid; q < (RS_T / 32) * 256; q += RS_T) (&s_wc[0][0])[q] = 0;
    __syncthreads();
    const int i = base + j * RS_T + tid;
    const bool on = i < n;
    uint64_t k = 0; uint32_t v = 0; unsigned d = 256u + lane;          // inactive lanes: private pseudo-digits
    if (on) { k = keys[i]; v = vals[i]; d = (unsigned)(k >> shift) & 255u; }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const unsigned below = __popc(peers & ((1u << lane) - 1u));
    if (on && below == 0) s_wc[wid][d] = (uint16_t)__popc(peers);     // the group's lowest lane publishes the warp's count
    __syncthreads();
    if (on) {
      uint32_t pos = s_run[d] + below;
      for (int w = 0; w < wid; w++) pos += s_wc[w][d];
      keys_out[pos] = k; vals_out[pos] = v;
    }
    __syncthreads();
    {                                                                  // digit `tid`: advance by this round's total
      uint32_t t = 0;
#pragma unroll
      for (int w = 0; w < RS_T / 32; w++) t += s_wc[w][tid];
      s_run[tid] += t;
    }
    __syncthreads();
  }
}

// per cell: [start, end) in the sorted keys, quota = round(0.2f * count) (0 below 20 points), output offsets.  grid 1, block 128.
__global__ void __launch_bounds__(128) k_top_cells(int n, const uint64_t* __restrict__ keys, int* __restrict__ cell_start,
                                                    int* __restrict__ cell_quota, int* __restrict__ cell_out, int* __restrict__ n_out) {
  __shared__ int s_q[TOP_CELLS];
  const int c = threadIdx.x;
  if (c < TOP_CELLS) {
    auto lower = [&](uint64_t key) { int lo = 0, hi = n; while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; } return lo; };
    const int st = lower((uint64_t)c << 32), en = lower((uint64_t)(c + 1) << 32);
    const int cnt = en - st;
    int need = __float2int_rz(roundf(__fmul_rn(0.2f, (float)cnt)));     // :123 (computed before the size test)
    if (cnt < TOP_MIN_PTS) need = 0;                                     // :124-126
    cell_start[c] = st; cell_quota[c] = need; s_q[c] = need;
  }
  __syncthreads();
  if (c == 0) {
    int run = 0;
    for (int k = 0; k < TOP_CELLS; k++) { cell_out[k] = run; run += s_q[k]; }   // cells in (grid_x, grid_y) order (:116-117)
    *n_out = run;
  }
}

__global__ void __launch_bounds__(256) k_top_gather(int n, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                     const int* __restrict__ cell_start, const int* __restrict__ cell_quota,
                                                     const int* __restrict__ cell_out, const float* __restrict__ x, const float* __restrict__ y,
                                                     float* __restrict__ ox, float* __restrict__ oy, uint32_t* __restrict__ oidx) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n) return;
  const unsigned c = (unsigned)(keys[p] >> 32);
  if (c >= (unsigned)TOP_CELLS) return;
  const int j = p - cell_start[c];
  if (j >= cell_quota[c]) return;
  const uint32_t i = vals[p];
  const int o = cell_out[c] + j;
  ox[o] = x[i]; oy[o] = y[i]; oidx[o] = i;                               // flat_point.z = 0 (:135)
}

}  // namespace bevgen

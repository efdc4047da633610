// microbench.cu — design-calibration probes for the ordering stage (not part of the product).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int S = 133312;

__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}
// scattered 16 B stores: idx[f][i] is a permutation of [0,S)
__global__ void k_scatter16(const uint32_t* __restrict__ idx, float4* __restrict__ rec, int F) {
  int f = blockIdx.y; int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) { uint32_t s = idx[(size_t)f * S + i]; rec[(size_t)f * S + s] = make_float4(i, f, s, 1.f); }
}
template <int MODE>
__global__ void k_scatter16v(const uint32_t* __restrict__ idx, float4* __restrict__ rec, int F) {
  int f = blockIdx.y; int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) {
    uint32_t s = idx[(size_t)f * S + i]; float4 v = make_float4(i, f, s, 1.f); float4* p = &rec[(size_t)f * S + s];
    if (MODE == 0) __stcg(p, v); else if (MODE == 1) __stcs(p, v); else if (MODE == 2) __stwt(p, v);
    else if (MODE == 3) { float2* q = reinterpret_cast<float2*>(p); *q = make_float2(v.x, v.y); }              // 8 B
    else if (MODE == 4) { float4* q = &rec[((size_t)f * S + s) * 2]; q[0] = v; q[1] = v; }                      // 32 B per slot
  }
}
// sorted-within-warp scatter: each warp sorts its 32 slots (bitonic via shuffles) before storing
__global__ void k_scatter16_blocksort(const uint32_t* __restrict__ idx, float4* __restrict__ rec, int F) {
  __shared__ uint32_t keys[1024];
  int f = blockIdx.y; int i = blockIdx.x * 1024 + threadIdx.x;
  uint32_t s = i < S ? idx[(size_t)f * S + i] : 0xFFFFFFFFu;
  keys[threadIdx.x] = s; __syncthreads();
  // rank by counting (O(n^2/1024) per thread = 1024 compares) - only to see whether store ORDER matters
  uint32_t r = 0; for (int j = 0; j < 1024; j++) r += keys[j] < s;
  __syncthreads(); if (s != 0xFFFFFFFFu) keys[r] = s; __syncthreads();
  uint32_t t = keys[threadIdx.x];
  if (i < S && t != 0xFFFFFFFFu) rec[(size_t)f * S + t] = make_float4(i, f, t, 1.f);
}
__global__ void k_scatter4(const uint32_t* __restrict__ idx, uint32_t* __restrict__ own, int F) {
  int f = blockIdx.y; int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) { uint32_t s = idx[(size_t)f * S + i]; own[(size_t)f * S + s] = i + 1; }
}
__global__ void k_gather16(const uint32_t* __restrict__ idx, const float4* __restrict__ rec, float4* __restrict__ out, int F) {
  int f = blockIdx.y; int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) { uint32_t s = idx[(size_t)f * S + i]; out[(size_t)f * S + i] = rec[(size_t)f * S + s]; }
}
// smem atomics: one CTA per frame, reads idx (coalesced), does op on smem table
template <int MODE>
__global__ void __launch_bounds__(1024) k_smem(const uint32_t* __restrict__ idx, uint32_t* __restrict__ out) {
  extern __shared__ uint32_t tab[];
  const int f = blockIdx.x;
  const int nwords = MODE == 0 ? (S + 31) / 32 : S / 4;   // bit table or quarter u32 table
  for (int i = threadIdx.x; i < nwords; i += 1024) tab[i] = 0;
  __syncthreads();
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < S; i += 1024) {
    uint32_t s = idx[(size_t)f * S + i];
    if (MODE == 0) { uint32_t old = atomicOr(&tab[s >> 5], 1u << (s & 31)); acc += old & 1; }
    else if (MODE == 1) { atomicMax(&tab[s % (S / 4)], (uint32_t)i + 1); }
    else if (MODE == 2) { tab[s % (S / 4)] = i + 1; }
    else if (MODE == 3) { uint32_t old = atomicMax(&tab[s % (S / 4)], (uint32_t)i + 1); acc += old; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nwords; i += 1024) acc += tab[i];
  if (acc == 0x12345678) out[f] = acc;
}
// cluster probe: 8 CTAs x 1024 threads x smem; cluster.sync cost, DSMEM read bw, remote atomics
__global__ void __launch_bounds__(1024) k_cluster(uint32_t* __restrict__ out, int iters, int mode) {
  extern __shared__ uint32_t sm[];
  cg::cluster_group cl = cg::this_cluster();
  const int nw = 32768;  // 128 KB
  for (int i = threadIdx.x; i < nw; i += 1024) sm[i] = i ^ blockIdx.x;
  cl.sync();
  const unsigned r = cl.block_rank(), n = cl.num_blocks();
  uint32_t acc = 0;
  long long t0 = clock64();
  if (mode == 0) { for (int it = 0; it < iters; it++) cl.sync(); }
  else if (mode == 1) {   // coalesced remote reads
    for (int it = 0; it < iters; it++) {
      const uint32_t* rem = cl.map_shared_rank(sm, (r + 1 + it % (n - 1)) % n);
      for (int i = threadIdx.x; i < nw; i += 1024) acc += rem[i];
    }
  } else if (mode == 2) { // random remote atomicOr
    uint32_t x = threadIdx.x * 2654435761u + r;
    for (int it = 0; it < iters; it++) {
      x = x * 1664525u + 1013904223u;
      uint32_t* rem = cl.map_shared_rank(sm, (x >> 28) % n);
      atomicOr(&rem[(x >> 8) % nw], 1u << (x & 31));
    }
  } else if (mode == 3) { // random remote 16B stores
    uint32_t x = threadIdx.x * 2654435761u + r;
    for (int it = 0; it < iters; it++) {
      x = x * 1664525u + 1013904223u;
      uint4* rem = reinterpret_cast<uint4*>(cl.map_shared_rank(sm, (x >> 28) % n));
      rem[(x >> 8) % (nw / 4)] = make_uint4(x, x, x, x);
    }
  }
  cl.sync();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (uint32_t)(t1 - t0);
  if (acc == 0x12345678) out[1000] = acc;
}

int main() {
  const int F = 1024;
  uint32_t* idx; float4 *rec, *rec2; uint32_t* own; uint32_t* out;
  CK(cudaMalloc(&idx, (size_t)F * S * 4)); CK(cudaMalloc(&rec, (size_t)F * S * 16)); CK(cudaMalloc(&rec2, (size_t)F * S * 16));
  CK(cudaMalloc(&own, (size_t)F * S * 4)); CK(cudaMalloc(&out, 1 << 20));
  {
    std::vector<uint32_t> h((size_t)F * S); std::mt19937 rng(1);
    std::vector<uint32_t> p(S); for (int i = 0; i < S; i++) p[i] = i;
    for (int f = 0; f < 16; f++) { std::shuffle(p.begin(), p.end(), rng); std::copy(p.begin(), p.end(), h.begin() + (size_t)f * S); }
    for (int f = 16; f < F; f++) std::copy(h.begin() + (size_t)(f % 16) * S, h.begin() + (size_t)(f % 16 + 1) * S, h.begin() + (size_t)f * S);
    CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, double bytes, auto fn) {
    for (int i = 0; i < 2; i++) fn();
    cudaEventRecord(e0); for (int i = 0; i < 5; i++) fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    cudaError_t e = cudaGetLastError();
    printf("%-28s %8.3f ms  %7.3f us/frame  %8.1f GB/s  %s\n", name, ms, ms * 1e3 / F, bytes / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  size_t n4 = (size_t)F * S;
  timeit("copy float4", 2.0 * n4 * 16, [&] { k_copy<<<148 * 8, 1024>>>(rec, rec2, n4); });
  dim3 g((S + 255) / 256, F);
  timeit("scatter16 (idx4 + st16)", n4 * 20.0, [&] { k_scatter16<<<g, 256>>>(idx, rec, F); });
  timeit("scatter16 stcg", n4 * 20.0, [&] { k_scatter16v<0><<<g, 256>>>(idx, rec, F); });
  timeit("scatter16 stcs", n4 * 20.0, [&] { k_scatter16v<1><<<g, 256>>>(idx, rec, F); });
  timeit("scatter16 stwt", n4 * 20.0, [&] { k_scatter16v<2><<<g, 256>>>(idx, rec, F); });
  timeit("scatter8", n4 * 12.0, [&] { k_scatter16v<3><<<g, 256>>>(idx, rec, F); });
  timeit("scatter32 (half the frames)", n4 * 0.5 * 36.0, [&] { k_scatter16v<4><<<dim3(g.x, F / 2), 256>>>(idx, rec, F / 2); });
  timeit("scatter16 sorted per 1024", n4 * 20.0, [&] { k_scatter16_blocksort<<<dim3((S + 1023) / 1024, F), 1024>>>(idx, rec, F); });
  timeit("scatter16 256 frames only", n4 * 0.25 * 20.0, [&] { k_scatter16<<<dim3(g.x, F / 4), 256>>>(idx, rec, F / 4); });
  timeit("scatter4  (idx4 + st4)", n4 * 8.0, [&] { k_scatter4<<<g, 256>>>(idx, own, F); });
  timeit("gather16  (idx4+ld16+st16)", n4 * 36.0, [&] { k_gather16<<<g, 256>>>(idx, rec, rec2, F); });
  CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000));
  CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000));
  CK(cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000));
  CK(cudaFuncSetAttribute(k_smem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000));
  timeit("smem atomicOr bits (ret)", n4 * 4.0, [&] { k_smem<0><<<F, 1024, 17000>>>(idx, out); });
  timeit("smem atomicMax u32 (noret)", n4 * 4.0, [&] { k_smem<1><<<F, 1024, 134000>>>(idx, out); });
  timeit("smem plain st u32", n4 * 4.0, [&] { k_smem<2><<<F, 1024, 134000>>>(idx, out); });
  timeit("smem atomicMax u32 (ret)", n4 * 4.0, [&] { k_smem<3><<<F, 1024, 134000>>>(idx, out); });

  // cluster probes
  for (int csz : {0}) { if (csz == 0) break;
    cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
    int smem = 200 * 1024;
    cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (csz > 8) cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cfg.gridDim = dim3(csz * (128 / csz)); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem;
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0; cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k_cluster, &cfg);
    printf("cluster size %2d: max active clusters %d (%s)\n", csz, ncl, cudaGetErrorString(e));
    cudaGetLastError();
    if (ncl == 0) continue;
    for (int mode = 0; mode < 4; mode++) {
      int iters = mode == 0 ? 100 : mode == 1 ? 20 : 200;
      e = cudaLaunchKernelEx(&cfg, k_cluster, out, iters, mode);
      cudaError_t e2 = cudaDeviceSynchronize();
      uint32_t cyc = 0; cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
      const char* nm[] = {"cluster.sync", "remote read 128KB", "remote atomicOr x1024thr", "remote st16 x1024thr"};
      double per = (double)cyc / iters;
      printf("  csz %2d %-26s %10.1f cyc/iter  (%s %s)", csz, nm[mode], per, cudaGetErrorString(e), cudaGetErrorString(e2));
      if (mode == 1) printf("  %.1f B/cyc/SM", 131072.0 / per);
      if (mode >= 2) printf("  %.2f cyc/warp-op", per / 32.0);
      printf("\n");
    }
  }
  return 0;
}

// nanoflann_shim.cpp — builds oracle/_ref/libnanoflann_ref.so from the reference's OWN vendored headers
// (/root/reference/include/nanoflann.hpp + KDTreeVectorOfVectorsAdaptor.h), compiled where they lie.
// TEST INFRASTRUCTURE ONLY (see oracle/bevgen_oracle.c header).  Nothing from the reference is copied here:
// this file only instantiates the same template the reference instantiates (BatchMultiBevGen.cpp:21-22) and
// drives it with the same arguments as the reference call sites (BatchMultiBevGen.cpp:534-550, 593-613), so
// that the oracle's exhaustive-scan k-NN (and its tie rule) can be validated against the real KD-tree.
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include "nanoflann.hpp"
#include "KDTreeVectorOfVectorsAdaptor.h"

using PosVecMat = std::vector<std::vector<float>>;                    // BatchMultiBevGen.cpp:21
using InvKeyTree = KDTreeVectorOfVectorsAdaptor<PosVecMat, float>;    // BatchMultiBevGen.cpp:22

static PosVecMat to_mat(const float* pts, int M) {
  PosVecMat m; m.reserve(M);
  for (int j = 0; j < M; j++) m.emplace_back(std::vector<float>{pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]});
  return m;
}

extern "C" {

// One k-NN query against a freshly built tree (dim 3, leaf 10, SearchParams(10)) — the reference's pattern.
__attribute__((visibility("default")))
int ref_knn(const float* pts, int M, const float* q, int k, uint64_t* idx, float* dist) {
  PosVecMat mat = to_mat(pts, M);
  std::unique_ptr<InvKeyTree> tree = std::make_unique<InvKeyTree>(3, mat, 10);
  std::vector<size_t> ci(k); std::vector<float> cd(k);
  nanoflann::KNNResultSet<float> rs(k);
  rs.init(&ci[0], &cd[0]);
  std::vector<float> qq{q[0], q[1], q[2]};
  tree->index->findNeighbors(rs, qq.data(), nanoflann::SearchParams(10));
  for (int i = 0; i < k; i++) { idx[i] = ci[i]; dist[i] = cd[i]; }
  return (int)rs.size();
}

// Many queries against ONE tree (getKeyFrameLabel's pattern, :593-613).
__attribute__((visibility("default")))
void ref_knn_many(const float* pts, int M, const float* qs, int Q, int k, uint64_t* idx, float* dist) {
  PosVecMat mat = to_mat(pts, M);
  std::unique_ptr<InvKeyTree> tree = std::make_unique<InvKeyTree>(3, mat, 10);
  for (int t = 0; t < Q; t++) {
    std::vector<size_t> ci(k); std::vector<float> cd(k);   // value-initialised, as at :604-605
    nanoflann::KNNResultSet<float> rs(k);
    rs.init(&ci[0], &cd[0]);
    std::vector<float> qq{qs[3 * t], qs[3 * t + 1], qs[3 * t + 2]};
    tree->index->findNeighbors(rs, qq.data(), nanoflann::SearchParams(10));
    for (int i = 0; i < k; i++) { idx[(size_t)t * k + i] = ci[i]; dist[(size_t)t * k + i] = cd[i]; }
  }
}

}  // extern "C"

// Internal C++ interface of the K1 (nearest-neighbour) kernels; see knn.cu.
#pragma once
#include "common.cuh"
namespace prgpu {

constexpr int KNN_TILE = 256;      // nodes per shared-memory tile; SoA stride is a multiple of this
constexpr int KNN_THREADS = 128;   // (virtual) queries per block

struct KnnScratch {                // per-handle device scratch for partial results
  double *part_dist = nullptr;
  uint32_t *part_idx = nullptr;
  size_t part_capacity = 0;        // entries
  double *part2_dist = nullptr;    // second level of a multi-pass merge
  uint32_t *part2_idx = nullptr;
  size_t part2_capacity = 0;
};

// SoA layout: coordinate d of node i at pts[d * stride + i]; stride % KNN_TILE == 0 and the
// buffer is allocated (not necessarily initialised) up to `stride` per dimension.
int knn_search(int dim, const double *pts_soa, uint64_t stride, uint32_t n, uint32_t index_base,
               const double *queries_dev, uint32_t nq, int k, const uint32_t *lo_dev, uint32_t *idx_dev,
               double *dist_dev, KnnScratch &scratch, int sm_count, cudaStream_t stream,
               KernelTimer *timer = nullptr);
int knn_transpose_append(int dim, const double *aos_dev, uint64_t n_new, double *pts_soa, uint64_t stride,
                         uint64_t offset, cudaStream_t stream);
int knn_gather_points(int dim, const double *pts_soa, uint64_t stride, uint64_t first, uint64_t n,
                      double *aos_dev, cudaStream_t stream);
int topk_merge(const uint32_t *idx_parts, const double *dist_parts, int nparts, uint32_t nq, int k,
               uint32_t *idx_out, double *dist_out, cudaStream_t stream);
} // namespace prgpu

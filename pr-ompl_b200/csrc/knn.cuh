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
  // filter path
  double *cand_dist = nullptr;     // merged candidates: nq x (k+margin)
  uint32_t *cand_idx = nullptr;
  size_t cand_capacity = 0;
  uint32_t *flags = nullptr;       // [0] = number of unproven queries, [1+i] = their ids
  size_t flags_capacity = 0;
  double *fb_q = nullptr, *fb_dist = nullptr;  // fallback staging
  uint32_t *fb_idx = nullptr, *fb_lo = nullptr;
  size_t fb_capacity = 0;
  unsigned long long filter_calls = 0, fallback_queries = 0;
};

// Device view of a tree's configuration buffer.
//  pts  : SoA fp64, coordinate d of node i at pts[d * stride + i]; stride % KNN_TILE == 0 and the buffer
//         is allocated (not necessarily initialised) up to `stride` per dimension.  The reference values.
//  ptsf : fp32 mirror for the filter pass, 8 floats per node: x_0..x_6 rounded to fp32 (0 beyond dim) and
//         |x|^2 in slot 7.  NULL when dim == 8 (no free slot): such trees always take the exact path.
//         The same allocation continues with the TF32 mirror of the tensor-core scan (ptsf + 8 * stride:
//         8 floats per node, x_0..x_6 rounded to TF32, 1.0 in slot 7) and the fp32 |x|^2 array
//         (ptsf + 16 * stride: 4 floats per node pair, (n2a, n2b, n2a, n2b)).
//  xmax2: device scalar, max |x|^2 over the nodes (bounds the filter's rounding error).
struct KnnTreeView {
  const double *pts;
  const float *ptsf;
  const double *xmax2;
  uint64_t stride;
  uint32_t n;
  uint32_t index_base;
};

constexpr int KNN_FILTER_SAMPLE_RANK = 8; // the threshold is the 8-th smallest key of a 1/64 sample of the tree
constexpr int KNN_FILTER_MAX_K = 16;   // largest k the filter path serves

// exact search (always bit-identical to the reference); knn_search picks the filter path
// (fp32 scan -> fp64 re-rank of k+8 candidates -> proof check -> exact fallback for unproven
// queries) when the shape allows it and returns the same results.
int knn_search_exact(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, int k,
                     const uint32_t *lo_dev, uint32_t *idx_dev, double *dist_dev, KnnScratch &scratch, int sm_count,
                     cudaStream_t stream, KernelTimer *timer = nullptr);
int knn_search(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, int k, const uint32_t *lo_dev,
               uint32_t *idx_dev, double *dist_dev, KnnScratch &scratch, int sm_count, cudaStream_t stream,
               KernelTimer *timer = nullptr, int force_exact = 0);
// measurement aid behind prgpu_diag_knn_scan: the production scan kernel over the whole tree with caller-given
// thresholds (rounded up to TF32 numbers in place); every reported node's D, its fp32 re-cut key and the two
// error bounds of the selection
int knn_diag_scan(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, float *thresh_dev, uint32_t cap,
                  uint32_t *counts_dev, uint32_t *cidx_dev, float *ckey_dev, float *key32_dev, double *eps_dev,
                  cudaStream_t stream);
// Deferred form for callers that enqueue more work behind the search (the sharded step): the filter path
// leaves the unproven-query list on the device (scratch.flags[0] = count, [1+i] = query ids) instead of
// synchronising; *pending says whether such a list exists.  knn_search_finish(nflag) then re-runs those
// queries with the exact kernel (what knn_search does by itself after its own read-back).
int knn_search_deferred(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, int k,
                        const uint32_t *lo_dev, uint32_t *idx_dev, double *dist_dev, KnnScratch &scratch, int sm_count,
                        cudaStream_t stream, KernelTimer *timer, int force_exact, bool *pending);
int knn_search_finish(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, int k,
                      const uint32_t *lo_dev, uint32_t *idx_dev, double *dist_dev, KnnScratch &scratch, int sm_count,
                      cudaStream_t stream, uint32_t nflag);
int topk_merge_strided(const uint32_t *idx_parts, const double *dist_parts, size_t idx_part_stride,
                       size_t dist_part_stride, int nparts, uint32_t nq, int k, uint32_t *idx_out, double *dist_out,
                       cudaStream_t stream);
int merge_passes(KnnScratch &sc, int nparts, uint32_t nq, int k, uint32_t *idx_out, double *dist_out,
                 cudaStream_t stream);
int scratch_reserve(double *&d, uint32_t *&i, size_t &cap, size_t need);
int knn_transpose_append(int dim, const double *aos_dev, uint64_t n_new, double *pts_soa, float *ptsf,
                         double *xmax2, uint64_t stride, uint64_t offset, cudaStream_t stream);
int knn_gather_points(int dim, const double *pts_soa, uint64_t stride, uint64_t first, uint64_t n,
                      double *aos_dev, cudaStream_t stream);
int topk_merge(const uint32_t *idx_parts, const double *dist_parts, int nparts, uint32_t nq, int k,
               uint32_t *idx_out, double *dist_out, cudaStream_t stream);
} // namespace prgpu

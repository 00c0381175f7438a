// Internal C++ interface of the k-d index (K1, sub-linear form); see kdtree.cu.
#pragma once
#include "common.cuh"
namespace prgpu {
struct __align__(32) KdNode {
  uint32_t dim;   // split dimension
  float cut;      // fp32 cut value: the left half sorts below or at it, the right half at or above it
  float lo, hi;   // bounds of this node's cell along `dim`
  float cut_lo;   // every node of the right half lies at or above this (one fp32 below the cut)
  float cut_hi;   // every node of the left half lies at or below this (one fp32 above the cut)
  float pad[2];
};
struct KdIndex {
  KdNode *nodes = nullptr;      // 2^levels - 1 internal nodes, heap order (children of i: 2i+1, 2i+2)
  double *leaf_pts = nullptr;   // [leaf][dim][32 * ppl] fp64, +inf in unused slots
  float *leaf_f32 = nullptr;    // [leaf][dim][slots] the same coordinates rounded to fp32: what the searches read first
  uint32_t *leaf_idx = nullptr; // [leaf][32 * ppl] insertion index (local to the handle), PRGPU_INVALID_INDEX if unused
  float *root_cell = nullptr;   // lo[8], hi[8] of the root cell
  unsigned long long *visited = nullptr; // leaves visited by the searches so far (when count_visits)
  uint32_t n = 0;               // nodes covered (0: no index)
  int levels = 0, dim = 0;
  int ppl = 2;                  // leaf capacity = 32 * ppl slots (chosen by kd_build from the node count)
  bool count_visits = true;
  float build_ms = 0.f;
  unsigned builds = 0;
  void release();
};
// Fused search + exchange (multi-GPU, queries sharded): instead of idx_dev / dist_dev the team kernel's last
// worker of a query writes the result row straight into EVERY rank's window -- peer memory mapped over
// NVLink -- at row (rank * nq + q): no collective call, the transfer rides on the search.
constexpr int KD_MAX_PEERS = 16;
struct KdPeerOut {
  char *base[KD_MAX_PEERS];   // every rank's window (this process's mapping), dist table at 0
  unsigned long long idx_off; // byte offset of the index table inside a window
  int nranks, rank;
};
// leaf capacity (32 * ppl slots) and depth kd_build takes for n nodes (n >= 128)
void kd_leaf_capacity(uint32_t n, int *ppl, int *levels);
// builds the index over nodes [0, n) of the fp64 SoA buffer; synchronises `stream`
int kd_build(KdIndex &kd, int dim, const double *pts_soa, uint64_t stride, uint32_t n, cudaStream_t stream);
// exact k nearest of every query under (sqrt-distance, index); idx = local index + index_base
int kd_search(const KdIndex &kd, uint32_t index_base, const double *queries_dev, uint32_t nq, int k, uint32_t *idx_dev,
              double *dist_dev, cudaStream_t stream, KernelTimer *timer, const KdPeerOut *peers = nullptr);
} // namespace prgpu

// Internal interface of the K3 (NNF) kernels; see nnf.cu.
#pragma once
#include "world.cuh"
namespace prgpu {
// tensor-product DAG in "sorted" vertex order (NN-graph vertices ordered by topological level)
struct NnfGraphDev {
  int R;              // reference (super-sampled) path length = TPG rows
  uint32_t V;         // NN-graph vertices (c0, c1, IK nodes, sub-sampled edge vertices)
  uint32_t Lv;        // number of topological levels of the NN graph
  const uint32_t *lvl_start; // Lv+1 : first sorted position of each level
  const uint32_t *level;     // V    : level of sorted position s
  const uint32_t *pred_off;  // V+1  : CSR of NN-graph predecessors (edge-creation order)
  const uint32_t *pred_pos;  // E    : sorted position of the predecessor
  const uint32_t *pred_eid;  // E    : NN edge id
  const uint8_t *blocked;    // E    : 1 = edge found in collision (markEdgeInCollision)
  const double *pnn;         // V x 3: end-effector translation of sorted vertex s
  const double *pref;        // R x 3: translation of reference pose r
  double *B;                 // R x V: bottleneck cost table
  uint32_t start_pos, goal_pos; // sorted positions of c0 / c1
};
int nnf_subsample_launch(const DeviceWorld *w_dev, int dof, uint32_t n_new, uint32_t first_vertex,
                         const uint32_t *src, const uint32_t *dst, const uint32_t *step, int D,
                         double *states, double *pose12, cudaStream_t stream);
int nnf_translations_launch(const double *pose12, const uint32_t *vertex_of_pos, uint32_t n, double *out3,
                            cudaStream_t stream);
int nnf_dp_launch(const NnfGraphDev &g, int sm_count, cudaStream_t stream, KernelTimer *timer);
// walks the backpointers goal -> start; out_path holds (row, sorted position) pairs, out_n the
// count, out_cost the goal's bottleneck cost
int nnf_backtrack_launch(const NnfGraphDev &g, uint32_t *out_path, uint32_t cap_pairs, uint32_t *out_n,
                         double *out_cost, cudaStream_t stream);
} // namespace prgpu

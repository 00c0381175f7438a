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
  const double *ftab;        // V x R node weights f[s*R + r] handed over by the host (user metric), or NULL:
                             // then f = |pref[r] - pnn[s]| is computed on the fly
  double *B;                 // R x V: bottleneck cost table
  uint32_t start_pos, goal_pos; // sorted positions of c0 / c1
};
// ---- banded (column-wise) form of the same dynamic program -----------------------------------
// Only cells whose node weight can matter for a path of cost <= tau are stored: vertex s keeps
// rows [band_lo, band_lo + band_w) (a superset interval of {r : f[r,s] <= tau}); everything outside
// is +inf by construction.  Every cell whose true cost is <= tau is computed exactly, so the goal
// cost and the backtracked path equal the full table's whenever the goal cost is <= tau.
struct NnfVertexMeta {   // one 32-byte record per sorted vertex
  uint64_t col_off;      // first cell of the vertex's column in Bc
  uint32_t band_lo, band_w;
  uint32_t pred_begin, pred_end; // range into the packed edge records
  uint32_t pad0, pad1;
};
struct NnfEdgeRec {      // one 16-byte record per (vertex, predecessor) pair, creation order
  uint64_t col_off;      // predecessor's column
  uint32_t band_lo;
  uint32_t band_w_blocked; // bit 31 = edge knocked out; low bits = predecessor's band width
};
struct NnfBandDev {
  int R;
  uint32_t V, Lv;
  const uint32_t *lvl_start; // Lv+1
  const NnfVertexMeta *meta; // V
  const NnfEdgeRec *edge;    // E (CSR order)
  const uint32_t *pred_pos;  // E (CSR order): sorted position of the predecessor
  const double *pnn, *pref;
  const double *ftab;        // as NnfGraphDev::ftab
  double *Bc;
  // hop[cell]: where the optimal path through this cell LEAVES the cell's vertex -- the predecessor cell on
  // another vertex, found by the fixed order of followBackpointers' restatement ((r-1, s) first, then the
  // predecessors in list order, (r, p) before (r-1, p)) and already composed over the rows the path stays on
  // the vertex.  Walking goal -> start costs one dependent load per NN edge of the path.  NULL: not kept.
  unsigned long long *hop;
  uint32_t start_pos, goal_pos;
};
// hop word: bits 63..28 predecessor cell (offset into Bc), bits 27..1 its sorted position, bit 0: the predecessor
// cell's row is (leaving row - 1) instead of the leaving row
constexpr unsigned long long NNF_HOP_START = ~0ull;       // the start cell (0, c0)
constexpr unsigned long long NNF_HOP_NONE = ~0ull - 1ull; // unreachable cell
constexpr uint32_t NNF_HOP_MAX_V = 1u << 27;

// Device-resident LazySP loop (NNFrechet::solve, frechet/src/NNFrechet.cpp:106-152): one persistent block
// chases the hops of the current optimum, evaluates the path's not-yet-known NN edges (evaluateEdge), does the
// reference's bookkeeping in path order, knocks the first colliding edge out and repairs the bottleneck table
// LOCALLY -- only vertices whose predecessors' columns changed are recomputed, level by level, which is what
// LPAStar::updateVertex / computeShortestPath do with their priority queue (LPAStar.cpp:17-159).
struct NnfLazyOut {
  uint32_t status, iters, n_evaluated, n_collisions, path_len, epoch, repaired_vertices, pad;
  double goal_cost, final_error;
  unsigned long long cycles[5]; // SM cycles spent in: hop chase, per-hop pass, edge evaluation, bookkeeping, repair
};
enum : uint32_t { NNF_LAZY_RUNNING = 0, NNF_LAZY_SOLVED = 1, NNF_LAZY_NO_PATH = 2, NNF_LAZY_CONTINUE = 3,
                  NNF_LAZY_NEED_REBUILD = 4, NNF_LAZY_ERROR = 5,
                  NNF_LAZY_NEED_FLOW = 6 }; // the repair wave outgrew one block: the host re-sweeps from `epoch`'s
                                            // level with the whole GPU (out.pad = that level) and relaunches
struct NnfLazyDev {
  const uint32_t *succ_off, *succ_pos; // CSR of NN-graph successors by sorted position
  const uint32_t *succ_pb;             // per successor entry: first predecessor record of that successor (prefetch)
  const uint32_t *level;               // level of a sorted position
  const uint32_t *vertex_of_pos;
  const uint32_t *pred_eid;            // CSR slot -> NN edge id
  NnfEdgeRec *edge_rw;                 // = NnfBandDev::edge, writable (blocked bit)
  uint32_t *stamp, *worklist;          // V each
  const double *states;                // V x dof by vertex id
  int8_t *known;                       // per NN edge id: -1 unknown, 0 collides, 1 free
  uint8_t *evaluated, *blocked;        // per NN edge id
  uint32_t *path_pos, *path_slot;      // cap_hops each
  unsigned long long *path_cell;       // cap_hops
  uint32_t cap_hops;
  uint32_t *out_path;                  // NN-graph vertex ids of the solution
  double *trace;                       // goal cost of every search, or NULL
  NnfLazyOut *out;
  uint32_t max_iters, epoch;
  uint32_t wave_budget;                // columns one block repairs per knock-out before the whole-GPU sweep takes over
  double tau_ub, check_resolution;
  int band_full, dof;
};
int nnf_lazysp_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const NnfBandDev &g, const NnfLazyDev &z,
                      cudaStream_t stream);
int nnf_rowmin_launch(int R, uint32_t V, const double *pnn, const double *pref, const double *ftab,
                      unsigned long long *rowmin_bits, cudaStream_t stream);
int nnf_band_launch(int R, uint32_t V, const double *pnn, const double *pref, const double *ftab, double tau,
                    uint32_t *band_lo, unsigned long long *band_w64, cudaStream_t stream);
int nnf_band_pack_launch(uint32_t V, uint32_t E, const uint32_t *band_lo, const unsigned long long *band_w64,
                         const unsigned long long *col_off, const uint32_t *pred_off, const uint32_t *pred_pos,
                         const uint32_t *pred_eid, const uint8_t *blocked, NnfVertexMeta *meta, NnfEdgeRec *edge,
                         cudaStream_t stream);
int nnf_band_dp_launch(const NnfBandDev &g, uint32_t first_level, cudaStream_t stream, KernelTimer *timer);
int nnf_flow_dp_launch(const NnfBandDev &g, uint32_t first_pos, uint32_t epoch, uint32_t *done, int sm_count,
                       cudaStream_t stream, KernelTimer *timer);
int nnf_band_backtrack_launch(const NnfBandDev &g, uint32_t *out_path, uint32_t cap_pairs, uint32_t *out_n,
                              double *out_cost, cudaStream_t stream);
int nnf_exclusive_scan_u64(const unsigned long long *in, unsigned long long *out, uint32_t n, void *temp,
                           size_t &temp_bytes, cudaStream_t stream);

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

/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * C ABI of the CPU oracle (liboracle.so) and of the compiled reference
 * (oracle/_ref/libref_rrtconnect.so exports ref_rrtc_solve with the signature of
 * orc_rrtc_solve; oracle/_ref/libref_frechet.so exports ref_nnf_run).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load these libraries.
 * Parity unpinned upstream: the reference ships no tests, fixtures or golden
 * vectors; the oracle is pinned instead against the
 * reference's own RRTConnect.cpp compiled here (oracle/_ref) and against
 * hand-computed known answers (tests/test_oracle_*.py). */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- nearest neighbours (NearestNeighborsLinear / findKnnNodes semantics) ---- */
/* pts: n x dim row-major; queries: nq x dim.  Candidates of query i are the
 * points with index >= lo[i] (lo == NULL: all).  Results ordered by
 * (sqrt-distance, index); idx padded with 0xFFFFFFFF, dist with +inf. */
int orc_knn(const double *pts, uint64_t n, int dim, const double *queries, uint64_t nq, int k,
            const uint32_t *lo, uint32_t *idx_out, double *dist_out, int nthreads);
/* the same answers from a k-d tree: the GNAT-class structure the reference would really get from
 * SelfConfig::getDefaultNearestNeighbors (oracle/kdtree.cpp) -- the honest CPU baseline */
void *orc_kdtree_build(const double *pts, uint64_t n, int dim);
void orc_kdtree_destroy(void *tree);
int orc_kdtree_knn(const void *tree, const double *queries, uint64_t nq, int k, uint32_t *idx_out,
                   double *dist_out, int nthreads);
double orc_distance(const double *a, const double *b, int dim);
void orc_interpolate(const double *from, const double *to, double t, int dim, double *out);

/* ---- synthetic world ---- */
void *orc_world_create(int dof, const double *dh /* dof x 5: a,d,cos(alpha),sin(alpha),theta0 */,
                       int n_link_spheres, const int32_t *sphere_link, const double *sphere_xyzr,
                       int n_obs_spheres, const double *obs_spheres /* x4 */, int n_obs_boxes,
                       const double *obs_boxes /* x6 */);
void orc_world_destroy(void *world);
double orc_clearance(const void *world, const double *q);
void orc_fk(const void *world, const double *q, double *pose12, double *sphere_centres);

/* ---- edge validation ---- */
enum { ORC_MODE_OMPL_DISCRETE = 0, ORC_MODE_NNF_ALPHA = 1 };
/* mode 0: DiscreteMotionValidator::checkMotion(from,to) with longest valid segment
 *         `resolution` (check_first != 0: isValid(from) && checkMotion(from,to)).
 * mode 1: NNFrechet::evaluateEdge with check resolution `resolution`; valid = !inCollision.
 * min_clearance (nullable): minimum clearance over ALL tested states (no early exit).
 * n_tested (nullable): number of states in the tested set. */
int orc_check_motion_batch(const void *world, const double *from, const double *to, uint64_t n,
                           double resolution, int mode, int check_first, uint8_t *valid,
                           double *min_clearance, uint32_t *n_tested, int nthreads);
int orc_is_valid_batch(const void *world, const double *q, uint64_t n, uint8_t *valid,
                       double *clearance, int nthreads);

/* ---- sample stream ---- */
double orc_sample_uniform(uint64_t seed, uint64_t query, uint64_t it, uint64_t dim, double lo, double hi);

/* ---- RRTConnect (port in liboracle.so; ref_rrtc_solve in oracle/_ref has the same signature) ---- */
typedef struct {
  int32_t status;        /* ompl::base::PlannerStatus::StatusType */
  uint32_t iterations;   /* uniform samples drawn */
  uint32_t n_start, n_goal;
  uint32_t path_len;
  uint32_t overflow;     /* 1 if an output buffer was too small */
} orc_rrtc_result;
int orc_rrtc_solve(const void *world, int dim, const double *lo, const double *hi, const double *start,
                   const double *goal, double range, int ext_first, int ext_second,
                   double resolution_fraction, uint64_t seed, uint64_t query_id, uint32_t max_iters,
                   uint32_t max_nodes, uint32_t max_path, orc_rrtc_result *res, double *start_states,
                   int32_t *start_parent, double *goal_states, int32_t *goal_parent, double *path);
/* ---- NNF (flat-array port of frechet/src/{NNFrechet,LPAStar,util}.cpp; the literal reference is
 * compiled unmodified into oracle/_ref/libref_frechet.so over the stand-ins of oracle/compat/ and is
 * driven through ref_nnf_run below) ---- */
/* util.cpp:3-21 linDoubleSpace/linIntSpace: (int)(min + i*delta), delta = (max-min)/(N-1) */
void orc_lin_int_space(double mn, double mx, uint64_t n, int32_t *out);
/* NNFrechet.cpp:398-448: IK budget per super-sampled reference index: counts[0] = counts[R-1] = M,
 * interior = multinomial of W*M - 2M draws of std::uniform_int_distribution<int>(1, R-2) on a
 * std::default_random_engine seeded with `seed` (a fresh distribution object per draw, :324). */
void orc_nnf_layer_counts(int seed, int R, int W, int M, uint32_t *counts);
typedef struct {
  uint32_t n_vertices, n_edges, n_ik;
  uint32_t lazy_iters;      /* LazySP iterations executed (TPG searches) */
  int32_t solved;           /* 1: collision-free path found; 0: no path left / iteration cap */
  double final_error;       /* mFinalError of the last extracted path (NNFrechet.cpp:615-645) */
  double goal_cost;         /* bottleneck cost g[goal] of the last search */
  uint32_t path_len;        /* NN-graph vertices on the final path */
  uint32_t n_evaluated, n_collisions;
  uint32_t overflow;
} orc_nnf_result;
/* ref_pose12: R x 12 (sub-sampled reference path); layer_counts: R; ik_states: n_ik x dof in vertex
 * order (layer 0, layer R-1, then layers 1..R-2).  Task metric = Euclidean distance of the pose
 * translations.  knn_out: n_ik x k neighbour vertex ids (0xFFFFFFFF padded). */
int orc_nnf_run(const void *world, int R, const double *ref_pose12, int k, int D, double check_res,
                const uint32_t *layer_counts, const double *ik_states, uint32_t max_lazy_iters, int nthreads,
                orc_nnf_result *res, uint32_t *knn_out, double *vertex_states, double *vertex_pose12,
                uint32_t cap_vertices, uint32_t *path_vertices, uint32_t cap_path, double *goal_cost_trace);
/* Independent check of a bottleneck cost (NOT the DP recurrence of orc_nnf_run): is the goal cell
 * (R-1, c1) reachable from (0, c0) through tensor-product cells whose task distance is <= tau
 * (strict != 0: < tau), along product edges A/B/C of NNFrechet.cpp:567-596 whose NN edge is not
 * blocked?  pose12: V x 12; edges: E x 2 (source, target) with source before target in some
 * topological order (any NN graph built by buildNNGraph); blocked: E bytes or NULL.
 * A cost c is the min-bottleneck cost iff reachable(c, 0) && !reachable(c, 1). */
int orc_nnf_reachable(int R, const double *ref_pose12, uint32_t V, const double *pose12, uint32_t E,
                      const uint32_t *edges, const uint8_t *blocked, double tau, int strict);

/* ---- the reference's own NNFrechet::NNFrechet (oracle/ref_nnf_driver.cpp -> oracle/_ref/libref_frechet.so;
 * libref_frechet_rev.so = same sources with the stand-in graph's edge-list iteration order reversed) ---- */
typedef struct {
  uint32_t R;               /* size of the sub-sampled reference path (NNFrechet.cpp:235-252) */
  uint32_t n_vertices, n_edges, n_ik; /* NN graph */
  uint64_t n_tpg_vertices, n_tpg_edges;
  int32_t solved;           /* solve() returned EXACT_SOLUTION */
  double final_error;       /* mFinalError (NNFrechet.hpp:282) */
  double goal_cost;         /* LPAStar::mDistanceMap[goal] after the last search */
  uint32_t path_len;
  uint32_t n_evaluated, n_collisions;
  uint32_t overflow;
  double setup_seconds, solve_seconds;
} ref_nnf_result;
/* ref_full_pose12: the caller's reference path BEFORE subsampleRefPath; ik_states: pre-drawn IK solutions
 * handed out in the reference's call order.  search_only != 0: setup() + one computeShortestPath +
 * extractNNPath (first-search cost); otherwise solve(solve_seconds). */
int ref_nnf_run(const void *world, uint32_t n_ref_full, const double *ref_full_pose12, int W, int M, int k, int D,
                int seed, const double *ik_states, uint32_t n_ik_avail, double solve_seconds, int search_only,
                ref_nnf_result *res, double *ref_sub_pose12, uint32_t cap_ref, uint32_t *ik_call_counts,
                uint32_t *knn_out, double *vertex_states, double *vertex_pose12, uint32_t cap_vertices,
                uint32_t *nn_edges, uint8_t *nn_edge_flags, uint32_t cap_edges, double *path_states,
                uint32_t cap_path);
#ifdef __cplusplus
}
#endif
#endif

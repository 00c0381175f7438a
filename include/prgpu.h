/* prgpu.h -- C ABI of the B200-native hot path behind pr-ompl's planners.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its
 * hot path sits behind three OMPL C++ plugin interfaces.  Each entry point below names the
 * reference call site it replaces; the C++ host mirror (pr-ompl_b200/plugin/) subclasses
 * those OMPL interfaces and calls only this header.
 *
 *   ompl::NearestNeighbors<Motion*>  (pr_ompl/include/pr_ompl/RRTConnect.h:154; add/nearest/
 *       size/list/clear used at pr_ompl/src/RRTConnect.cpp:181,209,235,278,172-174)
 *       and findKnnNodes (frechet/src/util.cpp:34-58)                 -> prgpu_knn_*
 *   ompl::base::StateValidityChecker / MotionValidator (si_->checkMotion, isValid at
 *       pr_ompl/src/RRTConnect.cpp:198; evaluateEdge frechet/src/NNFrechet.cpp:655-683)
 *                                                                     -> prgpu_world_*, prgpu_check_motion_*
 *   pr_ompl::RRTConnect::solve / growTree (pr_ompl/src/RRTConnect.cpp:178-376), many
 *       independent queries                                           -> prgpu_rrtc_batch_*
 *   NNFrechet::setup / solve (frechet/src/NNFrechet.cpp:95-221, LPAStar.cpp:125-159)
 *                                                                     -> prgpu_nnf_*
 *
 * Conventions: opaque handles; every function returns 0 on success and a negative
 * PRGPU_E* code on failure (text via prgpu_last_error(), thread-local); no C++ exceptions
 * or STL types cross the boundary; the caller owns every host buffer; a handle owns its
 * device buffers and a CUDA stream; calls are synchronous unless they take a stream
 * argument (`*_device` variants: device pointers in/out, enqueued on `stream`, a
 * cudaStream_t passed as void*); one handle is used from one thread at a time.  There is
 * NO CPU fallback: without a usable CUDA device every compute entry point fails with
 * PRGPU_ENODEV.
 */
#ifndef PRGPU_H
#define PRGPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PRGPU_OK 0
#define PRGPU_EINVAL (-1)   /* bad argument */
#define PRGPU_ENODEV (-2)   /* no CUDA device / wrong architecture */
#define PRGPU_ECUDA (-3)    /* CUDA runtime error (see prgpu_last_error) */
#define PRGPU_ENOMEM (-4)
#define PRGPU_ELIMIT (-5)   /* size beyond a documented limit */

#define PRGPU_INVALID_INDEX 0xFFFFFFFFu
#define PRGPU_MAX_DIM 8
#define PRGPU_MAX_K 64

const char *prgpu_last_error(void);
int prgpu_version(void);
int prgpu_device_count(void); /* number of CUDA devices, 0 if none; never fails */

/* ------------------------------------------------------------------------------------
 * K1: nearest neighbours over a configuration buffer (a planner's tree).
 * Points are appended in order; the insertion index is the point's identity and the
 * tie-break: results are ordered by (sqrt-distance, index), exactly the order of OMPL's
 * linear scan (first strict minimum wins) and of findKnnNodes' sort of (distance, vertex).
 * Distances are RealVectorStateSpace::distance bit for bit (sequential sum of squared
 * differences in double, no FMA, IEEE sqrt).
 * ---------------------------------------------------------------------------------- */
/* One-at-a-time use (a host planner: add(1 node), nearest(1 query) per growTree): calls of <= 64 queries on trees of
 * <= 65536 nodes and adds of <= 8 points pass their arguments and results through a window of mapped pinned host
 * memory instead of copy calls.  A small prgpu_knn_add() returns once its append is QUEUED on the handle's stream:
 * every later call on the handle is ordered behind it (the *_device entry points wait for it when given another
 * stream), and an error of the append surfaces at the next synchronising call.  PRGPU_NO_HOST_WINDOW=1 keeps the
 * copy path for every size. */
typedef struct prgpu_knn prgpu_knn_t;
int prgpu_knn_create(int device, int dim, uint64_t capacity_hint, prgpu_knn_t **out);
int prgpu_knn_destroy(prgpu_knn_t *h);
int prgpu_knn_clear(prgpu_knn_t *h);                                   /* NearestNeighbors::clear */
int prgpu_knn_add(prgpu_knn_t *h, const double *pts, uint64_t n);      /* ::add ; pts n x dim row-major, host */
int prgpu_knn_add_device(prgpu_knn_t *h, const double *pts_dev, uint64_t n, void *stream);
int prgpu_knn_size(const prgpu_knn_t *h, uint64_t *n);                 /* ::size */
/* global index reported for local point i is index_base + i (node-sharded trees) */
int prgpu_knn_set_index_base(prgpu_knn_t *h, uint32_t index_base);
/* copies points [first, first+n) back in insertion order (NearestNeighbors::list) */
int prgpu_knn_get_points(prgpu_knn_t *h, uint64_t first, uint64_t n, double *pts);
/* ::nearest (k=1) / ::nearestK ; findKnnNodes when lo != NULL: candidates of query i are
 * the points with local index >= lo[i].  idx: nq x k (PRGPU_INVALID_INDEX padded),
 * dist: nq x k (+inf padded) or NULL. */
int prgpu_knn_nearest_k(prgpu_knn_t *h, const double *queries, uint64_t nq, int k, const uint32_t *lo,
                        uint32_t *idx, double *dist);
/* (Host form: when idx / dist are pinned host memory and the k-d index serves the call, the search kernel writes the
 * result rows straight into them -- the transfer rides under the search instead of following it as copies.) */
int prgpu_knn_nearest_k_device(prgpu_knn_t *h, const double *queries_dev, uint64_t nq, int k,
                               const uint32_t *lo_dev, uint32_t *idx_dev, double *dist_dev, void *stream);
/* Batches in flight.  Searches served by the k-d index (below) use no per-handle scratch: device-buffer calls on
 * one handle may be issued on several streams at once (different output buffers), and the next batch's blocks then
 * take the SM slots the tail of the previous one leaves idle -- 1.5x the one-at-a-time throughput at 4096 queries
 * over 1e7 nodes (bench.py: extra.two_batches_in_flight).  The brute-force paths share the handle's scratch:
 * one call at a time there. */
/* Search strategy.  By default batched calls (>= 128 queries, k <= 16, dim <= 7, >= 8192 points) use an
 * TF32 tensor-core filter scan followed by an fp32 re-cut and an exact fp64 re-rank of the few surviving
 * candidates, with a per-query proof that no discarded point can precede them; unproven queries are
 * re-run by the exact fp64 kernel, so the results are bit-identical either way.  The filter path
 * synchronises the stream once per call (it reads the number of unproven queries).  force_exact != 0
 * disables the filter. */
int prgpu_knn_set_search_mode(prgpu_knn_t *h, int force_exact);
/* Builds the k-d index over the current nodes now (synchronous, ~4.4 ns per node).  Without this call the index
 * is built automatically once the batched searches on an unchanged node set have cost as much extra as the build
 * (see knn_dispatch in csrc/capi.cu); appends and clear() make it stale.  Results never depend on it. */
int prgpu_knn_build_index(prgpu_knn_t *h);
/* Sub-linear search.  Batched calls without `lo` (>= 16 queries, k <= 32) over >= 2^17 points are served by a
 * balanced k-d tree built on the device over the configuration buffer: same results bit for bit under the
 * (distance, index) order, ~1e-3 of the pairs.  The index is built at the second such call on an unchanged
 * point set (so a structure refilled before every search never builds one); add/clear make it stale.
 * prgpu_knn_set_search_mode(h, 2) keeps the brute-force paths only (mode 1: the exact fp64 kernel only;
 * mode 3: the index whenever it can answer -- no `lo`, k <= 32, >= 128 points -- built at once; mode 0: auto).
 * indexed_nodes: points covered by a valid index (0: none); leaves_visited: leaves (64 or 96 slots) evaluated by
 * all index searches so far (reading it synchronises the device).  Any pointer may be NULL. */
int prgpu_knn_index_stats(prgpu_knn_t *h, uint64_t *indexed_nodes, float *build_ms, uint32_t *builds,
                          uint64_t *index_calls, uint64_t *leaves_visited);
/* number of calls served by the filter path and of queries that needed the exact fallback */
int prgpu_knn_search_stats(prgpu_knn_t *h, uint64_t *filter_calls, uint64_t *fallback_queries);
/* measurement aid: when on, CUDA events bracket the search kernel of every nearest_k call on
 * the stream it is launched on; last_kernel_ms waits for and returns the latest duration. */
int prgpu_knn_set_profiling(prgpu_knn_t *h, int on);
int prgpu_knn_last_kernel_ms(prgpu_knn_t *h, float *ms);
/* merge `nparts` per-shard candidate tables (each nq x k, ordered, same padding) into the
 * global top-k under the same (distance, index) order: the step after the NCCL all-gather
 * of per-shard candidates.  Layout: part p of query q at ((p * nq) + q) * k. */
int prgpu_topk_merge_device(int device, const uint32_t *idx_parts_dev, const double *dist_parts_dev,
                            int nparts, uint64_t nq, int k, uint32_t *idx_dev, double *dist_dev,
                            void *stream);

/* ------------------------------------------------------------------------------------
 * K2: synthetic world (sphere-approximated FK of a DH arm vs sphere/box obstacles) and
 * batched discretised edge validation.
 * ---------------------------------------------------------------------------------- */
typedef struct prgpu_world prgpu_world_t;
#define PRGPU_MAX_LINK_SPHERES 32
#define PRGPU_MAX_OBSTACLES 4096
/* dh: dof x 5 (a, d, cos(alpha), sin(alpha), theta0); sphere_link[s] in [0,dof];
 * sphere_xyzr: S x 4; obs_spheres: x4 (x,y,z,R); obs_boxes: x6 (cx,cy,cz,hx,hy,hz).
 * Creation also builds the broad-phase structures on the device (an obstacle grid and a clearance
 * field of up to 2^24 fp32 cells, a few milliseconds); they only accelerate, every verdict is that of the
 * double-precision specification (clearance < 0 <=> collision), exactly, outside the 1e-6 contact band. */
int prgpu_world_create(int device, int dof, const double *dh, int n_link_spheres,
                       const int32_t *sphere_link, const double *sphere_xyzr, int n_obs_spheres,
                       const double *obs_spheres, int n_obs_boxes, const double *obs_boxes,
                       prgpu_world_t **out);
int prgpu_world_destroy(prgpu_world_t *w);
int prgpu_world_dof(const prgpu_world_t *w);
int prgpu_world_set_profiling(prgpu_world_t *w, int on);      /* brackets the check_motion kernel */
int prgpu_world_last_kernel_ms(prgpu_world_t *w, float *ms);

#define PRGPU_MODE_OMPL_DISCRETE 0 /* DiscreteMotionValidator::checkMotion: {to} U {j/nd, 0<j<nd} */
#define PRGPU_MODE_NNF_ALPHA 1     /* NNFrechet::evaluateEdge: alpha = 0, +1/numQ, ... while <= 1 */
/* valid[i] = 1 iff every state of edge i's tested set is collision free.
 * mode 0: resolution = the space's longest valid segment length; check_first != 0 also
 *         tests `from` (the goal-tree form isValid(from) && checkMotion(from,to),
 *         pr_ompl/src/RRTConnect.cpp:198).
 * mode 1: resolution = NNF's check resolution (0.05, frechet/include/NNFrechet.hpp:277);
 *         valid = !evaluateEdge(...).
 * States that a tested neighbour's clearance proves free are not tested (conservative advancement): the
 * result is still the AND over the whole tested set.  The host-buffer form stages its uploads under the
 * kernels for batches of 2^19 edges and more. */
int prgpu_check_motion_batch(prgpu_world_t *w, const double *from, const double *to, uint64_t n,
                             double resolution, int mode, int check_first, uint8_t *valid);
int prgpu_check_motion_batch_device(prgpu_world_t *w, const double *from_dev, const double *to_dev,
                                    uint64_t n, double resolution, int mode, int check_first,
                                    uint8_t *valid_dev, void *stream);
/* StateValidityChecker::isValid for n states (n x dof) */
int prgpu_is_valid_batch(prgpu_world_t *w, const double *q, uint64_t n, uint8_t *valid);
int prgpu_is_valid_batch_device(prgpu_world_t *w, const double *q_dev, uint64_t n, uint8_t *valid_dev,
                                void *stream);
/* end-effector pose (3x4 row-major [R|p]) of n states: the FK callback of NNF */
int prgpu_fk_batch(prgpu_world_t *w, const double *q, uint64_t n, double *pose12);
int prgpu_fk_batch_device(prgpu_world_t *w, const double *q_dev, uint64_t n, double *pose12_dev,
                          void *stream);

/* ------------------------------------------------------------------------------------
 * T1: many independent RRTConnect queries (pr_ompl::RRTConnect::solve, one planner per query,
 * pr_ompl/src/RRTConnect.cpp:219-376) against one world: one warp per query, then 8-warp teams for the
 * queries still unsolved after 256 samples (same trees, iteration counts and paths as one planner each).
 * Start/goal are single configurations (ProblemDefinition::setStartAndGoalStates); samples come
 * from the counter-based stream u(seed, query_id_base + i, iteration, dim) over [lo, hi];
 * termination = max_iters uniform samples (the ptc) or a full tree.
 * ---------------------------------------------------------------------------------- */
typedef struct prgpu_rrtc_batch prgpu_rrtc_batch_t;
#define PRGPU_EXT_EXTEND 0  /* RRTConnect::EXTTYPE_EXTEND  */
#define PRGPU_EXT_CONNECT 1 /* RRTConnect::EXTTYPE_CONNECT */
/* ompl::base::PlannerStatus::StatusType values reported per query */
#define PRGPU_STATUS_INVALID_START 1
#define PRGPU_STATUS_INVALID_GOAL 2
#define PRGPU_STATUS_TIMEOUT 4
#define PRGPU_STATUS_EXACT_SOLUTION 6
typedef struct {
  int32_t dim;
  double lo[PRGPU_MAX_DIM], hi[PRGPU_MAX_DIM]; /* RealVectorBounds */
  double range;               /* "range" param; < eps -> 0.2 * maximum extent (configurePlannerRange) */
  double resolution_fraction; /* si->setStateValidityCheckingResolution (OMPL default 0.01) */
  int32_t ext_first, ext_second; /* "extension_type": con-con / con-ext / ext-con (default) / ext-ext */
  uint64_t seed, query_id_base;
  uint32_t max_iters; /* uniform samples per query */
  uint32_t max_nodes; /* capacity of each tree */
} prgpu_rrtc_params;
int prgpu_rrtc_batch_create(prgpu_world_t *w, const prgpu_rrtc_params *params, uint32_t n_queries,
                            prgpu_rrtc_batch_t **out);
int prgpu_rrtc_batch_destroy(prgpu_rrtc_batch_t *h);
/* starts/goals: n_queries x dim (host).  Synchronous: upload, plan all queries, keep trees on
 * the device. */
int prgpu_rrtc_batch_run(prgpu_rrtc_batch_t *h, const double *starts, const double *goals);
int prgpu_rrtc_batch_run_device(prgpu_rrtc_batch_t *h, const double *starts_dev, const double *goals_dev,
                                void *stream);
/* per-query results (any pointer may be NULL): status, samples drawn, tree sizes */
int prgpu_rrtc_batch_summary(prgpu_rrtc_batch_t *h, int32_t *status, uint32_t *iterations,
                             uint32_t *n_start, uint32_t *n_goal);
/* one query's trees in insertion order (NearestNeighbors::list / getPlannerData,
 * pr_ompl/src/RRTConnect.cpp:378-415) and its solution path (start ... goal).  Buffers hold
 * max_nodes x dim / max_nodes / max_path x dim entries; path_len receives the path length
 * (0 if unsolved). */
int prgpu_rrtc_batch_fetch(prgpu_rrtc_batch_t *h, uint32_t query, double *start_states,
                           int32_t *start_parent, double *goal_states, int32_t *goal_parent,
                           uint32_t *path_len, double *path, uint32_t max_path);

/* ------------------------------------------------------------------------------------
 * K3: NNF (Nearest-Neighbor Frechet) end-effector path following
 * (NNFrechet::setup / solve, frechet/src/NNFrechet.cpp:184-221, :95-178).
 * The host planner supplies what the reference obtains through its callbacks: the
 * sub-sampled reference poses (subsampleRefPath :235-252) and the IK solutions of every
 * layer (sampleIKNodes :282-308; poses are recomputed with the world's FK as at :295).
 * create() = buildNNGraph + buildTensorProductGraph: suffix-masked kNN (findKnnNodes), the D
 * sub-sampled vertices of every edge with their FK, and the tensor-product DAG in flat
 * arrays.  solve() = the LazySP loop: wavefront min-bottleneck search, path extraction,
 * evaluateEdge on the path's NN edges (accumulated-alpha discretisation at
 * check_resolution), knock-out of the first colliding edge, repeat.
 * Task-space metric: Euclidean distance between pose translations.
 * ---------------------------------------------------------------------------------- */
typedef struct prgpu_nnf prgpu_nnf_t;
typedef struct {
  int32_t num_nn;          /* mNumNN (default 5, frechet/include/NNFrechet.hpp:269) */
  int32_t discretization;  /* mDiscretization D >= 1 (default 3, :274) */
  double check_resolution; /* mCheckResolution (0.05, :277) */
} prgpu_nnf_params;
typedef struct {
  uint32_t n_vertices, n_edges, n_ik;
  uint32_t lazy_iters;  /* TPG searches executed */
  int32_t solved;       /* 1: collision-free path found */
  double final_error;   /* mFinalError: max node weight over non-dummy path nodes (:615-645) */
  double goal_cost;     /* bottleneck cost of the goal in the last search */
  uint32_t path_len;
  uint32_t n_evaluated, n_collisions;
  uint32_t overflow;
} prgpu_nnf_result;
/* ref_pose12: R x 12; layer_counts: R (IK solutions per reference index); ik_states: sum(counts) x dof
 * in NN-graph vertex order: layer 0, layer R-1, then layers 1..R-2 (the order of buildNNGraph). */
int prgpu_nnf_create(prgpu_world_t *w, const prgpu_nnf_params *params, int R, const double *ref_pose12,
                     const uint32_t *layer_counts, const double *ik_states, prgpu_nnf_t **out);
int prgpu_nnf_destroy(prgpu_nnf_t *h);
int prgpu_nnf_graph_info(prgpu_nnf_t *h, uint32_t *n_vertices, uint32_t *n_edges, uint32_t *n_ik);
/* neighbours chosen for every IK node (n_ik x num_nn NN-graph vertex ids, PRGPU_INVALID_INDEX padded) */
int prgpu_nnf_get_knn(prgpu_nnf_t *h, uint32_t *knn);
/* all NN-graph vertices: configuration (n_vertices x dof; zeros for the two dummies) and pose */
int prgpu_nnf_get_vertices(prgpu_nnf_t *h, double *states, double *pose12);
/* how findKnnNodes was executed at create time: calls that took the tensor-core filter path of K1 and
 * queries that had to be re-run by the exact kernel (see prgpu_knn_search_stats) */
int prgpu_nnf_knn_stats(prgpu_nnf_t *h, uint64_t *filter_calls, uint64_t *fallback_queries);
/* NN-graph edges in creation order (n_edges x 2: source, target) and per-edge LazySP state:
 * bit 0 = evaluated (the EProp::evaluated marker, NNFrechet.cpp:658-660), bit 1 = knocked out
 * (markEdgeInCollision, :685-702).  Either pointer may be NULL. */
int prgpu_nnf_get_edges(prgpu_nnf_t *h, uint32_t *edges, uint8_t *flags);
/* path_vertices: NN-graph vertex ids of the solution (cap_path entries); goal_cost_trace: goal
 * cost of every search (max_lazy_iters entries) or NULL.  May be called again to continue. */
int prgpu_nnf_solve(prgpu_nnf_t *h, uint32_t max_lazy_iters, prgpu_nnf_result *res, uint32_t *path_vertices,
                    uint32_t cap_path, double *goal_cost_trace);
/* Host-callback form: the reference takes forward kinematics, the task-space metric and state
 * validity as opaque user functions (frechet/include/NNFrechet.hpp:43-59, :303-312; call sites NNFrechet.cpp:295,
 * :372 FK, :495 metric, :665,:675 isValid).  For a robot or metric the device has no model of, the host planner
 *   1. prgpu_nnf_create_graph(): findKnnNodes for every node + the NN graph with its sub-sampled vertices
 *      (configurations only) on the device;
 *   2. reads the V configurations back (prgpu_nnf_get_vertices, pose12 = NULL), runs ITS FK and metric and hands
 *      over the node-weight table weights[v * R + r] = distance(ref_pose[r], FK(state[v])) by NN-graph vertex id
 *      (rows of the two dummy vertices: the weights of ref_pose[0] / ref_pose[R-1] themselves, NNFrechet.cpp:409,414);
 *   3. installs an edge evaluator -- evaluateEdge (:655-683) over the caller's own validity checker: for n edges
 *      (from, to: n x dof) write free[i] = 1 if edge i is collision-free, 0 if in collision; return 0, or non-zero
 *      to abort the solve with an error;
 *   4. prgpu_nnf_solve() as in the device form: the bottleneck search, path extraction and LazySP bookkeeping are
 *      the same code, only the weights come from the table and the verdicts from the callback. */
typedef int (*prgpu_nnf_edge_eval_fn)(void *user, const double *from, const double *to, uint32_t n, uint8_t *free_out);
int prgpu_nnf_create_graph(int device, int dof, const prgpu_nnf_params *params, int R, const uint32_t *layer_counts,
                           const double *ik_states, prgpu_nnf_t **out);
int prgpu_nnf_set_weights(prgpu_nnf_t *h, const double *weights /* n_vertices x R */);
/* also accepted by a device-form handle: replaces the device's checkMotion batch by the caller's verdicts */
int prgpu_nnf_set_edge_evaluator(prgpu_nnf_t *h, prgpu_nnf_edge_eval_fn fn, void *user);
/* search kernel: 0 (default) = banded column-wise DP (only cells that can lie on a path of cost
 * <= tau are stored; tau is widened until the optimum is provably inside) executed as a dataflow
 * over per-vertex ready flags; 1 = full R x V table swept by anti-diagonals; 2 = the banded DP swept
 * level by level by one thread-block cluster.  All return identical costs and paths. */
int prgpu_nnf_set_mode(prgpu_nnf_t *h, int mode);
int prgpu_nnf_band_info(prgpu_nnf_t *h, double *tau, uint64_t *cells, uint32_t *rebuilds);
/* With search kernel 0 and a device world the whole LazySP loop of solve() runs in ONE persistent kernel: it
 * follows the optimum's hop words, evaluates the path's unknown NN edges, does the reference's bookkeeping in
 * path order, knocks the first colliding edge out and repairs the table locally (only vertices whose
 * predecessors' columns changed are recomputed -- what LPAStar::updateVertex / computeShortestPath do,
 * frechet/src/LPAStar.cpp:17-159); the host reads one status record per call or band rebuild.  launches: kernel
 * launches of that loop so far; repaired_vertices: columns recomputed by the local repairs; phase_cycles: SM
 * cycles of the loop's five phases (the block's thread-0 clock).  A repair wave that outgrows one block (more than
 * 2048 columns) is handed to the whole-GPU sweep from its current level, so an iteration never costs more than it
 * did with the host-side loop.  (Environment
 * PRGPU_NNF_HOST_LOOP=1 keeps the loop on the host, one full search per iteration: same results.) */
int prgpu_nnf_loop_stats(prgpu_nnf_t *h, uint64_t *launches, uint64_t *repaired_vertices,
                         uint64_t *phase_cycles /* 6, or NULL: SM cycles of hop chase, per-hop pass, edge evaluation,
                                                  bookkeeping, repair; [5] = repair waves handed to the whole-GPU sweep */);
/* duration of the latest search kernel (measurement aid, always on) */
int prgpu_nnf_last_dp_ms(prgpu_nnf_t *h, float *ms);

/* ------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU.  No counterpart in the reference (single-threaded,
 * no collectives); this is how the C++ host gets the sharded paths without Python.
 *   rank 0:  prgpu_comm_unique_id(id)  ->  the application broadcasts the 128 bytes (MPI_Bcast, a file,
 *            torch.distributed ...)  ->  every rank: prgpu_comm_init(device, rank, nranks, id, &comm)
 * NCCL is bound at run time (libnccl.so.2); prgpu_comm_init_custom takes the application's own all-gather
 * instead (an existing communicator, or the CPU/gloo tests; device = -1 makes a host-only communicator
 * that only prgpu_comm_allgather accepts).
 * ---------------------------------------------------------------------------------- */
typedef struct prgpu_comm prgpu_comm_t;
#define PRGPU_COMM_ID_BYTES 128
/* all-gather of `bytes_per_rank` bytes from every rank into recv (rank r's block at r * bytes_per_rank),
 * enqueued on `stream`; returns 0 on success */
typedef int (*prgpu_allgather_fn)(void *ctx, const void *send, void *recv, uint64_t bytes_per_rank, void *stream);
int prgpu_comm_unique_id(void *id128);
int prgpu_comm_init(int device, int rank, int nranks, const void *id128, prgpu_comm_t **out);
int prgpu_comm_init_custom(int device, int rank, int nranks, prgpu_allgather_fn fn, void *ctx, prgpu_comm_t **out);
int prgpu_comm_destroy(prgpu_comm_t *c);
int prgpu_comm_rank(const prgpu_comm_t *c);
int prgpu_comm_size(const prgpu_comm_t *c);
int prgpu_comm_allgather(prgpu_comm_t *c, const void *send, void *recv, uint64_t bytes_per_rank, void *stream);
/* sharded steps that had to be repeated because some shard left queries unproven (see below) */
int prgpu_comm_stats(const prgpu_comm_t *c, uint64_t *redo_steps);
/* Symmetric windows for the FUSED search + exchange: every rank allocates a window of 2 x bytes (+ flags), the
 * CUDA IPC handles travel over the communicator's own all-gather, and every rank maps every window (NVLink peer
 * memory).  With windows enabled, prgpu_knn_nearest_k_query_sharded_device lets the index search kernel write
 * each result row straight into every rank's window from its epilogue and exchanges arrival flags with one tiny
 * kernel: no collective call, no staging copy.  Collective.  Returns PRGPU_ENODEV (on every rank) when peer
 * memory cannot be mapped between these ranks: the all-gather path then stays in use.  bytes >= 12 * k * nq_local *
 * nranks of the largest step. */
int prgpu_comm_enable_peer_windows(prgpu_comm_t *c, uint64_t bytes);
/* steps that took the fused form so far */
int prgpu_comm_stats2(const prgpu_comm_t *c, uint64_t *fused_steps);
/* nearestK over a tree whose nodes are sharded over the ranks (each rank's handle holds its shard, with
 * prgpu_knn_set_index_base = the shard's first global index; queries replicated): local search, ONE
 * all-gather of the per-shard (distance, index) records, ordered merge.  Every rank receives the global
 * result, identical to a single-GPU search of the whole tree.  Collective: all ranks must call it. */
int prgpu_knn_nearest_k_sharded_device(prgpu_knn_t *h, prgpu_comm_t *c, const double *queries_dev, uint64_t nq, int k,
                                       uint32_t *idx_dev, double *dist_dev, void *stream);
int prgpu_knn_nearest_k_sharded(prgpu_knn_t *h, prgpu_comm_t *c, const double *queries, uint64_t nq, int k,
                                uint32_t *idx, double *dist);
/* The other partitioning (the tree fits one GPU and is held WHOLE by every rank; the QUERIES are split): this rank's
 * nq_local queries are searched locally -- the k-d index makes that sub-linear -- and one all-gather hands every rank
 * the answers of all ranks: idx_all / dist_all are [nranks][nq_local][k], rank-major.  Every rank passes the same
 * nq_local.  Collective; asynchronous on `stream` unless the local search had to fall back (rare). */
int prgpu_knn_nearest_k_query_sharded_device(prgpu_knn_t *h, prgpu_comm_t *c, const double *queries_dev, uint64_t nq_local,
                                             int k, uint32_t *idx_all_dev, double *dist_all_dev, void *stream);
/* the same with host buffers (queries in, all ranks' answers out), synchronous */
int prgpu_knn_nearest_k_query_sharded(prgpu_knn_t *h, prgpu_comm_t *c, const double *queries, uint64_t nq_local, int k,
                                      uint32_t *idx_all, double *dist_all);
/* checkMotion over an edge batch replicated on every rank: rank r validates its slice, one all-gather of
 * 1 byte per edge; every rank receives all n verdicts.  Collective. */
int prgpu_check_motion_batch_sharded_device(prgpu_world_t *w, prgpu_comm_t *c, const double *from_dev,
                                            const double *to_dev, uint64_t n, double resolution, int mode,
                                            int check_first, uint8_t *valid_dev, void *stream);

/* ------------------------------------------------------------------------------------
 * Measurement aids (no counterpart in the reference)
 * ---------------------------------------------------------------------------------- */
/* kernels launched by this library in this process so far */
uint64_t prgpu_diag_launch_count(void);
/* live micro-benchmark: issue peak of mma.sync.m16n8k8 TF32 (the instruction K1's scan uses), TFLOP/s */
int prgpu_diag_mma_tf32_peak(int device, double *tflops);
/* The numbers the exactness proofs of K1's filter path rest on, as the device computes them.  Runs the PRODUCTION
 * scan kernel over the whole tree for nq queries with the caller's thresholds T (thresh: in = any float, out = T
 * rounded up to a TF32 number, as the search does).  Per query: counts = nodes the scan reported (D < 0; only
 * min(counts, cap) are stored), idx = their local indices, D = (|x|^2 - 2 q.x - T) from the tensor cores,
 * key32 = (|x|^2 - 2 q.x) from the fp32 FMA chain of the re-cut, eps = (eps_tc, eps_32) as knn_select_kernel
 * computes them.  tests/test_gpu_bounds.py checks |D - exact| <= eps_tc, |key32 - exact| <= eps_32 and that no node
 * with exact key < T - eps_tc is missing. */
int prgpu_diag_knn_scan(prgpu_knn_t *h, const double *queries, uint32_t nq, float *thresh, uint32_t cap, uint32_t *counts,
                        uint32_t *idx /* nq x cap */, float *D /* nq x cap */, float *key32 /* nq x cap */,
                        double *eps /* nq x 2 */);
/* K1's index: the leaf capacity (slots: 64 or 96) and depth (2^levels leaves) prgpu_knn_build_index takes for n
 * nodes -- the capacity that leaves the fewest slots of the balanced tree's leaves empty.  Host arithmetic only. */
int prgpu_diag_kd_leaf_capacity(uint64_t n, int *slots, int *levels);
/* K2's filter: largest distance between a link-sphere centre from the production fp32 forward-kinematics chain and
 * the fp64 chain of the specification, per state (max_err: n); *delta = the bound prgpu_world_create derived. */
int prgpu_diag_fk_centre_error(prgpu_world_t *w, const double *states, uint64_t n, double *max_err, double *delta);

#ifdef __cplusplus
}
#endif
#endif /* PRGPU_H */

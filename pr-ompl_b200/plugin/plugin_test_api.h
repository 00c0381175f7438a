/* C entry points of libprgpu_plugin.so used by the parity tests: they drive the C++ OMPL-plugin
 * mirror (pr_gpu::RRTConnect over GpuNearestNeighbors / GpuMotionValidator / CounterSampler)
 * exactly the way an OMPL application would, and return the planner's trees. */
#ifndef PLUGIN_TEST_API_H
#define PLUGIN_TEST_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  int32_t status;
  uint32_t iterations, n_start, n_goal, path_len, overflow;
} plg_rrtc_result;
void *plg_world_create(int device, int dof, const double *dh, int ns, const int32_t *link, const double *xyzr,
                       int nos, const double *os, int nob, const double *ob);
void plg_world_destroy(void *world);
const char *plg_last_error(void);
/* nn_kind: 0 = GpuNearestNeighbors (default of the planner). */
int plg_rrtc_solve(void *world, int dim, const double *lo, const double *hi, const double *start,
                   const double *goal, double range, const char *extension_type, double resolution_fraction,
                   uint64_t seed, uint64_t query_id, uint32_t max_iters, uint32_t max_nodes, uint32_t max_path,
                   plg_rrtc_result *res, double *start_states, int32_t *start_parent, double *goal_states,
                   int32_t *goal_parent, double *path, uint32_t *planner_data_counts /* 4: vertices, edges, starts, goals */);
/* The next plg_rrtc_solve call stops after first_iters samples (use an even number: the tree alternation
 * restarts with the start tree on every solve(), RRTConnect.cpp:263-268), then -- clear_between != 0:
 * clear() and setup() again, sample counter back to 0 (RRTConnect.cpp:166-176) -- calls solve() a second
 * time up to max_iters.  first_iters == 0 restores the single-call behaviour. */
void plg_set_two_phase(uint32_t first_iters, int clear_between);
/* pr_gpu::BatchedRRTConnect through the plugin library: n queries, per query status / iterations / tree
 * sizes / path (paths: n x max_path x dim). */
int plg_batched_rrtc_solve(void *world, int dim, const double *lo, const double *hi, uint32_t n,
                           const double *starts, const double *goals, double range, const char *extension_type,
                           double resolution_fraction, uint64_t seed, uint64_t query_id_base, uint32_t max_iters,
                           uint32_t max_nodes, int32_t *status, uint32_t *iterations, uint32_t *n_start,
                           uint32_t *n_goal, uint32_t *path_len, double *paths, uint32_t max_path);
/* exercises the OMPL-facing classes directly: returns 0 if every check passes */
int plg_selftest_interfaces(void *world, const double *states, uint32_t n, uint8_t *is_valid_out,
                            uint8_t *motion_out /* n-1 consecutive edges */, uint32_t *nearest_out /* n */);
typedef struct {
  int32_t status;            /* PlannerStatus */
  uint32_t lazy_iters, n_evaluated, n_collisions, path_len, R;
  double final_error, goal_cost;
} plg_nnf_result;
/* drives NNFrechet::NNFrechet (the drop-in name) through its public API: the reference path is
 * ref_pose12 (P x 12), the IK callback pops pre-generated solutions (ik_states, in the planner's call
 * order) -- returns the planner's layer counts, path (NN-graph vertex ids) and path states. */
int plg_nnf_run(void *world, int W, int M, int k, int D, int seed, const double *ref_pose12, uint32_t P,
                const double *ik_states, uint32_t n_ik_available, double solve_seconds, plg_nnf_result *res,
                uint32_t *layer_counts /* R */, uint32_t *path_vertices, uint32_t cap_path, double *path_states);
/* the same, choosing the planner's form: metric_kind 0 = translation distance, 1 = translation + 0.25 x Frobenius
 * norm of the rotation difference; with_world = 0 leaves setWorld() out (FK callback = the world's FK, validity =
 * the SpaceInformation's checker, as a user of the reference would supply them).  *used_host_form reports which
 * form setup() took. */
int plg_nnf_run_host(void *world, int metric_kind, int with_world, int *used_host_form, int W, int M, int k, int D,
                     int seed, const double *ref_pose12, uint32_t P, const double *ik_states,
                     uint32_t n_ik_available, double solve_seconds, plg_nnf_result *res, uint32_t *layer_counts,
                     uint32_t *path_vertices, uint32_t cap_path, double *path_states);
#ifdef __cplusplus
}
#endif
#endif

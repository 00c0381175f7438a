// Internal interface of the batched RRTConnect kernel; see rrtc_batch.cu.
#pragma once
#include "world.cuh"
namespace prgpu {
struct RrtcDeviceParams {
  int dim;
  double lo[WORLD_MAX_DOF], hi[WORLD_MAX_DOF];
  double range;  // resolved (auto range already applied)
  double lvs;    // longest valid segment = extent * resolution fraction
  int ext_first, ext_second;
  uint64_t seed, query_id_base;
  uint32_t max_iters, max_nodes;
};
struct RrtcLoopState { // the solve loop's variables: what a suspended query needs to carry on in the team kernel
  uint32_t it, nS, nG, sampledGoals;
  int32_t connS, connG;
  int status;
  uint32_t startTree, full;
};
struct RrtcDeviceBuffers {
  const double *starts, *goals;       // B x dim
  double *start_states, *goal_states; // B x max_nodes x dim
  int32_t *start_parent, *goal_parent; // B x max_nodes
  int32_t *status;                    // B
  uint32_t *iters, *n_start, *n_goal; // B
  int32_t *conn_start, *conn_goal;    // B: ends of the solution path in each tree (-1 if none)
  uint32_t *overflow;                 // B
  uint32_t *work_counter;             // 3 consecutive words zeroed by the launch: next unstarted query,
  uint32_t *susp_count;               //   number of suspended queries,
  uint32_t *team_counter;             //   next suspended query of the team kernel
  uint32_t *susp_list;                // B: suspended query ids
  RrtcLoopState *resume;              // B
};
int rrtc_batch_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const RrtcDeviceParams &p,
                      const RrtcDeviceBuffers &b, uint32_t n_queries, int sm_count, cudaStream_t stream);
} // namespace prgpu

// K3: NNF (Nearest-Neighbor Frechet) device kernels.
//
//  * nnf_subsample_kernel  -- addSubsampledEdge (frechet/src/NNFrechet.cpp:345-396): the D
//    intermediate configurations first + (second-first)*(1+i)/(D+1) of every NN edge and their
//    end-effector poses (the FK callback), one thread per new vertex.
//  * nnf_dp_kernel         -- the min-bottleneck search of LPAStar::computeShortestPath
//    (frechet/src/LPAStar.cpp:125-159) on the tensor-product graph of
//    connectTensorProductNodes (NNFrechet.cpp:549-599), restated as a wavefront dynamic program:
//        B[r,n] = max(f[r,n], min(B[r-1,n], min over NN predecessors p of (B[r,p], B[r-1,p])))
//    with f[r,n] = |p_ref[r] - p_nn[n]| computed on the fly (addTensorProductNodes :495; the edge
//    weight max(f[u],f[v]) of :530 folds into the recurrence because B[u] >= f[u]).  Rows are
//    reference poses, columns NN-graph vertices sorted by topological level; all cells with
//    row + level = d are independent, so one cooperative persistent kernel sweeps the
//    anti-diagonals with a grid-wide barrier between them.  Knocked-out NN edges
//    (markEdgeInCollision :685-702) are a byte mask consulted per predecessor.
//  * nnf_backtrack_kernel  -- followBackpointers (LPAStar.cpp:171-189) by re-deriving, at each
//    cell, the first predecessor that attains the cell's cost.
#include <cfloat>
#include <cooperative_groups.h>
#include "nnf.cuh"

namespace cg = cooperative_groups;

namespace prgpu {
namespace {

constexpr double NNF_INF = DBL_MAX; // LPAStar.h:57

__global__ void nnf_subsample_kernel(const DeviceWorld *__restrict__ wp, int dof, uint32_t n_new,
                                     uint32_t first_vertex, const uint32_t *__restrict__ src,
                                     const uint32_t *__restrict__ dst, const uint32_t *__restrict__ step, int D,
                                     double *__restrict__ states, double *__restrict__ pose12) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_new)
    return;
  const DeviceWorld &w = *wp;
  const double t = (1.0 + (double)step[i]) / (double)(D + 1); // NNFrechet.cpp:367
  const double *a = states + (size_t)src[i] * dof;
  const double *b = states + (size_t)dst[i] * dof;
  double q[WORLD_MAX_DOF];
#pragma unroll
  for (int k = 0; k < WORLD_MAX_DOF; ++k)
    q[k] = k < dof ? lerp_exact(a[k], b[k], t) : 0.0;
  const size_t v = (size_t)first_vertex + i;
  for (int k = 0; k < dof; ++k)
    states[v * dof + k] = q[k];
  double F[12 * (WORLD_MAX_DOF + 1)];
  fk_frames(w, q, F);
  for (int k = 0; k < 12; ++k)
    pose12[v * 12 + k] = F[12 * w.dof + k];
}

__global__ void nnf_translations_kernel(const double *__restrict__ pose12, const uint32_t *__restrict__ vertex_of_pos,
                                        uint32_t n, double *__restrict__ out3) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
    return;
  const size_t v = vertex_of_pos ? vertex_of_pos[s] : s;
  out3[(size_t)s * 3 + 0] = pose12[v * 12 + 3];
  out3[(size_t)s * 3 + 1] = pose12[v * 12 + 7];
  out3[(size_t)s * 3 + 2] = pose12[v * 12 + 11];
}

__device__ __forceinline__ double task_distance(const double *a3, const double *b3) {
  const double dx = __dsub_rn(a3[0], b3[0]), dy = __dsub_rn(a3[1], b3[1]), dz = __dsub_rn(a3[2], b3[2]);
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

__global__ void __launch_bounds__(256) nnf_dp_kernel(const NnfGraphDev g) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t gsize = gridDim.x * blockDim.x;
  const int R = g.R;
  const uint32_t V = g.V;
  const uint32_t ndiag = (uint32_t)R + g.Lv - 1;
  for (uint32_t diag = 0; diag < ndiag; ++diag) {
    const uint32_t l_lo = diag >= (uint32_t)R ? diag - (uint32_t)(R - 1) : 0u;
    const uint32_t l_hi = diag < g.Lv ? diag : g.Lv - 1;
    const uint32_t s0 = g.lvl_start[l_lo], s1 = g.lvl_start[l_hi + 1];
    for (uint32_t s = s0 + gtid; s < s1; s += gsize) {
      const uint32_t r = diag - g.level[s];
      double best = NNF_INF;
      if (r == 0 && s == g.start_pos) {
        best = 0.0; // rhs(start) = 0 (LPAStar.cpp:13)
      } else {
        double m = NNF_INF; // min over predecessors of B
        if (r > 0)
          m = g.B[(size_t)(r - 1) * V + s]; // step A (NNFrechet.cpp:567-574)
        const uint32_t e1 = g.pred_off[s + 1];
        for (uint32_t e = g.pred_off[s]; e < e1; ++e) {
          if (g.blocked[g.pred_eid[e]])
            continue;
          const uint32_t p = g.pred_pos[e];
          m = fmin(m, g.B[(size_t)r * V + p]); // step B (:577-583)
          if (r > 0)
            m = fmin(m, g.B[(size_t)(r - 1) * V + p]); // step C (:586-596)
        }
        if (m != NNF_INF) {
          const double fv = task_distance(g.pref + 3 * (size_t)r, g.pnn + 3 * (size_t)s);
          best = fmax(m, fv);
        }
      }
      g.B[(size_t)r * V + s] = best;
    }
    grid.sync();
  }
}

__global__ void nnf_backtrack_kernel(const NnfGraphDev g, uint32_t *__restrict__ out_path, uint32_t cap_pairs,
                                     uint32_t *__restrict__ out_n, double *__restrict__ out_cost) {
  if (blockIdx.x != 0 || threadIdx.x != 0)
    return;
  const uint32_t V = g.V;
  uint32_t r = (uint32_t)g.R - 1, s = g.goal_pos, n = 0;
  const double goal = g.B[(size_t)r * V + s];
  *out_cost = goal;
  if (goal == NNF_INF) { // followBackpointers: empty path (LPAStar.cpp:174-175)
    *out_n = 0;
    return;
  }
  while (true) {
    if (n < cap_pairs) {
      out_path[2 * n] = r;
      out_path[2 * n + 1] = s;
    }
    ++n;
    if (r == 0 && s == g.start_pos)
      break;
    const double target = g.B[(size_t)r * V + s];
    const double fv = task_distance(g.pref + 3 * (size_t)r, g.pnn + 3 * (size_t)s);
    bool found = false;
    uint32_t br = 0, bs = 0;
    // first predecessor (fixed order) whose cost explains this cell's cost
    if (r > 0) {
      const double bu = g.B[(size_t)(r - 1) * V + s];
      if (bu != NNF_INF && fmax(bu, fv) == target) {
        found = true;
        br = r - 1;
        bs = s;
      }
    }
    for (uint32_t e = g.pred_off[s]; !found && e < g.pred_off[s + 1]; ++e) {
      if (g.blocked[g.pred_eid[e]])
        continue;
      const uint32_t p = g.pred_pos[e];
      const double b0 = g.B[(size_t)r * V + p];
      if (b0 != NNF_INF && fmax(b0, fv) == target) {
        found = true;
        br = r;
        bs = p;
        break;
      }
      if (r > 0) {
        const double b1 = g.B[(size_t)(r - 1) * V + p];
        if (b1 != NNF_INF && fmax(b1, fv) == target) {
          found = true;
          br = r - 1;
          bs = p;
          break;
        }
      }
    }
    if (!found) { // cannot happen on a consistent table
      n = 0xFFFFFFFFu;
      break;
    }
    r = br;
    s = bs;
  }
  *out_n = n;
}

} // namespace

int nnf_subsample_launch(const DeviceWorld *w_dev, int dof, uint32_t n_new, uint32_t first_vertex,
                         const uint32_t *src, const uint32_t *dst, const uint32_t *step, int D, double *states,
                         double *pose12, cudaStream_t stream) {
  if (n_new == 0)
    return PRGPU_OK;
  nnf_subsample_kernel<<<(n_new + 127) / 128, 128, 0, stream>>>(w_dev, dof, n_new, first_vertex, src, dst, step, D,
                                                               states, pose12);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int nnf_translations_launch(const double *pose12, const uint32_t *vertex_of_pos, uint32_t n, double *out3,
                            cudaStream_t stream) {
  if (n == 0)
    return PRGPU_OK;
  nnf_translations_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pose12, vertex_of_pos, n, out3);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int nnf_dp_launch(const NnfGraphDev &g, int sm_count, cudaStream_t stream, KernelTimer *timer) {
  int per_sm = 0;
  PRGPU_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nnf_dp_kernel, 256, 0));
  if (per_sm < 1)
    return fail(PRGPU_ECUDA, "nnf_dp_kernel cannot be made resident");
  dim3 grid((unsigned)(per_sm * sm_count)), block(256);
  NnfGraphDev arg = g;
  void *args[] = {&arg};
  if (timer)
    timer->begin(stream);
  PRGPU_CUDA_TRY(cudaLaunchCooperativeKernel((void *)nnf_dp_kernel, grid, block, args, 0, stream));
  if (timer)
    timer->end(stream);
  return PRGPU_OK;
}

int nnf_backtrack_launch(const NnfGraphDev &g, uint32_t *out_path, uint32_t cap_pairs, uint32_t *out_n,
                         double *out_cost, cudaStream_t stream) {
  nnf_backtrack_kernel<<<1, 32, 0, stream>>>(g, out_path, cap_pairs, out_n, out_cost);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

} // namespace prgpu

// K2: batched discretised edge validation.
//
// Replaces si_->checkMotion / isValid as called by the reference (pr_ompl/src/RRTConnect.cpp:198,
// i.e. OMPL's DiscreteMotionValidator::checkMotion over a StateValidityChecker) and
// NNFrechet::evaluateEdge (frechet/src/NNFrechet.cpp:655-683).
//
// One warp per edge.  The warp derives the edge's tested set exactly as the reference does
// (nd = ceil(distance / longestValidSegment) with the bit-exact distance; states j/nd via
// from + (to - from) * t; or NNF's accumulated alpha sequence), lanes take the states of a round
// in an order that spreads every round over the whole edge, and a ballot ends the edge at the
// first round containing an invalid state (the result is the AND over the set, so evaluation
// order is free).  Obstacles sit in shared memory as fp32 tables; see world.cuh for the
// filter / exact split.
#include "check_motion.cuh"

namespace prgpu {
namespace {

constexpr int CM_THREADS = 128;

template <int SP>
__global__ void __launch_bounds__(CM_THREADS)
check_motion_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ from, const double *__restrict__ to,
                    uint64_t n, double resolution, int mode, int check_first,
                    uint8_t *__restrict__ valid) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  float4 *s_os = s_obs;
  float4 *s_ob = s_obs + w.n_os;
  load_obstacles(w, s_os, s_ob);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const uint64_t warps_total = (uint64_t)gridDim.x * (CM_THREADS / 32);
  const int dof = w.dof;

  for (uint64_t e = (uint64_t)blockIdx.x * (CM_THREADS / 32) + (threadIdx.x >> 5); e < n; e += warps_total) {
    double a[WORLD_MAX_DOF], b[WORLD_MAX_DOF];
#pragma unroll
    for (int i = 0; i < WORLD_MAX_DOF; ++i) {
      a[i] = i < dof ? from[e * dof + i] : 0.0;
      b[i] = i < dof ? to[e * dof + i] : 0.0;
    }
    // RealVectorStateSpace::distance, bit exact
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < WORLD_MAX_DOF; ++i) {
      const double diff = __dsub_rn(a[i], b[i]);
      sum = __dadd_rn(sum, __dmul_rn(diff, diff));
    }
    const double len = __dsqrt_rn(sum);

    uint32_t np;      // number of states in the tested set
    uint32_t nd = 0;  // mode 0: segment count
    double step = 0.0;
    if (mode == PRGPU_MODE_OMPL_DISCRETE) {
      nd = (uint32_t)ceil(len / resolution);
      np = 1u + (nd >= 2 ? nd - 1 : 0u) + (check_first ? 1u : 0u);
    } else {
      // for (alpha = 0; alpha <= 1; alpha += 1.0/numQ): count the accumulated sequence
      const int numQ = (int)ceil(len / resolution);
      step = 1.0 / (double)numQ; // numQ == 0 -> +inf -> a single state at alpha = 0
      np = 0;
      for (double alpha = 0.0; alpha <= 1.0; alpha = __dadd_rn(alpha, step))
        ++np;
    }

    const uint32_t rounds = (np + 31) / 32;
    bool ok = true;
    for (uint32_t r = 0; r < rounds; ++r) {
      const uint32_t p = r + (uint32_t)lane * rounds; // spreads each round over the edge
      bool good = true;
      if (p < np) {
        double q[WORLD_MAX_DOF];
        if (mode == PRGPU_MODE_OMPL_DISCRETE) {
          const uint32_t extra = check_first ? 1u : 0u;
          if (p == 0) {
#pragma unroll
            for (int i = 0; i < WORLD_MAX_DOF; ++i)
              q[i] = b[i];
          } else if (p <= extra) {
#pragma unroll
            for (int i = 0; i < WORLD_MAX_DOF; ++i)
              q[i] = a[i];
          } else {
            const uint32_t j = p - extra; // 1 .. nd-1
            const double t = (double)j / (double)nd;
#pragma unroll
            for (int i = 0; i < WORLD_MAX_DOF; ++i)
              q[i] = lerp_exact(a[i], b[i], t);
          }
        } else {
          double alpha = 0.0;
          for (uint32_t i = 0; i < p; ++i)
            alpha = __dadd_rn(alpha, step);
#pragma unroll
          for (int i = 0; i < WORLD_MAX_DOF; ++i)
            q[i] = lerp_exact(a[i], b[i], alpha);
        }
        good = state_valid<SP>(w, s_os, s_ob, q);
      }
      if (__any_sync(0xffffffffu, !good)) {
        ok = false;
        break;
      }
    }
    if (lane == 0)
      valid[e] = ok ? 1 : 0;
  }
}

template <int SP>
__global__ void __launch_bounds__(CM_THREADS)
is_valid_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ qs, uint64_t n, uint8_t *__restrict__ valid) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  float4 *s_os = s_obs;
  float4 *s_ob = s_obs + w.n_os;
  load_obstacles(w, s_os, s_ob);
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double q[WORLD_MAX_DOF];
#pragma unroll
    for (int d = 0; d < WORLD_MAX_DOF; ++d)
      q[d] = d < w.dof ? qs[i * w.dof + d] : 0.0;
    valid[i] = state_valid<SP>(w, s_os, s_ob, q) ? 1 : 0;
  }
}

__global__ void fk_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ qs, uint64_t n,
                          double *__restrict__ pose12) {
  const DeviceWorld &w = *wp;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  double q[WORLD_MAX_DOF];
#pragma unroll
  for (int d = 0; d < WORLD_MAX_DOF; ++d)
    q[d] = d < w.dof ? qs[i * w.dof + d] : 0.0;
  double F[12 * (WORLD_MAX_DOF + 1)];
  fk_frames(w, q, F);
  for (int k = 0; k < 12; ++k)
    pose12[i * 12 + k] = F[12 * w.dof + k];
}

size_t obstacle_smem(const DeviceWorld &w) { return ((size_t)w.n_os + 2 * (size_t)w.n_ob) * sizeof(float4); }

} // namespace

#define PRGPU_SP_DISPATCH(S, ...)                                                                  \
  do {                                                                                             \
    if ((S) <= 8) { constexpr int SP = 8; __VA_ARGS__; }                                           \
    else if ((S) <= 16) { constexpr int SP = 16; __VA_ARGS__; }                                    \
    else if ((S) <= 24) { constexpr int SP = 24; __VA_ARGS__; }                                    \
    else { constexpr int SP = 32; __VA_ARGS__; }                                                   \
  } while (0)

int check_motion_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *from_dev, const double *to_dev, uint64_t n,
                        double resolution, int mode, int check_first, uint8_t *valid_dev, int sm_count,
                        cudaStream_t stream, KernelTimer *timer) {
  if (n == 0)
    return PRGPU_OK;
  if (!(resolution > 0.0))
    return fail(PRGPU_EINVAL, "check_motion: resolution must be > 0");
  if (mode != PRGPU_MODE_OMPL_DISCRETE && mode != PRGPU_MODE_NNF_ALPHA)
    return fail(PRGPU_EINVAL, "check_motion: unknown mode");
  const size_t smem = obstacle_smem(w);
  const uint64_t warps = n;
  uint64_t blocks = (warps + CM_THREADS / 32 - 1) / (CM_THREADS / 32);
  const uint64_t cap = (uint64_t)sm_count * 8;
  if (blocks > cap)
    blocks = cap;
  if (timer)
    timer->begin(stream);
  PRGPU_SP_DISPATCH(w.S, {
    auto kern = check_motion_kernel<SP>;
    PRGPU_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, CM_THREADS, smem, stream>>>(w_dev, from_dev, to_dev, n, resolution, mode,
                                                         check_first, valid_dev);
  });
  if (timer)
    timer->end(stream);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int is_valid_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *q_dev, uint64_t n, uint8_t *valid_dev,
                    cudaStream_t stream) {
  if (n == 0)
    return PRGPU_OK;
  const size_t smem = obstacle_smem(w);
  uint64_t blocks = (n + CM_THREADS - 1) / CM_THREADS;
  if (blocks > 148 * 8)
    blocks = 148 * 8;
  PRGPU_SP_DISPATCH(w.S, {
    auto kern = is_valid_kernel<SP>;
    PRGPU_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, CM_THREADS, smem, stream>>>(w_dev, q_dev, n, valid_dev);
  });
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int fk_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *q_dev, uint64_t n, double *pose12_dev, cudaStream_t stream) {
  if (n == 0)
    return PRGPU_OK;
  fk_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(w_dev, q_dev, n, pose12_dev);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

} // namespace prgpu

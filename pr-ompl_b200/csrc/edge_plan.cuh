// The tested set of one edge, exactly as the reference derives it: OMPL's DiscreteMotionValidator
// (nd = ceil(distance / resolution) segments; states from + (to - from) * j / nd, the end state first) or
// NNFrechet::evaluateEdge's accumulated alpha (frechet/src/NNFrechet.cpp:668-673).  Shared by the K2 kernels
// (check_motion.cu) and the device-resident LazySP loop of K3 (nnf.cu).
#pragma once
#include "world.cuh"
namespace prgpu {
struct EdgePlan {
  uint32_t np, nd, extra, stride; // tested-set size, segment count, 1 if `from` is tested, sample stride
  double step;                    // NNF mode: 1/numQ
};

__device__ __forceinline__ EdgePlan plan_edge(const double *a, const double *b, double resolution, int mode,
                                              int check_first) {
  double sum = 0.0;
#pragma unroll
  for (int i = 0; i < WORLD_MAX_DOF; ++i) {
    const double diff = __dsub_rn(a[i], b[i]);
    sum = __dadd_rn(sum, __dmul_rn(diff, diff));
  }
  const double len = __dsqrt_rn(sum); // RealVectorStateSpace::distance, bit exact
  EdgePlan p;
  p.nd = 0;
  p.extra = 0;
  p.step = 0.0;
  if (mode == PRGPU_MODE_OMPL_DISCRETE) {
    p.nd = (uint32_t)ceil(len / resolution);
    p.extra = check_first ? 1u : 0u;
    p.np = 1u + (p.nd >= 2 ? p.nd - 1 : 0u) + p.extra;
  } else {
    const int numQ = (int)ceil(len / resolution);
    p.step = 1.0 / (double)numQ;
    p.np = 0;
    for (double alpha = 0.0; alpha <= 1.0; alpha = __dadd_rn(alpha, p.step))
      ++p.np;
  }
  p.stride = (p.np + 7) / 8;
  if (p.stride < 1)
    p.stride = 1;
  return p;
}

__device__ __forceinline__ void edge_state(const EdgePlan &pl, const double *a, const double *b, int mode, uint32_t p,
                                           double *q) {
  if (mode == PRGPU_MODE_OMPL_DISCRETE) {
    if (p == 0) {
#pragma unroll
      for (int i = 0; i < WORLD_MAX_DOF; ++i)
        q[i] = b[i];
    } else if (p <= pl.extra) {
#pragma unroll
      for (int i = 0; i < WORLD_MAX_DOF; ++i)
        q[i] = a[i];
    } else {
      const double t = (double)(p - pl.extra) / (double)pl.nd;
#pragma unroll
      for (int i = 0; i < WORLD_MAX_DOF; ++i)
        q[i] = lerp_exact(a[i], b[i], t);
    }
  } else {
    double alpha = 0.0;
    for (uint32_t i = 0; i < p; ++i)
      alpha = __dadd_rn(alpha, pl.step);
#pragma unroll
    for (int i = 0; i < WORLD_MAX_DOF; ++i)
      q[i] = lerp_exact(a[i], b[i], alpha);
  }
}
} // namespace prgpu

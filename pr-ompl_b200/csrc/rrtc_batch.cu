// T1: many independent RRTConnect queries, one WARP per query, resident on the device from the first
// sample to the solution.
//
// Follows pr_ompl::RRTConnect::solve / growTree (pr_ompl/src/RRTConnect.cpp:219-376, :178-217)
// statement by statement -- tree alternation before the TRAPPED `continue` (:267 vs :303), goal
// seeding (:270-286), CONNECT as repeated growTree with a fresh nearest query each step (:296-300),
// rstate overwritten by the last added state (:313-314), goal-tree edges validated from the new
// state toward the tree (:198), one step back on the start side at the connection (:332-335).
// The 32 lanes of a warp cooperate inside each growTree: the linear nearest scan with the
// (distance, index) order of NearestNeighborsLinear (butterfly reduction, no shared memory), then the
// DiscreteMotionValidator tested set spread over the lanes with a vote for the early exit.  A query is a
// long dependent chain (sample -> nearest -> validate -> append), so its speed is set by latency, not
// width: one warp per query needs no block barrier at all, puts twice as many queries on an SM as a
// 64-thread block per query did, and the warps are persistent -- each takes the next unstarted query
// from a global counter when it finishes one, so the few queries that run to the sample budget do
// not hold finished neighbours' registers.  The obstacle tables in shared memory are shared by the
// RB_WARPS queries of a block.
#include <cstdio>
#include "rrtc_batch.cuh"
#include "sample_stream.cuh"

namespace prgpu {
namespace {

constexpr int RB_WARPS = 8;            // queries in flight per block
constexpr int RB_THREADS = RB_WARPS * 32;
enum { GS_TRAPPED = 0, GS_ADVANCED = 1, GS_REACHED = 2 };
enum { ST_INVALID_START = 1, ST_INVALID_GOAL = 2, ST_TIMEOUT = 4, ST_EXACT = 6 }; // PlannerStatus values

struct Shared {
  double rstate[WORLD_MAX_DOF];
  double xstate[WORLD_MAX_DOF];
  double near[WORLD_MAX_DOF];
};

template <int SP>
__global__ void __launch_bounds__(RB_THREADS)
rrtc_batch_kernel(const DeviceWorld *__restrict__ wp, const RrtcDeviceParams p, const RrtcDeviceBuffers b,
                  uint32_t n_queries) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  __shared__ Shared sh_all[RB_WARPS];
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  Shared &sh = sh_all[threadIdx.x >> 5];
  const int tid = lane; // lane of the query's warp
  const int dim = p.dim;
 for (;;) { // persistent warp: next unstarted query
  uint32_t qid = 0;
  if (lane == 0)
    qid = atomicAdd(b.work_counter, 1u);
  qid = __shfl_sync(0xffffffffu, qid, 0);
  if (qid >= n_queries)
    return;
  double *tS = b.start_states + (size_t)qid * p.max_nodes * dim;
  double *tG = b.goal_states + (size_t)qid * p.max_nodes * dim;
  int32_t *pS = b.start_parent + (size_t)qid * p.max_nodes;
  int32_t *pG = b.goal_parent + (size_t)qid * p.max_nodes;
  const double *start = b.starts + (size_t)qid * dim;
  const double *goal = b.goals + (size_t)qid * dim;
  uint32_t nS = 0, nG = 0; // uniform across the warp
#ifdef RB_PROFILE
  long long prof_scan = 0, prof_plan = 0, prof_valid = 0, prof_total = clock64(), prof_grows = 0, prof_states = 0;
#endif

  // warp-uniform helpers ------------------------------------------------------------------
  auto in_bounds = [&](const double *q) { // RealVectorStateSpace::satisfiesBounds
    bool ok = true;
    for (int i = 0; i < dim; ++i)
      if (q[i] - 2.220446049250313e-16 > p.hi[i] || q[i] + 2.220446049250313e-16 < p.lo[i])
        ok = false;
    return ok;
  };
  auto single_valid = [&](const double *q) { // isValid of one state, evaluated by lane 0
    int f = 0;
    if (tid == 0) {
      double qq[WORLD_MAX_DOF];
#pragma unroll
      for (int i = 0; i < WORLD_MAX_DOF; ++i)
        qq[i] = i < dim ? q[i] : 0.0;
      f = (in_bounds(qq) && state_valid_call<SP>(wp, &tabs, qq)) ? 1 : 0;
    }
    return __shfl_sync(0xffffffffu, f, 0) != 0;
  };
  auto append = [&](double *T, int32_t *P, uint32_t &n, const double *q, int32_t parent) {
    if (tid < dim)
      T[(size_t)n * dim + tid] = q[tid];
    if (tid == 0)
      P[n] = parent;
    ++n;
    __syncwarp(); // the new node is visible to the next nearest scan
  };

  // growTree (RRTConnect.cpp:178-217); returns GrowState, sets xmotion on success
  auto grow = [&](double *T, int32_t *P, uint32_t &n, bool tgi_start, int32_t &xmotion, bool &full) -> int {
    // :181 nearest = first strict minimum of sqrt-distance in insertion order
    // The scan is the latency chain of a long-running query (its trees grow to ~10^3 nodes): four nodes
    // per lane are summed as independent chains, and the square root -- monotone, so a node whose
    // squared sum is not below the best one cannot have a smaller distance -- is taken only for a node
    // that may replace the best.
#ifdef RB_PROFILE
    long long c0 = clock64();
    ++prof_grows;
#endif
    double bd = CUDART_INF, bs = CUDART_INF;
    int32_t bi = 0x7FFFFFFF;
    double rq[WORLD_MAX_DOF];
#pragma unroll
    for (int k = 0; k < WORLD_MAX_DOF; ++k)
      rq[k] = k < dim ? sh.rstate[k] : 0.0;
    auto consider = [&](double sq, uint32_t i) {
      if (sq < bs) { // increasing i per lane: strict < keeps the lowest index
        const double d = __dsqrt_rn(sq);
        if (d < bd) {
          bd = d;
          bi = (int32_t)i;
          bs = sq;
        }
      }
    };
    uint32_t i = tid;
    for (; i + 96 < n; i += 128) {
      const double *x0 = T + (size_t)i * dim, *x1 = x0 + 32 * dim, *x2 = x0 + 64 * dim, *x3 = x0 + 96 * dim;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int k = 0; k < WORLD_MAX_DOF; ++k)
        if (k < dim) {
          const double d0 = __dsub_rn(x0[k], rq[k]), d1 = __dsub_rn(x1[k], rq[k]);
          const double d2 = __dsub_rn(x2[k], rq[k]), d3 = __dsub_rn(x3[k], rq[k]);
          s0 = __dadd_rn(s0, __dmul_rn(d0, d0));
          s1 = __dadd_rn(s1, __dmul_rn(d1, d1));
          s2 = __dadd_rn(s2, __dmul_rn(d2, d2));
          s3 = __dadd_rn(s3, __dmul_rn(d3, d3));
        }
      consider(s0, i);
      consider(s1, i + 32);
      consider(s2, i + 64);
      consider(s3, i + 96);
    }
    for (; i < n; i += 32) {
      const double *x = T + (size_t)i * dim;
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < WORLD_MAX_DOF; ++k)
        if (k < dim) {
          const double diff = __dsub_rn(x[k], rq[k]);
          sq = __dadd_rn(sq, __dmul_rn(diff, diff));
        }
      consider(sq, i);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, off);
      const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (od < bd || (od == bd && oi < bi)) {
        bd = od;
        bi = oi;
      }
    }
    const int32_t nmotion = bi; // the butterfly leaves the same (distance, index) minimum in every lane
    if (tid < dim)
      sh.near[tid] = T[(size_t)nmotion * dim + tid];
    __syncwarp();

#ifdef RB_PROFILE
    long long c1 = clock64();
    prof_scan += c1 - c0;
#endif
    // :188-194 distance and step clamp (bit exact; every thread computes the same values)
    double a[WORLD_MAX_DOF], c[WORLD_MAX_DOF]; // a = nmotion->state, c = dstate
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < WORLD_MAX_DOF; ++k) {
      a[k] = k < dim ? sh.near[k] : 0.0;
      c[k] = k < dim ? sh.rstate[k] : 0.0;
      const double diff = __dsub_rn(a[k], c[k]);
      s = __dadd_rn(s, __dmul_rn(diff, diff));
    }
    const double d = __dsqrt_rn(s);
    bool reach = true;
    if (d > p.range) {
      const double t = p.range / d;
#pragma unroll
      for (int k = 0; k < WORLD_MAX_DOF; ++k)
        c[k] = k < dim ? lerp_exact(a[k], c[k], t) : 0.0;
      reach = false;
    }
    // :198 start tree: checkMotion(nmotion, dstate); goal tree: isValid(dstate) && checkMotion(dstate, nmotion)
    // DiscreteMotionValidator: s1 = from, s2 = to; tested set {s2} U {from + (to-from) j/nd} (+ dstate itself)
    double *from = tgi_start ? a : c;
    double *to = tgi_start ? c : a;
    double s2 = 0.0;
#pragma unroll
    for (int k = 0; k < WORLD_MAX_DOF; ++k) {
      const double diff = __dsub_rn(from[k], to[k]);
      s2 = __dadd_rn(s2, __dmul_rn(diff, diff));
    }
    const uint32_t nd = (uint32_t)ceil(__dsqrt_rn(s2) / p.lvs);
    const uint32_t extra = tgi_start ? 0u : 1u;
    const uint32_t np = 1u + (nd >= 2 ? nd - 1 : 0u) + extra;
    const uint32_t rounds = (np + 31) / 32;
#ifdef RB_PROFILE
    long long c2 = clock64();
    prof_plan += c2 - c1;
    prof_states += np;
#endif
    bool ok = true;
    for (uint32_t r = 0; r < rounds; ++r) {
      const uint32_t pt = r + (uint32_t)tid * rounds;
      bool good = true;
      if (pt < np) {
        double q[WORLD_MAX_DOF];
        if (pt == 0) {
#pragma unroll
          for (int k = 0; k < WORLD_MAX_DOF; ++k)
            q[k] = to[k];
        } else if (pt <= extra) {
#pragma unroll
          for (int k = 0; k < WORLD_MAX_DOF; ++k)
            q[k] = from[k];
        } else {
          const double t = (double)(pt - extra) / (double)nd;
#pragma unroll
          for (int k = 0; k < WORLD_MAX_DOF; ++k)
            q[k] = lerp_exact(from[k], to[k], t);
        }
        good = state_valid_call<SP>(wp, &tabs, q);
      }
      if (__any_sync(0xffffffffu, !good)) {
        ok = false;
        break;
      }
    }
#ifdef RB_PROFILE
    prof_valid += clock64() - c2;
#endif
    if (!ok)
      return GS_TRAPPED; // :215
    if (n >= p.max_nodes) {
      full = true;
      return GS_TRAPPED;
    }
    if (tid < dim)
      sh.xstate[tid] = c[tid];
    __syncwarp();
    xmotion = (int32_t)n;
    append(T, P, n, sh.xstate, nmotion); // :200-209
    return reach ? GS_REACHED : GS_ADVANCED;
  };

  // solve (RRTConnect.cpp:219-376) ----------------------------------------------------------
  int status = ST_TIMEOUT;
  int32_t connS = -1, connG = -1;
  uint32_t it = 0;
  bool full = false;
  if (single_valid(start)) { // :230-236 via PlannerInputStates::nextStart
    if (tid < dim)
      sh.xstate[tid] = start[tid];
    __syncwarp();
    append(tS, pS, nS, sh.xstate, -1);
  }
  if (nS == 0) {
    status = ST_INVALID_START; // :238-242
  } else {
    bool startTree = true;
    uint32_t sampledGoals = 0;
    while (it < p.max_iters && !full) { // :263, ptc = sample budget
      const bool growStart = startTree; // tree = startTree ? tStart : tGoal
      bool tgi_start = startTree;
      startTree = !startTree; // :267
      if (nG == 0 || sampledGoals < nG / 2) { // :270
        if (sampledGoals < 1) { // a GoalState yields one sample ever (Appendix B.7)
          sampledGoals++;
          if (single_valid(goal)) {
            if (tid < dim)
              sh.xstate[tid] = goal[tid];
            __syncwarp();
            append(tG, pG, nG, sh.xstate, -1);
          }
        }
        if (nG == 0) // :281-285
          break;
      }
      if (tid < dim) // :289
        sh.rstate[tid] = stream_uniform(p.seed, p.query_id_base + qid, it, tid, p.lo[tid], p.hi[tid]);
      __syncwarp();
      ++it;

      int gs = GS_TRAPPED;
      int32_t xmotion = -1, added = -1; // :295
      bool second_done = false;
      for (int phase = 0; phase < 2; ++phase) {
        // phase 0 grows `tree` toward the sample (:296-300), phase 1 grows `otherTree` toward the
        // node just added (:316-321)
        const bool onStart = (phase == 0) ? growStart : !growStart;
        const int ext = (phase == 0) ? p.ext_first : p.ext_second;
        do {
          gs = grow(onStart ? tS : tG, onStart ? pS : pG, onStart ? nS : nG, tgi_start, xmotion, full);
        } while (gs == GS_ADVANCED && ext == 1);
        if (phase == 1) {
          second_done = true;
          break;
        }
        if (xmotion < 0) // :303
          break;
        added = xmotion; // :307
        xmotion = -1;
        if (gs != GS_REACHED) { // :313-314
          const double *src = (growStart ? tS : tG) + (size_t)added * dim;
          __syncwarp();
          if (tid < dim)
            sh.rstate[tid] = src[tid];
          __syncwarp();
        }
        tgi_start = startTree; // :316
      }
      if (!second_done)
        continue;

      if (gs == GS_REACHED) { // :327 (isStartGoalPairValid is true for a GoalState)
        int32_t startMotion = startTree ? xmotion : added; // :323-324
        int32_t goalMotion = startTree ? added : xmotion;
        if (pS[startMotion] >= 0) // :332-335
          startMotion = pS[startMotion];
        else
          goalMotion = pG[goalMotion];
        connS = startMotion;
        connG = goalMotion;
        status = ST_EXACT;
        break;
      }
    }
  }
#ifdef RB_PROFILE
  if (tid == 0 && it >= p.max_iters && (qid % 37) == 0)
    printf("q %u it %u grows %lld states %lld total %lld scan %lld plan %lld valid %lld\n", qid, it, prof_grows, prof_states,
           clock64() - prof_total, prof_scan, prof_plan, prof_valid);
#endif
  if (tid == 0) {
    b.status[qid] = status;
    b.iters[qid] = it;
    b.n_start[qid] = nS;
    b.n_goal[qid] = nG;
    b.conn_start[qid] = connS;
    b.conn_goal[qid] = connG;
    b.overflow[qid] = full ? 1u : 0u;
  }
  __syncwarp();
 } // next query
}

} // namespace

int rrtc_batch_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const RrtcDeviceParams &p,
                      const RrtcDeviceBuffers &b, uint32_t n_queries, int sm_count, cudaStream_t stream) {
  if (n_queries == 0)
    return PRGPU_OK;
  const size_t smem = obstacle_tables_bytes(w);
  PRGPU_CUDA_TRY(cudaMemsetAsync(b.work_counter, 0, sizeof(uint32_t), stream));
#define PRGPU_RB_LAUNCH(SPV)                                                                       \
  do {                                                                                             \
    auto kern = rrtc_batch_kernel<SPV>;                                                            \
    PRGPU_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int per_sm = 0;                                                                                \
    PRGPU_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RB_THREADS, smem)); \
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_queries + RB_WARPS - 1) / RB_WARPS,    \
                                                         (uint64_t)std::max(per_sm, 1) * sm_count); \
    kern<<<blocks, RB_THREADS, smem, stream>>>(w_dev, p, b, n_queries);                            \
  } while (0)
  if (w.grid_n[0] > 0)
    PRGPU_RB_LAUNCH(0);
  else if (w.S <= 8)
    PRGPU_RB_LAUNCH(8);
  else if (w.S <= 16)
    PRGPU_RB_LAUNCH(16);
  else if (w.S <= 24)
    PRGPU_RB_LAUNCH(24);
  else
    PRGPU_RB_LAUNCH(32);
#undef PRGPU_RB_LAUNCH
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

} // namespace prgpu

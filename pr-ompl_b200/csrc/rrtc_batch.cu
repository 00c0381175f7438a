// T1: many independent RRTConnect queries, resident on the device from the first sample to the solution.
//
// Follows pr_ompl::RRTConnect::solve / growTree (pr_ompl/src/RRTConnect.cpp:219-376, :178-217)
// statement by statement -- tree alternation before the TRAPPED `continue` (:267 vs :303), goal
// seeding (:270-286), CONNECT as repeated growTree with a fresh nearest query each step (:296-300),
// rstate overwritten by the last added state (:313-314), goal-tree edges validated from the new
// state toward the tree (:198), one step back on the start side at the connection (:332-335).
//
// A query is a long dependent chain (sample -> nearest -> validate -> append): its speed is set by
// latency, not width.  Two kernels:
//
//  rrtc_batch_kernel  one WARP per query.  The 32 lanes cooperate inside each growTree: the linear
//      nearest scan with the (distance, index) order of NearestNeighborsLinear (four independent fp64
//      chains per lane, the square root only for a node that may replace the best, butterfly reduction),
//      then the DiscreteMotionValidator tested set spread over the lanes with a vote for the early
//      exit.  No block barrier.  Warps are persistent: each takes the next unstarted query from a
//      global counter; the obstacle tables in shared memory serve the RB_WARPS queries of a block.
//      A query still unsolved after RB_CUT samples is suspended (its trees are in global memory, the
//      loop state is a few words) and queued for the second kernel.
//
//  rrtc_team_kernel   one BLOCK (RB_WARPS warps) per suspended query.  These are the queries that mostly
//      get TRAPPED: an iteration whose first growTree is trapped changes nothing but the sample counter
//      and the tree alternation, so warp w evaluates the first growTree of iteration it+w against the
//      unchanged trees; the iterations before the first untrapped one are skipped wholesale, the warp
//      that found it carries its iteration on (append, the other tree's CONNECT, the connection test).
//      Meanwhile the warps without a check for their next iteration make it against the trees as they
//      were; it is kept iff no appended node is strictly nearer to its sample than the nearest it found.
//      Same samples, same order of effects, bit-identical trees.
#include <cstdio>
#include "rrtc_batch.cuh"
#include "sample_stream.cuh"

namespace prgpu {
namespace {

#ifndef RB_WARPS_N
#define RB_WARPS_N 8
#endif
constexpr int RB_WARPS = RB_WARPS_N;   // queries in flight per block (first kernel) / team size (second)
constexpr int RB_THREADS = RB_WARPS * 32;
#ifndef RB_CUT
#define RB_CUT 256                     // samples after which an unsolved query moves to the team kernel
#endif
enum { GS_TRAPPED = 0, GS_ADVANCED = 1, GS_REACHED = 2 };
enum { ST_INVALID_START = 1, ST_INVALID_GOAL = 2, ST_TIMEOUT = 4, ST_EXACT = 6 }; // PlannerStatus values
enum { IT_CONTINUE = 0, IT_END = 1 };

// dynamic shared memory: the obstacle tables, then per warp the centres and the work list of a validity round
__host__ __device__ __forceinline__ size_t warp_scratch_offset(const DeviceWorld &w) {
  return (obstacle_tables_bytes(w) + 15) & ~(size_t)15;
}
__host__ __device__ __forceinline__ size_t warp_scratch_bytes(const DeviceWorld &w) {
  return ((size_t)32 * w.S * (3 * sizeof(float) + sizeof(uint16_t)) + 15) & ~(size_t)15;
}

struct Shared { // per warp
  double rstate[WORLD_MAX_DOF];
  double xstate[WORLD_MAX_DOF];
  double near[WORLD_MAX_DOF];
};

// the tested set of one edge (DiscreteMotionValidator: `to` first, then `from` if asked, then the interior)
struct EdgeStates {
  const double *from, *to;
  uint32_t nd, extra, rounds, rd;
  __device__ __forceinline__ void state(uint32_t pt, double *q) const {
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
  }
  // state held by lane `st` in round `rd`
  __device__ __forceinline__ void get(uint32_t st, double *q) const { state(rd + st * rounds, q); }
};

// One validity round of the warp: lane s holds state s (`have`, `q`).  A state test is a latency chain
// (FK, then 16 field loads, then obstacle lists for the spheres the field cannot certify) and only ~20
// lanes have a state, so the round is split by KIND of work instead of by state:
//   A  each lane runs the fp32 FK chain of its state and leaves the sphere centres in shared memory;
//   B  the (state, sphere) pairs are dealt to all 32 lanes, four field loads in flight per lane;
//   C  the uncertified pairs are dealt to all lanes for the obstacle lists, a vote after every 32.
// `es->get(s, q)` regenerates state s (the fp64 verdict of a thin-band pair needs it).  Requires the
// clearance field; identical verdicts to state_valid_grid by construction (same functions).
static __device__ __noinline__ bool round_valid(const DeviceWorld *wp, const ObstacleTables *tabs, float *cen, uint16_t *work,
                                              int tid, bool have, const double *q, const EdgeStates *es) {
  const DeviceWorld &w = *wp;
  const ArmTableF32 &arm = *tabs->arm;
  const int S = arm.link_first[w.dof + 1];
  const unsigned hm = __ballot_sync(0xffffffffu, have);
  const int nst = __popc(hm); // the states sit in lanes 0 .. nst-1
  if (have) {
    FrameF32 f;
    f.reset();
    for (int i = 0; i <= w.dof; ++i) {
      if (i > 0)
        f.advance(arm, q, i);
      for (int j = arm.link_first[i]; j < arm.link_first[i + 1]; ++j) {
        float cx, cy, cz;
        f.centre(arm.sphere[j], cx, cy, cz);
        float *c = cen + ((size_t)tid * S + j) * 3;
        c[0] = cx;
        c[1] = cy;
        c[2] = cz;
      }
    }
  }
  __syncwarp();
  const int nitems = nst * S;
  const int dnx = w.df_n[0], dny = w.df_n[1], dnz = w.df_n[2];
  const float dox = w.df_org[0], doy = w.df_org[1], doz = w.df_org[2], dinv = w.df_inv;
  int nwork = 0;
  const bool pow2 = (S & (S - 1)) == 0;
  const int sh = 31 - __clz(S);
  const float dout = w.df_out;
  const uint32_t *df = w.df;
  const float two_delta = 2.0f * w.delta;
  bool hit = false; // some sphere certainly penetrates an obstacle (the field's upper bound)
  constexpr int RV_U = 8; // field loads in flight per lane
  for (int base = 0; base < nitems; base += 32 * RV_U) {
    float dv[RV_U];
#pragma unroll
    for (int u = 0; u < RV_U; ++u) {
      const int it = base + u * 32 + tid;
      dv[u] = CUDART_INF_F;
      if (it < nitems) {
        const float *c = cen + (size_t)it * 3;
        const float fx = (c[0] - dox) * dinv, fy = (c[1] - doy) * dinv, fz = (c[2] - doz) * dinv;
        const bool in = fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)dnx && fy < (float)dny && fz < (float)dnz;
        // (outside the field's box, or NaN: the margin)
        const float2 d = in ? df_unpack(__ldg(df + ((size_t)(int)fz * dny + (int)fy) * dnx + (int)fx))
                            : make_float2(dout, CUDART_INF_F);
        const float rs = arm.sphere[pow2 ? (it & (S - 1)) : (it % S)].w;
        dv[u] = d.x - rs;
        hit |= d.y < rs - two_delta;
      }
    }
    if (__any_sync(0xffffffffu, hit))
      return false;
#pragma unroll
    for (int u = 0; u < RV_U; ++u) {
      const int it = base + u * 32 + tid;
      const bool unc = it < nitems && !(dv[u] > 0.f);
      const unsigned bal = __ballot_sync(0xffffffffu, unc);
      if (unc)
        work[nwork + __popc(bal & ((1u << tid) - 1u))] = (uint16_t)it;
      nwork += __popc(bal);
    }
  }
  __syncwarp();
  for (int b0 = 0; b0 < nwork; b0 += 32) {
    bool bad = false;
    if (b0 + tid < nwork) {
      const int it = work[b0 + tid];
      const int st = pow2 ? (it >> sh) : (it / S), j = pow2 ? (it & (S - 1)) : (it % S);
      double qs[WORLD_MAX_DOF];
      const float *c = cen + (size_t)it * 3;
      bad = !grid_sphere_free(
          w, *tabs,
          [&]() {
            es->get((uint32_t)st, qs); // only a thin-band pair gets here
            return (const double *)qs;
          },
          j, c[0], c[1], c[2]);
    }
    if (__any_sync(0xffffffffu, bad))
      return false;
  }
  return true;
}


// One query as seen by one warp.  Every member is warp-uniform.
template <int SP>
struct Planner {
  const DeviceWorld *wp;
  const ObstacleTables *tabs;
  const RrtcDeviceParams &p;
  Shared &sh;
  int tid; // lane
  int dim;
  uint32_t qid;
  double *tS, *tG;
  int32_t *pS, *pG;
  const double *goal;
  float *cen;     // per warp: sphere centres of the states of a validity round, [32][S][3]
  uint16_t *work; // per warp: (state, sphere slot) items the clearance field could not certify, [32 * S]
  RrtcLoopState s;

  struct Check { // result of the first half of growTree
    int32_t nmotion;
    bool ok, reach;
    double bd;               // distance of the nearest node
    double c[WORLD_MAX_DOF]; // dstate
  };

  __device__ __forceinline__ Planner(const DeviceWorld *wp_, const ObstacleTables *tabs_, const RrtcDeviceParams &p_,
                                     const RrtcDeviceBuffers &b, Shared &sh_, int lane, uint32_t qid_,
                                     unsigned char *warp_scratch)
      : wp(wp_), tabs(tabs_), p(p_), sh(sh_), tid(lane), dim(p_.dim), qid(qid_) {
    cen = reinterpret_cast<float *>(warp_scratch);
    work = reinterpret_cast<uint16_t *>(cen + 32 * wp_->S * 3);
    tS = b.start_states + (size_t)qid * p.max_nodes * dim;
    tG = b.goal_states + (size_t)qid * p.max_nodes * dim;
    pS = b.start_parent + (size_t)qid * p.max_nodes;
    pG = b.goal_parent + (size_t)qid * p.max_nodes;
    goal = b.goals + (size_t)qid * dim;
  }

  __device__ __forceinline__ bool in_bounds(const double *q) const { // RealVectorStateSpace::satisfiesBounds
    bool ok = true;
    for (int i = 0; i < dim; ++i)
      if (q[i] - 2.220446049250313e-16 > p.hi[i] || q[i] + 2.220446049250313e-16 < p.lo[i])
        ok = false;
    return ok;
  }
  __device__ __forceinline__ bool single_valid(const double *q) { // isValid of one state, evaluated by lane 0
    int f = 0;
    if (tid == 0) {
      double qq[WORLD_MAX_DOF];
#pragma unroll
      for (int i = 0; i < WORLD_MAX_DOF; ++i)
        qq[i] = i < dim ? q[i] : 0.0;
      f = (in_bounds(qq) && state_valid_call<SP>(wp, tabs, qq)) ? 1 : 0;
    }
    return __shfl_sync(0xffffffffu, f, 0) != 0;
  }
  __device__ __forceinline__ void append(double *T, int32_t *P, uint32_t &n, const double *q, int32_t parent) {
    if (tid < dim)
      T[(size_t)n * dim + tid] = q[tid];
    if (tid == 0)
      P[n] = parent;
    ++n;
    __syncwarp(); // the new node is visible to the next nearest scan
  }
  __device__ __forceinline__ void draw_sample(uint32_t it) { // :289
    if (tid < dim)
      sh.rstate[tid] = stream_uniform(p.seed, p.query_id_base + qid, it, tid, p.lo[tid], p.hi[tid]);
    __syncwarp();
  }

  // growTree (RRTConnect.cpp:178-217), first half: nearest (:181), step clamp (:188-194), motion check (:198)
  __device__ __forceinline__ void check(const double *T, uint32_t n, bool tgi_start, Check &r) {
    // The scan is part of a query's latency chain: four nodes per lane are summed as independent chains,
    // and the square root -- monotone, so a node whose squared sum is not below the best one cannot
    // have a smaller distance -- is taken only for a node that may replace the best.
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
    r.nmotion = bi; // the butterfly leaves the same (distance, index) minimum in every lane
    r.bd = bd;
    if (tid < dim)
      sh.near[tid] = T[(size_t)r.nmotion * dim + tid];
    __syncwarp();

    // :188-194 distance and step clamp (bit exact; every lane computes the same values)
    double a[WORLD_MAX_DOF]; // nmotion->state
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < WORLD_MAX_DOF; ++k) {
      a[k] = k < dim ? sh.near[k] : 0.0;
      r.c[k] = rq[k];
      const double diff = __dsub_rn(a[k], r.c[k]);
      s = __dadd_rn(s, __dmul_rn(diff, diff));
    }
    const double d = __dsqrt_rn(s);
    r.reach = true;
    if (d > p.range) {
      const double t = p.range / d;
#pragma unroll
      for (int k = 0; k < WORLD_MAX_DOF; ++k)
        r.c[k] = k < dim ? lerp_exact(a[k], r.c[k], t) : 0.0;
      r.reach = false;
    }
    // :198 start tree: checkMotion(nmotion, dstate); goal tree: isValid(dstate) && checkMotion(dstate, nmotion)
    // DiscreteMotionValidator: s1 = from, s2 = to; tested set {s2} U {from + (to-from) j/nd} (+ dstate itself)
    const double *from = tgi_start ? a : r.c;
    const double *to = tgi_start ? r.c : a;
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
    r.ok = true;
    EdgeStates es;
    es.from = from;
    es.to = to;
    es.nd = nd;
    es.extra = extra;
    es.rounds = rounds;
    for (uint32_t rd = 0; rd < rounds; ++rd) {
      const uint32_t pt = rd + (uint32_t)tid * rounds;
      double q[WORLD_MAX_DOF];
      const bool have = pt < np;
      es.rd = rd;
      if (have)
        es.state(pt, q);
      bool all_good;
      if (SP == 0 && wp->df) {
        all_good = round_valid(wp, tabs, cen, work, tid, have, q, &es);
      } else {
        const bool good = have ? state_valid_call<SP>(wp, tabs, q) : true;
        all_good = !__any_sync(0xffffffffu, !good);
      }
      if (!all_good) {
        r.ok = false;
        break;
      }
    }
  }
  // A check made against n0 nodes stays the check against n1 > n0 nodes unless one of the new nodes is
  // strictly nearer to the sample (sh.rstate) than the nearest it found (an equal distance loses to the
  // lower index).  Same arithmetic as the scan.
  __device__ __forceinline__ bool still_nearest(const double *T, uint32_t n0, uint32_t n1, const Check &r) {
    bool nearer = false;
    for (uint32_t i = n0 + tid; i < n1; i += 32) {
      const double *x = T + (size_t)i * dim;
      double sq = 0.0;
#pragma unroll
      for (int k = 0; k < WORLD_MAX_DOF; ++k)
        if (k < dim) {
          const double diff = __dsub_rn(x[k], sh.rstate[k]);
          sq = __dadd_rn(sq, __dmul_rn(diff, diff));
        }
      nearer = nearer || __dsqrt_rn(sq) < r.bd;
    }
    return !__any_sync(0xffffffffu, nearer);
  }
  // growTree, second half (:199-216)
  __device__ __forceinline__ int commit(double *T, int32_t *P, uint32_t &n, const Check &r, int32_t &xmotion) {
    if (!r.ok)
      return GS_TRAPPED; // :215
    if (n >= p.max_nodes) {
      s.full = 1;
      return GS_TRAPPED;
    }
    if (tid < dim)
      sh.xstate[tid] = r.c[tid];
    __syncwarp();
    xmotion = (int32_t)n;
    append(T, P, n, sh.xstate, r.nmotion); // :200-209
    return r.reach ? GS_REACHED : GS_ADVANCED;
  }

  // :225-257 up to the loop
  __device__ __forceinline__ void begin(const double *start) {
    s.it = 0;
    s.nS = s.nG = 0;
    s.sampledGoals = 0;
    s.connS = s.connG = -1;
    s.status = ST_TIMEOUT;
    s.startTree = 1;
    s.full = 0;
    if (single_valid(start)) { // :230-236 via PlannerInputStates::nextStart
      if (tid < dim)
        sh.xstate[tid] = start[tid];
      __syncwarp();
      append(tS, pS, s.nS, sh.xstate, -1);
    }
    if (s.nS == 0)
      s.status = ST_INVALID_START; // :238-242
  }

  // one pass of the solve loop (:263-376).  `pre`: the first growTree's check, already done by this warp
  // against the current trees for this iteration's sample (team kernel), or nullptr.
  __device__ __forceinline__ int iteration(const Check *pre) {
    const bool growStart = s.startTree != 0; // tree = startTree ? tStart : tGoal
    bool tgi_start = growStart;
    s.startTree = s.startTree ? 0u : 1u; // :267
    if (s.nG == 0 || s.sampledGoals < s.nG / 2) { // :270
      if (s.sampledGoals < 1) { // a GoalState yields one sample ever (Appendix B.7)
        s.sampledGoals++;
        if (single_valid(goal)) {
          if (tid < dim)
            sh.xstate[tid] = goal[tid];
          __syncwarp();
          append(tG, pG, s.nG, sh.xstate, -1);
        }
      }
      if (s.nG == 0) // :281-285
        return IT_END;
    }
    draw_sample(s.it);
    ++s.it;

    int gs = GS_TRAPPED;
    int32_t xmotion = -1, added = -1; // :295
    bool second_done = false;
    for (int phase = 0; phase < 2; ++phase) {
      // phase 0 grows `tree` toward the sample (:296-300), phase 1 grows `otherTree` toward the
      // node just added (:316-321)
      const bool onStart = (phase == 0) ? growStart : !growStart;
      const int ext = (phase == 0) ? p.ext_first : p.ext_second;
      do {
        double *T = onStart ? tS : tG;
        int32_t *P = onStart ? pS : pG;
        uint32_t &n = onStart ? s.nS : s.nG;
        if (pre) {
          gs = commit(T, P, n, *pre, xmotion);
          pre = nullptr;
        } else {
          Check r;
          check(T, n, tgi_start, r);
          gs = commit(T, P, n, r, xmotion);
        }
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
      tgi_start = s.startTree != 0; // :316
    }
    if (!second_done)
      return IT_CONTINUE;
    if (gs == GS_REACHED) { // :327 (isStartGoalPairValid is true for a GoalState)
      int32_t startMotion = s.startTree ? xmotion : added; // :323-324
      int32_t goalMotion = s.startTree ? added : xmotion;
      if (pS[startMotion] >= 0) // :332-335
        startMotion = pS[startMotion];
      else
        goalMotion = pG[goalMotion];
      s.connS = startMotion;
      s.connG = goalMotion;
      s.status = ST_EXACT;
      return IT_END;
    }
    return IT_CONTINUE;
  }

  __device__ __forceinline__ void finish(const RrtcDeviceBuffers &b) const {
    if (tid == 0) {
      b.status[qid] = s.status;
      b.iters[qid] = s.it;
      b.n_start[qid] = s.nS;
      b.n_goal[qid] = s.nG;
      b.conn_start[qid] = s.connS;
      b.conn_goal[qid] = s.connG;
      b.overflow[qid] = s.full ? 1u : 0u;
    }
  }
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
  for (;;) { // persistent warp: next unstarted query
    uint32_t qid = 0;
    if (lane == 0)
      qid = atomicAdd(b.work_counter, 1u);
    qid = __shfl_sync(0xffffffffu, qid, 0);
    if (qid >= n_queries)
      return;
    Planner<SP> pl(wp, &tabs, p, b, sh_all[threadIdx.x >> 5], lane, qid,
                   reinterpret_cast<unsigned char *>(s_obs) + warp_scratch_offset(w) +
                       (size_t)(threadIdx.x >> 5) * warp_scratch_bytes(w));
    pl.begin(b.starts + (size_t)qid * p.dim);
    bool ended = pl.s.status == ST_INVALID_START;
    while (!ended && pl.s.it < p.max_iters && !pl.s.full && pl.s.it < (uint32_t)RB_CUT) // :263, ptc = sample budget
      ended = pl.iteration(nullptr) == IT_END;
    if (!ended && pl.s.it < p.max_iters && !pl.s.full) {
      // suspended: the team kernel carries on from here
      if (lane == 0) {
        b.resume[qid] = pl.s;
        b.susp_list[atomicAdd(b.susp_count, 1u)] = qid;
      }
    } else {
      pl.finish(b);
    }
    __syncwarp();
  }
}

template <int SP>
__global__ void __launch_bounds__(RB_THREADS)
rrtc_team_kernel(const DeviceWorld *__restrict__ wp, const RrtcDeviceParams p, const RrtcDeviceBuffers b) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  __shared__ Shared sh_all[RB_WARPS];
  __shared__ RrtcLoopState team_state;
  __shared__ int team_flag[RB_WARPS]; // by window slot; 0: first growTree not trapped, 1: trapped, 2: past the
                                      // sample budget, 3: not evaluated yet
  __shared__ uint32_t team_item;
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n_susp = *b.susp_count;
  for (;;) { // persistent block: next suspended query
    if (threadIdx.x == 0)
      team_item = atomicAdd(b.team_counter, 1u);
    __syncthreads();
    const uint32_t item = team_item;
    if (item >= n_susp)
      return;
    const uint32_t qid = b.susp_list[item];
    Planner<SP> pl(wp, &tabs, p, b, sh_all[warp], lane, qid,
                   reinterpret_cast<unsigned char *>(s_obs) + warp_scratch_offset(w) + (size_t)warp * warp_scratch_bytes(w));
    pl.s = b.resume[qid];
    // a suspended query has drawn its goal sample and has a goal tree: the seeding block of
    // `iteration` is a no-op from here on.
    // Warp w owns iteration my_it (the eight warps own eight consecutive iterations, the window); `have`:
    // r is the first growTree's check of iteration my_it against the trees as they are now.
    uint32_t my_it = pl.s.it + (uint32_t)warp;
    bool have = false;
    typename Planner<SP>::Check r;
    r.ok = false;
#ifdef RB_ACCOUNT
    long long pr_rounds = 0, pr_step1 = 0, pr_lazy = 0, pr_inval = 0, pr_look = 0, pr_chk = 0, pr_exec = 0, pr_t0 = clock64();
    uint32_t pr_it0 = pl.s.it;
#endif
    auto speculate = [&](uint32_t base_it, uint32_t base_start_tree) { // check of iteration my_it, trees as they are
      const bool growStart = (base_start_tree != 0) != (((my_it - base_it) & 1u) != 0);
      pl.draw_sample(my_it);
      pl.check(growStart ? pl.tS : pl.tG, growStart ? pl.s.nS : pl.s.nG, growStart, r);
      have = true;
    };
    for (;;) {
      if (!(pl.s.it < p.max_iters) || pl.s.full)
        break;
      const uint32_t slot = my_it - pl.s.it; // 0 .. RB_WARPS-1
#ifdef RB_ACCOUNT
      long long pc0 = clock64();
      ++pr_rounds;
#endif
      // (1) the checks of the window.  The last slot is evaluated only if the others are all trapped: it
      //     belongs to the warp that carried the previous untrapped iteration on and could not look ahead
      if (!have && my_it < p.max_iters && slot + 1 < RB_WARPS) {
        speculate(pl.s.it, pl.s.startTree);
#ifdef RB_ACCOUNT
        ++pr_step1;
#endif
      }
      if (lane == 0)
        team_flag[slot] = my_it >= p.max_iters ? 2 : (have ? (r.ok ? 0 : 1) : 3);
      __syncthreads();
      int first = RB_WARPS;
      for (int k = RB_WARPS - 1; k >= 0; --k)
        if (team_flag[k] != 1)
          first = k;
      int first_flag = first < RB_WARPS ? team_flag[first] : 1;
      const bool lazy = first_flag == 3; // uniform
      __syncthreads(); // everyone has read the flags
      if (lazy) {
#ifdef RB_ACCOUNT
        ++pr_lazy;
#endif
        if ((int)slot == first) {
          speculate(pl.s.it, pl.s.startTree);
          if (lane == 0)
            team_flag[slot] = r.ok ? 0 : 1;
        }
        __syncthreads();
        first_flag = team_flag[first];
        if (first_flag == 1)
          first = RB_WARPS; // it was the last slot: the whole window is trapped
        __syncthreads();
      }
#ifdef RB_ACCOUNT
      pr_chk += clock64() - pc0;
      long long pe0 = clock64();
#endif
      // (2) the `first` iterations before it were trapped: they moved the sample counter and the alternation
      const bool run = first < RB_WARPS && first_flag == 0;
      const uint32_t old_nS = pl.s.nS, old_nG = pl.s.nG;
      pl.s.it += (uint32_t)first;
      if (first & 1)
        pl.s.startTree = pl.s.startTree ? 0u : 1u;
      const bool executor = run && (int)slot == first;
      const bool freed = (int)slot < first || executor; // my iteration is consumed: take the next free one
      bool ended = false;
      if (executor) {
        ended = pl.iteration(&r) == IT_END;
        if (lane == 0) {
          team_state = pl.s;
          team_flag[0] = ended ? 1 : 0;
        }
        my_it += RB_WARPS;
        have = false;
      } else {
        if (freed) {
          my_it += RB_WARPS;
          have = false;
        }
        // (3) look ahead while the executor works: whoever has no check for its iteration (a freed warp, or
        //     the owner of the last slot whose check was not needed) makes it against the trees as they
        //     were; validated below against what the executor appends
        if (run && !have && my_it < p.max_iters) {
          speculate(pl.s.it, pl.s.startTree);
#ifdef RB_ACCOUNT
          ++pr_look;
#endif
        }
      }
      if (run) {
        __syncthreads();
        const uint32_t cur_it = pl.s.it; // iteration the executor carried on
        const uint32_t cur_start = pl.s.startTree;
        pl.s = team_state;
        ended = team_flag[0] != 0;
        if (!ended && have) {
          // the tree my check was made against: parity of my_it relative to the executed iteration
          const bool growStart = (cur_start != 0) != (((my_it - cur_it) & 1u) != 0);
          const bool okn = growStart ? pl.still_nearest(pl.tS, old_nS, pl.s.nS, r)
                                     : pl.still_nearest(pl.tG, old_nG, pl.s.nG, r);
          if (!okn) {
            have = false;
#ifdef RB_ACCOUNT
            ++pr_inval;
#endif
          }
        }
      }
      __syncthreads(); // team_flag / team_state are rewritten by the next round
#ifdef RB_ACCOUNT
      pr_exec += clock64() - pe0;
#endif
      if (ended)
        break;
    }
#ifdef RB_ACCOUNT
    if (lane == 0 && pl.s.it >= p.max_iters && (qid % 29) == 0)
      printf("TEAM q %u w %d its %u rounds %lld step1 %lld lazy %lld look %lld inval %lld chk %lld exec %lld total %lld\n", qid, warp,
             pl.s.it - pr_it0, pr_rounds, pr_step1, pr_lazy, pr_look, pr_inval, pr_chk, pr_exec, clock64() - pr_t0);
#endif
    if (warp == 0)
      pl.finish(b);
    __syncthreads();
  }
}

} // namespace

int rrtc_batch_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const RrtcDeviceParams &p,
                      const RrtcDeviceBuffers &b, uint32_t n_queries, int sm_count, cudaStream_t stream) {
  if (n_queries == 0)
    return PRGPU_OK;
  const size_t smem = warp_scratch_offset(w) + RB_WARPS * warp_scratch_bytes(w);
  // work_counter, susp_count, team_counter are consecutive words
  PRGPU_CUDA_TRY(cudaMemsetAsync(b.work_counter, 0, 3 * sizeof(uint32_t), stream));
#define PRGPU_RB_LAUNCH(SPV)                                                                       \
  do {                                                                                             \
    auto kern = rrtc_batch_kernel<SPV>;                                                            \
    auto team = rrtc_team_kernel<SPV>;                                                             \
    PRGPU_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    PRGPU_CUDA_TRY(cudaFuncSetAttribute(team, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    int per_sm = 0, per_sm_t = 0;                                                                  \
    PRGPU_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RB_THREADS, smem)); \
    PRGPU_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, team, RB_THREADS, smem)); \
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_queries + RB_WARPS - 1) / RB_WARPS,    \
                                                         (uint64_t)std::max(per_sm, 1) * sm_count); \
    kern<<<blocks, RB_THREADS, smem, stream>>>(w_dev, p, b, n_queries); PRGPU_COUNT_LAUNCH(1);                            \
    const unsigned tblocks =                                                                       \
        (unsigned)std::min<uint64_t>(n_queries, (uint64_t)std::max(per_sm_t, 1) * sm_count);       \
    team<<<tblocks, RB_THREADS, smem, stream>>>(w_dev, p, b); PRGPU_COUNT_LAUNCH(1);                                      \
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

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
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <algorithm>
#include <cstdlib>
#include "check_motion.cuh"
#include <map>
#include <mutex>
#include <utility>
#include "edge_plan.cuh"

namespace prgpu {
namespace {

constexpr int CM_THREADS = 128;

// cudaFuncSetAttribute costs microseconds; a planner's single-edge call is a few of them long.  The opt-in is
// remembered per kernel and device and raised only when a larger world asks for more.
template <class K> cudaError_t set_smem_once(K kern, size_t smem) {
  static std::map<std::pair<const void *, int>, size_t> have; // (kernels of one signature share this instantiation)
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  size_t &h = have[std::make_pair(reinterpret_cast<const void *>(kern), dev)];
  if (smem <= h)
    return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess)
    h = smem;
  return e;
}
#ifndef CM_MINB
#define CM_MINB 8 // resident blocks per SM the two FK-heavy kernels of the pair-dealt pipeline are compiled for
#endif

template <int SP>
__global__ void __launch_bounds__(CM_THREADS)
check_motion_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ from, const double *__restrict__ to,
                    uint64_t n, double resolution, int mode, int check_first,
                    uint8_t *__restrict__ valid) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs, n > 64); // (a planner's single edge: tables read in place)
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
        good = state_valid_any<SP>(wp, &tabs, q);
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

// ---- large batches: two flattened phases, one thread per tested state --------------------------
// The warp-per-edge kernel above leaves lanes idle (an edge's state count is rarely a multiple of 32)
// and can only stop an edge between rounds.  For large batches the tested sets are flattened instead:
//   phase A: up to 8 states per edge, spread evenly over its tested set (stride m = ceil(np/8)), 8
//            lanes per edge; an edge with an invalid sample is finished, and so is one with np <= 8.
//            Three kernels, so that every lane of each does the same kind of work: `sample` tests the
//            samples against the clearance field only, `resolve` takes the samples the field could not
//            certify through the obstacle lists, `gaps` sizes what is left.
//            A certified sample also has a lower bound C of its clearance (clearance field, world.cuh);
//            no link-sphere centre moves by more than travel = sum_i |to_i - from_i| rho_i / nd between
//            neighbouring tested states, so the floor(C / travel) states on either side of the sample
//            are free WITHOUT being tested (conservative advancement): they are struck off the gap
//            between this sample and the next;
//   scan   : exclusive prefix sum of the states still owed by undecided edges;
//   expand : owed item -> (edge, tested-state index);
//   phase B: one thread per owed state; a thread first looks at the edge's verdict and skips the state
//            if the edge is already invalid.
// Every lane of both phases carries a state, and the result is the AND over the reference's tested set
// (the struck-off states are proven free, not assumed).

// phase A, step 1: the 8 samples of every edge against the clearance field only.  A certified sample
// leaves its reach (the number of neighbouring states its clearance covers); the others are queued with
// the mask of their uncertified spheres.  Every lane does the same work: no obstacle lists here.
template <int SP>
__global__ void __launch_bounds__(CM_THREADS)
cm_sample_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ from, const double *__restrict__ to,
                 uint32_t n, double resolution, int mode, int check_first, uint8_t *__restrict__ valid,
                 uint32_t *__restrict__ np_out, uint2 *__restrict__ gaps, uint2 *__restrict__ queue,
                 uint32_t *__restrict__ queue_n) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const int dof = w.dof;
  const int lane = threadIdx.x & 31;
  const uint32_t groups_total = gridDim.x * (CM_THREADS / 8);
  for (uint32_t e0 = blockIdx.x * (CM_THREADS / 8); e0 < n; e0 += groups_total) {
    const uint32_t e = e0 + (threadIdx.x >> 3);
    const uint32_t slot = threadIdx.x & 7;
    uint32_t reach = 0, unc = 0;
    if (e < n) {
      double a[WORLD_MAX_DOF], b[WORLD_MAX_DOF];
#pragma unroll
      for (int i = 0; i < WORLD_MAX_DOF; ++i) {
        a[i] = i < dof ? from[(size_t)e * dof + i] : 0.0;
        b[i] = i < dof ? to[(size_t)e * dof + i] : 0.0;
      }
      const EdgePlan pl = plan_edge(a, b, resolution, mode, check_first);
      if (slot == 0) {
        np_out[e] = pl.np;
        valid[e] = 1;
      }
      const uint32_t p = slot * pl.stride;
      if (p < pl.np) {
        double q[WORLD_MAX_DOF];
        edge_state(pl, a, b, mode, p, q);
        // no link-sphere centre moves by more than `travel` between neighbouring tested states
        float travel = 0.f;
#pragma unroll
        for (int i = 0; i < WORLD_MAX_DOF; ++i)
          travel += (float)fabs(b[i] - a[i]) * (i < dof ? w.rho[i] : 0.f);
        travel *= (float)(mode == PRGPU_MODE_OMPL_DISCRETE ? 1.0 / (double)pl.nd : pl.step) * (1.0f + 1e-5f);
        unc = 0xFFFFFFFFu;
        if constexpr (SP == 0) {
          if (w.df) {
            float clear = 0.f;
            unc = state_certify_grid(wp, &tabs, q, &clear);
            if (unc == 0 && clear > 0.f) {
              const float r = travel > 0.f ? floorf(clear / travel * (1.0f - 1e-5f)) : (float)pl.np;
              reach = r >= (float)pl.np ? pl.np : (uint32_t)r;
            }
          }
        }
      }
      gaps[(size_t)e * 8 + slot] = make_uint2(reach, 0u);
    }
    // queue the uncertified samples (one atomic per warp)
    const unsigned bal = __ballot_sync(0xffffffffu, unc != 0);
    if (bal) {
      const int leader = __ffs(bal) - 1;
      uint32_t base = 0;
      if (lane == leader)
        base = atomicAdd(queue_n, (uint32_t)__popc(bal));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (unc != 0)
        queue[base + __popc(bal & ((1u << lane) - 1u))] = make_uint2(e * 8u + slot, unc);
    }
  }
}

// phase A, step 2: the queued samples walk the obstacle lists of their uncertified spheres (worlds without a
// clearance field: the whole test).  An invalid sample finishes its edge.
template <int SP>
__global__ void __launch_bounds__(CM_THREADS)
cm_resolve_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ from, const double *__restrict__ to,
                  double resolution, int mode, int check_first, volatile uint8_t *valid,
                  const uint2 *__restrict__ queue, const uint32_t *__restrict__ queue_n) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const int dof = w.dof;
  const uint32_t total = *queue_n;
  for (uint32_t g = blockIdx.x * CM_THREADS + threadIdx.x; g < total; g += gridDim.x * CM_THREADS) {
    const uint2 item = queue[g];
    const uint32_t e = item.x >> 3, slot = item.x & 7u;
    if (valid[e] == 0)
      continue;
    double a[WORLD_MAX_DOF], b[WORLD_MAX_DOF];
#pragma unroll
    for (int i = 0; i < WORLD_MAX_DOF; ++i) {
      a[i] = i < dof ? from[(size_t)e * dof + i] : 0.0;
      b[i] = i < dof ? to[(size_t)e * dof + i] : 0.0;
    }
    const EdgePlan pl = plan_edge(a, b, resolution, mode, check_first);
    double q[WORLD_MAX_DOF];
    edge_state(pl, a, b, mode, slot * pl.stride, q);
    bool good;
    if constexpr (SP == 0)
      good = w.df ? state_lists_grid(wp, &tabs, q, item.y) : state_valid_grid(wp, &tabs, q);
    else
      good = state_valid_any<SP>(wp, &tabs, q);
    if (!good)
      valid[e] = 0;
  }
}

// phase A, step 3: what the samples leave untested.  Lane `slot` of an edge owns the tested states
// strictly between its sample and the next one (or the end); the reaches of the two samples strike
// states off both ends of that gap.
__global__ void __launch_bounds__(CM_THREADS)
cm_gaps_kernel(uint32_t n, int mode, const uint8_t *__restrict__ valid, uint32_t *__restrict__ owed,
               uint2 *__restrict__ gaps) {
  const uint32_t idx = blockIdx.x * CM_THREADS + threadIdx.x;
  const uint32_t e = idx >> 3, slot = idx & 7u;
  const bool in = e < n;
  const uint32_t np = in ? owed[e] : 0u; // the sample kernel left the tested-set size here
  const uint32_t reach = in ? gaps[idx].x : 0u;
  const bool ok = in && valid[e] != 0;
  __syncwarp();
  const uint32_t stride = max(1u, (np + 7u) / 8u);
  const int gbase = (threadIdx.x & 31) & ~7;
  const uint32_t reach_next = __shfl_sync(0xffffffffu, reach, gbase + ((slot + 1) & 7));
  const uint32_t reach_first = __shfl_sync(0xffffffffu, reach, gbase);
  uint32_t cnt = 0, lo = 0;
  const uint32_t p = slot * stride;
  if (ok && p < np && stride > 1) {
    const uint32_t pe = min(p + stride, np); // next sample or one past the last state
    const bool last = pe == np;               // no sample at pe
    uint32_t left = reach, right;
    if (mode == PRGPU_MODE_OMPL_DISCRETE) {
      // p = 0 is the END state `to`; the states after it in p order start at `from`
      if (slot == 0)
        left = 0;
      right = last ? reach_first : reach_next; // the trailing states run up to `to` (sample 0)
    } else {
      right = last ? 0u : reach_next;
    }
    const uint64_t l = (uint64_t)p + 1 + left, h = (uint64_t)pe - 1; // h inclusive before the right cut
    if (h >= l && h - l + 1 > (uint64_t)right) {
      lo = (uint32_t)l;
      cnt = (uint32_t)(h - l + 1 - right);
    }
  }
  if (in)
    gaps[idx] = make_uint2(lo, cnt);
  uint32_t tot = cnt;
  tot += __shfl_xor_sync(0xffffffffu, tot, 1);
  tot += __shfl_xor_sync(0xffffffffu, tot, 2);
  tot += __shfl_xor_sync(0xffffffffu, tot, 4);
  if (in && slot == 0)
    owed[e] = ok ? tot : 0u;
}

// ---- the same phases with the divergent work dealt out by (state, link sphere) PAIRS ------------------------
// (worlds with a clearance field; r01 ncu: the per-sample list walk above ran at 10 of 32 lanes per instruction
// because one thread re-ran the FK chain and then walked the lists of ITS uncertified spheres)
//   sample2  : as cm_sample_kernel, but (a) an uncertified sample leaves one ITEM per uncertified sphere --
//              (state reference, sphere slot, fp32 centre) -- so the list walk needs no FK and one thread owns
//              one cell list; (b) the 8 reaches of an edge sit in one warp: the gaps between the samples are
//              struck off by shuffles right here (no gaps kernel, no second pass over the gap table);
//   lists    : one thread per item walks the obstacle list of the centre's cell; phase A and phase B share it;
//   certify_b: phase B's states against the clearance field only, uncertified spheres become items again.
// If the item table is full the lane walks its own lists on the spot (correctness never depends on capacity).
constexpr uint32_t CM_NO_ITEM = 0xFFFFFFFFu; // a reserved slot that was never filled (see cm_push_items)
struct CmItems {
  uint32_t *ref;   // phase A: edge * 8 + slot; phase B: index of the owed state
  float4 *centre;  // x, y, z, sphere slot (as int bits)
  uint32_t *count;
  uint32_t cap;
};

// appends the uncertified spheres of this lane's state; returns false if they did not fit
__device__ __forceinline__ bool cm_push_items(const CmItems &it, uint32_t ref, uint32_t unc, const float *centres) {
  __syncwarp(); // also orders the callers' earlier stores to `valid` before any later one of another lane
  const int lane = threadIdx.x & 31;
  const uint32_t mine = (uint32_t)__popc(unc);
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o)
      incl += v;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0)
    return true;
  uint32_t base = 0;
  if (lane == 31)
    base = atomicAdd(it.count, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (base + total > it.cap) {
    // the counter keeps the overshoot and the lists kernel clamps it to the capacity: the slots this push had
    // reserved below the capacity stay without an item, so they are marked (a stale item of an earlier call
    // would otherwise be walked against this call's edges)
    for (uint32_t o = base + (uint32_t)lane; o < it.cap; o += 32)
      it.ref[o] = CM_NO_ITEM;
    return false;
  }
  uint32_t o = base + incl - mine;
  while (unc) {
    const int j = __ffs(unc) - 1;
    unc &= unc - 1;
    it.ref[o] = ref;
    it.centre[o] = make_float4(centres[3 * j], centres[3 * j + 1], centres[3 * j + 2], __int_as_float(j));
    ++o;
  }
  return true;
}

__global__ void __launch_bounds__(CM_THREADS, CM_MINB)
cm_sample2_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ from, const double *__restrict__ to,
                  uint32_t n, double resolution, int mode, int check_first, volatile uint8_t *valid,
                  uint32_t *__restrict__ owed, uint2 *__restrict__ gaps, CmItems items) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const int dof = w.dof;
  const int lane = threadIdx.x & 31;
  const uint32_t groups_total = gridDim.x * (CM_THREADS / 8);
  for (uint32_t e0 = blockIdx.x * (CM_THREADS / 8); e0 < n; e0 += groups_total) {
    const uint32_t e = e0 + (threadIdx.x >> 3);
    const uint32_t slot = threadIdx.x & 7;
    const bool in = e < n;
    uint32_t reach = 0, unc = 0, np = 0, stride = 1;
    float centres[3 * WORLD_MAX_S];
    double q[WORLD_MAX_DOF];
    bool hit = false;
    if (in) {
      double a[WORLD_MAX_DOF], b[WORLD_MAX_DOF];
#pragma unroll
      for (int i = 0; i < WORLD_MAX_DOF; ++i) {
        a[i] = i < dof ? from[(size_t)e * dof + i] : 0.0;
        b[i] = i < dof ? to[(size_t)e * dof + i] : 0.0;
      }
      const EdgePlan pl = plan_edge(a, b, resolution, mode, check_first);
      np = pl.np;
      stride = pl.stride;
      const uint32_t p = slot * pl.stride;
      if (p < pl.np) {
        edge_state(pl, a, b, mode, p, q);
        float travel = 0.f;
#pragma unroll
        for (int i = 0; i < WORLD_MAX_DOF; ++i)
          travel += (float)fabs(b[i] - a[i]) * (i < dof ? w.rho[i] : 0.f);
        travel *= (float)(mode == PRGPU_MODE_OMPL_DISCRETE ? 1.0 / (double)pl.nd : pl.step) * (1.0f + 1e-5f);
        float clear = 0.f;
        unc = grid_certify(w, *tabs.arm, q, &clear, centres, &hit);
        if (unc == 0 && clear > 0.f && !hit) {
          const float r = travel > 0.f ? floorf(clear / travel * (1.0f - 1e-5f)) : (float)pl.np;
          reach = r >= (float)pl.np ? pl.np : (uint32_t)r;
        }
      }
    }
    // a sample that certainly collides (two-sided field) finishes its edge: none of the edge's 8 lanes has
    // anything left to look up
    const bool edge_hit = (__ballot_sync(0xffffffffu, hit) & (0xFFu << (lane & ~7))) != 0;
    if (edge_hit)
      unc = 0;
    if (in && slot == 0)
      valid[e] = edge_hit ? 0 : 1;
    // the uncertified spheres become items (whole warp takes part in the prefix sum)
    if (!cm_push_items(items, e * 8u + slot, unc, centres) && unc != 0) {
      if (!grid_lists_kept(w, tabs, q, unc, centres))
        valid[e] = 0;
    }
    // what the samples leave untested: lane `slot` owns the states strictly between its sample and the next
    // one (or the end); the reaches of the two samples strike states off both ends of that gap
    const int gbase = lane & ~7;
    const uint32_t reach_next = __shfl_sync(0xffffffffu, reach, gbase + ((slot + 1) & 7));
    const uint32_t reach_first = __shfl_sync(0xffffffffu, reach, gbase);
    uint32_t cnt = 0, lo = 0;
    const uint32_t p = slot * stride;
    if (in && p < np && stride > 1) {
      const uint32_t pe = min(p + stride, np); // next sample or one past the last state
      const bool last = pe == np;               // no sample at pe
      uint32_t left = reach, right;
      if (mode == PRGPU_MODE_OMPL_DISCRETE) {
        if (slot == 0)
          left = 0; // p = 0 is the END state `to`; the states after it in p order start at `from`
        right = last ? reach_first : reach_next; // the trailing states run up to `to` (sample 0)
      } else {
        right = last ? 0u : reach_next;
      }
      const uint64_t l = (uint64_t)p + 1 + left, h = (uint64_t)pe - 1;
      if (h >= l && h - l + 1 > (uint64_t)right) {
        lo = (uint32_t)l;
        cnt = (uint32_t)(h - l + 1 - right);
      }
    }
    if (in)
      gaps[(size_t)e * 8 + slot] = make_uint2(lo, cnt);
    uint32_t tot = cnt;
    tot += __shfl_xor_sync(0xffffffffu, tot, 1);
    tot += __shfl_xor_sync(0xffffffffu, tot, 2);
    tot += __shfl_xor_sync(0xffffffffu, tot, 4);
    if (in && slot == 0)
      owed[e] = tot; // edges the lists kernel finds invalid are masked out by the scan's input iterator
  }
}

// one thread per (state, link sphere) item.  PHASE 0: ref = edge * 8 + sample slot; PHASE 1: ref = owed-state index.
template <int PHASE>
__global__ void __launch_bounds__(CM_THREADS)
cm_lists_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ from, const double *__restrict__ to,
                double resolution, int mode, int check_first, volatile uint8_t *valid, CmItems items,
                const uint32_t *__restrict__ item_edge, const uint32_t *__restrict__ item_p) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const int dof = w.dof;
  const uint32_t total = min(*items.count, items.cap);
  for (uint32_t g = blockIdx.x * CM_THREADS + threadIdx.x; g < total; g += gridDim.x * CM_THREADS) {
    const uint32_t ref = items.ref[g];
    if (ref == CM_NO_ITEM)
      continue;
    const uint32_t e = PHASE == 0 ? ref >> 3 : item_edge[ref];
    if (valid[e] == 0)
      continue;
    const float4 c = items.centre[g];
    double q[WORLD_MAX_DOF];
    auto state = [&]() -> const double * { // thin-band pairs only: the fp64 verdict needs the state itself
      double a[WORLD_MAX_DOF], b[WORLD_MAX_DOF];
#pragma unroll
      for (int i = 0; i < WORLD_MAX_DOF; ++i) {
        a[i] = i < dof ? from[(size_t)e * dof + i] : 0.0;
        b[i] = i < dof ? to[(size_t)e * dof + i] : 0.0;
      }
      const EdgePlan pl = plan_edge(a, b, resolution, mode, check_first);
      edge_state(pl, a, b, mode, PHASE == 0 ? (ref & 7u) * pl.stride : item_p[ref], q);
      return q;
    };
    if (!grid_sphere_free(w, tabs, state, __float_as_int(c.w), c.x, c.y, c.z))
      valid[e] = 0;
  }
}

// phase B against the clearance field only
__global__ void __launch_bounds__(CM_THREADS, CM_MINB)
cm_certify_b_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ from, const double *__restrict__ to,
                    uint32_t total, double resolution, int mode, int check_first, volatile uint8_t *valid,
                    const uint32_t *__restrict__ item_edge, const uint32_t *__restrict__ item_p, CmItems items) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const int dof = w.dof;
  const uint32_t gsize = gridDim.x * CM_THREADS;
  const uint32_t rounds = (total + gsize - 1) / gsize; // whole warps stay in the loop (shuffles in cm_push_items)
  for (uint32_t r = 0; r < rounds; ++r) {
    const uint32_t g = r * gsize + blockIdx.x * CM_THREADS + threadIdx.x;
    uint32_t unc = 0, e = 0;
    float centres[3 * WORLD_MAX_S];
    double q[WORLD_MAX_DOF];
    if (g < total) {
      e = item_edge[g];
      if (valid[e] != 0) { // else: another state of this edge already failed
        double a[WORLD_MAX_DOF], b[WORLD_MAX_DOF];
#pragma unroll
        for (int i = 0; i < WORLD_MAX_DOF; ++i) {
          a[i] = i < dof ? from[(size_t)e * dof + i] : 0.0;
          b[i] = i < dof ? to[(size_t)e * dof + i] : 0.0;
        }
        const EdgePlan pl = plan_edge(a, b, resolution, mode, check_first);
        edge_state(pl, a, b, mode, item_p[g], q);
        float clear;
        bool hit;
        unc = grid_certify(w, *tabs.arm, q, &clear, centres, &hit);
        if (hit) { // certainly colliding: the edge is finished
          valid[e] = 0;
          unc = 0;
        }
      }
    }
    if (!cm_push_items(items, g, unc, centres) && unc != 0) {
      if (!grid_lists_kept(w, tabs, q, unc, centres))
        valid[e] = 0;
    }
  }
}

// scan input: states still owed by an edge that no sample has invalidated
struct CmOwedOp {
  const uint32_t *owed;
  const uint8_t *valid;
  __host__ __device__ __forceinline__ uint32_t operator()(uint32_t e) const { return valid[e] ? owed[e] : 0u; }
};
__global__ void cm_expand2_kernel(uint32_t n, const uint8_t *__restrict__ valid, const uint32_t *__restrict__ owed,
                                  const uint32_t *__restrict__ off, const uint2 *__restrict__ gaps,
                                  uint32_t *__restrict__ item_edge, uint32_t *__restrict__ item_p,
                                  uint32_t *__restrict__ total_out) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n)
    return;
  const uint32_t mine = valid[e] ? owed[e] : 0u;
  if (e == n - 1)
    *total_out = off[e] + mine;
  if (mine == 0)
    return;
  uint32_t o = off[e];
  for (int s = 0; s < 8; ++s) {
    const uint2 g = gaps[(size_t)e * 8 + s];
    for (uint32_t i = 0; i < g.y; ++i) {
      item_edge[o] = e;
      item_p[o] = g.x + i;
      ++o;
    }
  }
}

// item g of the owed states is tested state item_p[g] of edge item_edge[g]
__global__ void cm_expand_kernel(uint32_t n, const uint32_t *__restrict__ owed, const uint32_t *__restrict__ off,
                                 const uint2 *__restrict__ gaps, uint32_t *__restrict__ item_edge,
                                 uint32_t *__restrict__ item_p) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || owed[e] == 0)
    return;
  uint32_t o = off[e];
  for (int s = 0; s < 8; ++s) {
    const uint2 g = gaps[(size_t)e * 8 + s];
    for (uint32_t i = 0; i < g.y; ++i) {
      item_edge[o] = e;
      item_p[o] = g.x + i;
      ++o;
    }
  }
}

template <int SP>
__global__ void __launch_bounds__(CM_THREADS)
cm_phase_b_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ from, const double *__restrict__ to,
                  uint32_t total, double resolution, int mode, int check_first, volatile uint8_t *valid,
                  const uint32_t *__restrict__ item_edge, const uint32_t *__restrict__ item_p) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const int dof = w.dof;
  const uint32_t gsize = gridDim.x * CM_THREADS;
  for (uint32_t g = blockIdx.x * CM_THREADS + threadIdx.x; g < total; g += gsize) {
    const uint32_t e = item_edge[g];
    if (valid[e] == 0)
      continue; // another state of this edge already failed
    double a[WORLD_MAX_DOF], b[WORLD_MAX_DOF];
#pragma unroll
    for (int i = 0; i < WORLD_MAX_DOF; ++i) {
      a[i] = i < dof ? from[(size_t)e * dof + i] : 0.0;
      b[i] = i < dof ? to[(size_t)e * dof + i] : 0.0;
    }
    const EdgePlan pl = plan_edge(a, b, resolution, mode, check_first);
    const uint32_t p = item_p[g];
    double q[WORLD_MAX_DOF];
    edge_state(pl, a, b, mode, p, q);
    if (!state_valid_any<SP>(wp, &tabs, q))
      valid[e] = 0;
  }
}

template <int SP>
__global__ void __launch_bounds__(CM_THREADS)
is_valid_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ qs, uint64_t n, uint8_t *__restrict__ valid) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs, n > 1024); // (a handful of states: tables read in place)
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double q[WORLD_MAX_DOF];
#pragma unroll
    for (int d = 0; d < WORLD_MAX_DOF; ++d)
      q[d] = d < w.dof ? qs[i * w.dof + d] : 0.0;
    valid[i] = state_valid_any<SP>(wp, &tabs, q) ? 1 : 0;
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

// clearance field: one thread per cell, exact fp64 distance from the cell's box to every obstacle
__global__ void dfield_build_kernel(const double *__restrict__ os_d, int n_os, const double *__restrict__ ob_d, int n_ob,
                                    int nx, int ny, int nz, double ox, double oy, double oz, double cell, double slack,
                                    double cap, uint32_t *__restrict__ df) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t ncell = (size_t)nx * ny * nz;
  if (i >= ncell)
    return;
  const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((size_t)nx * ny));
  const double lo[3] = {ox + x * cell, oy + y * cell, oz + z * cell};
  const double hi[3] = {lo[0] + cell, lo[1] + cell, lo[2] + cell};
  // best: min over obstacles of the SMALLEST distance from a point of the cell (lower bound of any point's
  // clearance); worst: min over obstacles of the LARGEST distance from a point of the cell -- every point of the
  // cell is at most that far from the surface of that obstacle (the distance to a convex obstacle is a convex
  // function of the point per axis: its maximum over the cell's box is attained at a corner, axis by axis)
  double best = cap, worst = cap;
  for (int o = 0; o < n_os; ++o) {
    double d2 = 0.0, f2 = 0.0;
    for (int c = 0; c < 3; ++c) {
      const double v = os_d[4 * o + c];
      const double g = fmax(fmax(lo[c] - v, v - hi[c]), 0.0);
      d2 += g * g;
      const double f = fmax(fabs(lo[c] - v), fabs(hi[c] - v));
      f2 += f * f;
    }
    best = fmin(best, sqrt(d2) - os_d[4 * o + 3]);
    worst = fmin(worst, sqrt(f2) - os_d[4 * o + 3]);
  }
  for (int o = 0; o < n_ob; ++o) {
    double d2 = 0.0, f2 = 0.0;
    for (int c = 0; c < 3; ++c) {
      const double bl = ob_d[6 * o + c] - ob_d[6 * o + 3 + c], bh = ob_d[6 * o + c] + ob_d[6 * o + 3 + c];
      const double g = fmax(fmax(bl - hi[c], lo[c] - bh), 0.0);
      d2 += g * g;
      const double f = fmax(fmax(fmax(bl - lo[c], lo[c] - bh), fmax(bl - hi[c], hi[c] - bh)), 0.0);
      f2 += f * f;
    }
    best = fmin(best, sqrt(d2));
    worst = fmin(worst, sqrt(f2));
  }
  // fp16 pair, rounded outward (values here are <= cap; resolution ~1e-4 m at link-sphere radii)
  const __half lo16 = __float2half_rd(__double2float_rd(best * (1.0 - 1e-12) - slack));
  const __half hi16 = __float2half_ru(__double2float_ru(worst * (1.0 + 1e-12) + slack));
  df[i] = (uint32_t)__half_as_ushort(lo16) | ((uint32_t)__half_as_ushort(hi16) << 16);
}

size_t obstacle_smem(const DeviceWorld &w) { return obstacle_tables_bytes(w); }

// Measurement aid (prgpu_diag_fk_centre_error): the link-sphere centres of the PRODUCTION fp32 chain
// (grid_certify's FrameF32, the one every K2 / T1 kernel filters with) against the fp64 chain of the
// specification; err[i] = the largest centre distance over the spheres of state i.  DeviceWorld::delta must
// bound it (prgpu_world_create).
__global__ void __launch_bounds__(CM_THREADS)
diag_fk_centre_kernel(const DeviceWorld *__restrict__ wp, const double *__restrict__ qs, uint64_t n,
                      double *__restrict__ err) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs);
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  double q[WORLD_MAX_DOF];
#pragma unroll
  for (int k = 0; k < WORLD_MAX_DOF; ++k)
    q[k] = k < w.dof ? qs[i * w.dof + k] : 0.0;
  float cmin, c[3 * WORLD_MAX_S];
  grid_certify<true>(w, *tabs.arm, q, &cmin, c);
  double worst = 0.0;
  for (int j = 0; j < w.S; ++j) {
    double e[3];
    sphere_centre_exact(w, q, tabs.arm->sphere_id[j], e);
    const double dx = (double)c[3 * j] - e[0], dy = (double)c[3 * j + 1] - e[1], dz = (double)c[3 * j + 2] - e[2];
    worst = fmax(worst, sqrt(dx * dx + dy * dy + dz * dz));
  }
  err[i] = worst;
}

} // namespace

#define PRGPU_SP_DISPATCH(S, ...)                                                                  \
  do {                                                                                             \
    if (w.grid_n[0] > 0) { constexpr int SP = 0; __VA_ARGS__; }                                    \
    else if ((S) <= 8) { constexpr int SP = 8; __VA_ARGS__; }                                      \
    else if ((S) <= 16) { constexpr int SP = 16; __VA_ARGS__; }                                    \
    else if ((S) <= 24) { constexpr int SP = 24; __VA_ARGS__; }                                    \
    else { constexpr int SP = 32; __VA_ARGS__; }                                                   \
  } while (0)

int check_motion_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *from_dev, const double *to_dev, uint64_t n,
                        double resolution, int mode, int check_first, uint8_t *valid_dev, int sm_count,
                        cudaStream_t stream, KernelTimer *timer, CmScratch *scratch) {
  if (n == 0)
    return PRGPU_OK;
  if (scratch && n >= 4096) {
    if (!(resolution > 0.0))
      return fail(PRGPU_EINVAL, "check_motion: resolution must be > 0");
    if (mode != PRGPU_MODE_OMPL_DISCRETE && mode != PRGPU_MODE_NNF_ALPHA)
      return fail(PRGPU_EINVAL, "check_motion: unknown mode");
    const size_t smem = obstacle_smem(w);
    const uint64_t PASS = 1ull << 20; // edges per pass: keeps the owed-state prefix sums in 32 bits
    if (timer)
      timer->begin(stream);
    for (uint64_t b0 = 0; b0 < n; b0 += PASS) {
      const uint32_t m = (uint32_t)std::min<uint64_t>(PASS, n - b0);
      size_t tmp = 0;
      PRGPU_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)m, stream));
      if (scratch->cap < m || scratch->tmp_bytes < tmp) {
        if (scratch->owed)
          cudaFree(scratch->owed);
        if (scratch->off)
          cudaFree(scratch->off);
        if (scratch->tmp)
          cudaFree(scratch->tmp);
        if (scratch->gaps)
          cudaFree(scratch->gaps);
        if (scratch->queue)
          cudaFree(scratch->queue);
        if (scratch->queue_n)
          cudaFree(scratch->queue_n);
        scratch->gaps = scratch->queue = nullptr;
        scratch->queue_n = nullptr;
        scratch->owed = scratch->off = nullptr;
        scratch->tmp = nullptr;
        scratch->cap = 0;
        const size_t cap = std::max<size_t>(m, std::min<uint64_t>(n, PASS));
        PRGPU_CUDA_TRY(cudaMalloc(&scratch->owed, cap * sizeof(uint32_t)));
        PRGPU_CUDA_TRY(cudaMalloc(&scratch->off, cap * sizeof(uint32_t)));
        PRGPU_CUDA_TRY(cudaMalloc(&scratch->gaps, cap * 8 * sizeof(uint2)));
        PRGPU_CUDA_TRY(cudaMalloc(&scratch->queue, cap * 8 * sizeof(uint2)));
        PRGPU_CUDA_TRY(cudaMalloc(&scratch->queue_n, sizeof(uint32_t)));
        PRGPU_CUDA_TRY(cudaMalloc(&scratch->tmp, std::max<size_t>(tmp, 16) * 2));
        scratch->cap = cap;
        scratch->tmp_bytes = std::max<size_t>(tmp, 16) * 2;
      }
      const double *fa = from_dev + b0 * w.dof, *tb = to_dev + b0 * w.dof;
      uint8_t *vd = valid_dev + b0;
      const unsigned blocks_a = (unsigned)std::min<uint64_t>((m + CM_THREADS / 8 - 1) / (CM_THREADS / 8), (uint64_t)sm_count * 16);
      const unsigned blocks_b = (unsigned)sm_count * 8;
      if (w.grid_n[0] > 0 && w.df) {
        // ---- pair-dealt pipeline (worlds with an obstacle grid and a clearance field) ----
        const size_t item_cap = (size_t)scratch->cap * 12; // (state, sphere) items; overflow is handled in-kernel
        if (scratch->pair_cap < item_cap) {
          if (scratch->pair_ref)
            cudaFree(scratch->pair_ref);
          if (scratch->pair_centre)
            cudaFree(scratch->pair_centre);
          scratch->pair_ref = nullptr;
          scratch->pair_centre = nullptr;
          scratch->pair_cap = 0;
          PRGPU_CUDA_TRY(cudaMalloc(&scratch->pair_ref, item_cap * sizeof(uint32_t)));
          PRGPU_CUDA_TRY(cudaMalloc(&scratch->pair_centre, item_cap * sizeof(float4)));
          scratch->pair_cap = item_cap;
        }
        if (!scratch->pair_n) // [0] phase A items, [1] phase B items, [2] owed states of the pass
          PRGPU_CUDA_TRY(cudaMalloc(&scratch->pair_n, 4 * sizeof(uint32_t)));
        // (PRGPU_CM_ITEM_CAP shrinks the table for tests of the in-kernel overflow path: results must not change)
        const char *cap_env = getenv("PRGPU_CM_ITEM_CAP");
        const size_t use_cap = cap_env ? std::min<size_t>(scratch->pair_cap, (size_t)atoll(cap_env)) : scratch->pair_cap;
        CmItems ia{scratch->pair_ref, reinterpret_cast<float4 *>(scratch->pair_centre), scratch->pair_n, (uint32_t)use_cap};
        CmItems ib = ia;
        ib.count = scratch->pair_n + 1;
        PRGPU_CUDA_TRY(cudaFuncSetAttribute(cm_sample2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PRGPU_CUDA_TRY(cudaFuncSetAttribute(cm_lists_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PRGPU_CUDA_TRY(cudaFuncSetAttribute(cm_lists_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PRGPU_CUDA_TRY(cudaFuncSetAttribute(cm_certify_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PRGPU_CUDA_TRY(cudaMemsetAsync(scratch->pair_n, 0, 4 * sizeof(uint32_t), stream));
        cm_sample2_kernel<<<blocks_a, CM_THREADS, smem, stream>>>(w_dev, fa, tb, m, resolution, mode, check_first, vd,
                                                                 scratch->owed, scratch->gaps, ia); PRGPU_COUNT_LAUNCH(1);
        cm_lists_kernel<0><<<blocks_b, CM_THREADS, smem, stream>>>(w_dev, fa, tb, resolution, mode, check_first, vd, ia,
                                                                   nullptr, nullptr); PRGPU_COUNT_LAUNCH(1);
        auto owed_in = thrust::make_transform_iterator(thrust::make_counting_iterator<uint32_t>(0u), CmOwedOp{scratch->owed, vd});
        size_t tb2 = 0;
        PRGPU_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb2, owed_in, scratch->off, (int)m, stream));
        if (tb2 > scratch->tmp_bytes)
          return fail(PRGPU_ECUDA, "check_motion: scan scratch too small");
        tb2 = scratch->tmp_bytes;
        PRGPU_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scratch->tmp, tb2, owed_in, scratch->off, (int)m, stream)); PRGPU_COUNT_LAUNCH(2);
        // number of owed states (one small read-back per pass sizes the item table and the grid)
        uint32_t last_off = 0, last_owed = 0;
        uint8_t last_valid = 0;
        PRGPU_CUDA_TRY(cudaMemcpyAsync(&last_off, scratch->off + (m - 1), 4, cudaMemcpyDeviceToHost, stream));
        PRGPU_CUDA_TRY(cudaMemcpyAsync(&last_owed, scratch->owed + (m - 1), 4, cudaMemcpyDeviceToHost, stream));
        PRGPU_CUDA_TRY(cudaMemcpyAsync(&last_valid, vd + (m - 1), 1, cudaMemcpyDeviceToHost, stream));
        PRGPU_CUDA_TRY(cudaStreamSynchronize(stream));
        const uint32_t total = last_off + (last_valid ? last_owed : 0u);
        if (total > 0) {
          if (scratch->item_cap < total) {
            if (scratch->item_edge)
              cudaFree(scratch->item_edge);
            scratch->item_edge = nullptr;
            scratch->item_cap = 0;
            PRGPU_CUDA_TRY(cudaMalloc(&scratch->item_edge, (size_t)total * 5 / 4 * 2 * sizeof(uint32_t)));
            scratch->item_cap = (size_t)total * 5 / 4;
          }
          uint32_t *item_p = scratch->item_edge + scratch->item_cap;
          cm_expand2_kernel<<<(m + 255) / 256, 256, 0, stream>>>(m, vd, scratch->owed, scratch->off, scratch->gaps,
                                                                 scratch->item_edge, item_p, scratch->pair_n + 2); PRGPU_COUNT_LAUNCH(1);
          const unsigned nb = (unsigned)std::min<uint64_t>(((uint64_t)total + CM_THREADS - 1) / CM_THREADS, blocks_b);
          cm_certify_b_kernel<<<nb, CM_THREADS, smem, stream>>>(w_dev, fa, tb, total, resolution, mode, check_first, vd,
                                                                scratch->item_edge, item_p, ib); PRGPU_COUNT_LAUNCH(1);
          cm_lists_kernel<1><<<blocks_b, CM_THREADS, smem, stream>>>(w_dev, fa, tb, resolution, mode, check_first, vd, ib,
                                                                     scratch->item_edge, item_p); PRGPU_COUNT_LAUNCH(1);
        }
        PRGPU_CUDA_TRY(cudaGetLastError());
        continue;
      }
      PRGPU_SP_DISPATCH(w.S, {
        auto ka = cm_sample_kernel<SP>;
        auto kr = cm_resolve_kernel<SP>;
        auto kb = cm_phase_b_kernel<SP>;
        PRGPU_CUDA_TRY(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PRGPU_CUDA_TRY(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PRGPU_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PRGPU_CUDA_TRY(cudaMemsetAsync(scratch->queue_n, 0, sizeof(uint32_t), stream));
        ka<<<blocks_a, CM_THREADS, smem, stream>>>(w_dev, fa, tb, m, resolution, mode, check_first, vd, scratch->owed,
                                                   scratch->gaps, scratch->queue, scratch->queue_n); PRGPU_COUNT_LAUNCH(1);
        kr<<<blocks_b, CM_THREADS, smem, stream>>>(w_dev, fa, tb, resolution, mode, check_first, vd, scratch->queue,
                                                   scratch->queue_n); PRGPU_COUNT_LAUNCH(1);
        cm_gaps_kernel<<<(unsigned)(((uint64_t)m * 8 + CM_THREADS - 1) / CM_THREADS), CM_THREADS, 0, stream>>>(
            m, mode, vd, scratch->owed, scratch->gaps); PRGPU_COUNT_LAUNCH(1);
        size_t tb2 = scratch->tmp_bytes;
        PRGPU_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scratch->tmp, tb2, scratch->owed, scratch->off, (int)m, stream)); PRGPU_COUNT_LAUNCH(2);
        // number of owed states (one small read-back per pass sizes the item table and the grid)
        uint32_t last_off = 0, last_owed = 0;
        PRGPU_CUDA_TRY(cudaMemcpyAsync(&last_off, scratch->off + (m - 1), 4, cudaMemcpyDeviceToHost, stream));
        PRGPU_CUDA_TRY(cudaMemcpyAsync(&last_owed, scratch->owed + (m - 1), 4, cudaMemcpyDeviceToHost, stream));
        PRGPU_CUDA_TRY(cudaStreamSynchronize(stream));
        const uint32_t total = last_off + last_owed;
        if (total > 0) {
          if (scratch->item_cap < total) {
            if (scratch->item_edge)
              cudaFree(scratch->item_edge);
            scratch->item_edge = nullptr;
            scratch->item_cap = 0;
            PRGPU_CUDA_TRY(cudaMalloc(&scratch->item_edge, (size_t)total * 5 / 4 * 2 * sizeof(uint32_t)));
            scratch->item_cap = (size_t)total * 5 / 4;
          }
          uint32_t *item_p = scratch->item_edge + scratch->item_cap;
          cm_expand_kernel<<<(m + 255) / 256, 256, 0, stream>>>(m, scratch->owed, scratch->off, scratch->gaps,
                                                                scratch->item_edge, item_p); PRGPU_COUNT_LAUNCH(1);
          const unsigned nb = (unsigned)std::min<uint64_t>(((uint64_t)total + CM_THREADS - 1) / CM_THREADS, blocks_b);
          kb<<<nb, CM_THREADS, smem, stream>>>(w_dev, fa, tb, total, resolution, mode, check_first, vd,
                                               scratch->item_edge, item_p); PRGPU_COUNT_LAUNCH(1);
        }
      });
      PRGPU_CUDA_TRY(cudaGetLastError());
    }
    if (timer)
      timer->end(stream);
    return PRGPU_OK;
  }
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
    PRGPU_CUDA_TRY(set_smem_once(kern, smem));
    kern<<<(unsigned)blocks, CM_THREADS, smem, stream>>>(w_dev, from_dev, to_dev, n, resolution, mode,
                                                         check_first, valid_dev); PRGPU_COUNT_LAUNCH(1);
  });
  if (timer)
    timer->end(stream);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int dfield_build_launch(const double *os_d, int n_os, const double *ob_d, int n_ob, const int *n3, const double *org,
                        double cell, double slack, double cap, uint32_t *df_dev, cudaStream_t stream) {
  const size_t ncell = (size_t)n3[0] * n3[1] * n3[2];
  dfield_build_kernel<<<(unsigned)((ncell + 127) / 128), 128, 0, stream>>>(os_d, n_os, ob_d, n_ob, n3[0], n3[1], n3[2],
                                                                          org[0], org[1], org[2], cell, slack, cap,
                                                                          df_dev); PRGPU_COUNT_LAUNCH(1);
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
    PRGPU_CUDA_TRY(set_smem_once(kern, smem));
    kern<<<(unsigned)blocks, CM_THREADS, smem, stream>>>(w_dev, q_dev, n, valid_dev); PRGPU_COUNT_LAUNCH(1);
  });
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int diag_fk_centre_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *q_dev, uint64_t n, double *err_dev,
                          cudaStream_t stream) {
  if (n == 0)
    return PRGPU_OK;
  if (!w.df)
    return fail(PRGPU_EINVAL, "diag_fk_centre_error: the world has no clearance field (the production fp32 chain is not in use)");
  const size_t smem = obstacle_smem(w);
  PRGPU_CUDA_TRY(cudaFuncSetAttribute(diag_fk_centre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  diag_fk_centre_kernel<<<(unsigned)((n + CM_THREADS - 1) / CM_THREADS), CM_THREADS, smem, stream>>>(w_dev, q_dev, n, err_dev); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int fk_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *q_dev, uint64_t n, double *pose12_dev, cudaStream_t stream) {
  if (n == 0)
    return PRGPU_OK;
  fk_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(w_dev, q_dev, n, pose12_dev); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

} // namespace prgpu

// K1, sub-linear form: a balanced k-d tree over the tree's configuration buffer, built
// and searched on the device, returning EXACTLY what the brute-force kernels return -- the k nearest under
// the (sqrt-distance, insertion index) order of NearestNeighborsLinear / findKnnNodes
// (pr_ompl/src/RRTConnect.cpp:181, frechet/src/util.cpp:34-58) -- while touching ~3e3 of 1e7 nodes per query.
//
// Build (one sort per level, no recursion): the node order is a permutation `perm`; level l splits every
// segment [floor(s n / 2^l), floor((s+1) n / 2^l)) at its middle position after sorting the segment by the
// coordinate of the segment's widest cell side.  All segments of a level are sorted by ONE radix sort of
// (segment id << 32 | order-preserving bits of the fp32-rounded coordinate).  L = ceil(log2(n / 32)) levels
// leave leaves of 16..32 nodes, stored as fp64 SoA blocks of 32 slots (one slot per lane; unused slots hold
// +inf) with the nodes' insertion indices.  An internal node keeps 16 bytes: split dimension, the fp32 cut
// value, and its cell's bounds along that dimension.
//
// Search: one warp per query, depth-first, nearer child first.  The squared distance `rd` from the query to
// the current cell is maintained incrementally (entering the far child replaces the split dimension's term
// by the squared distance to the cutting plane; cells are nested boxes, so the term being replaced is known
// from the node's own bounds) and a subtree is skipped when rd exceeds the k-th best so far.  Because the
// sort key is the fp32 ROUNDING of a coordinate, the planes are widened by one fp32 ulp on either side; rd is
// compared with a 1e-12 relative slack; so pruning never drops a node the reference would rank.  A leaf is
// evaluated by the whole warp in the reference's arithmetic (sub, mul, add in dimension order, no FMA;
// sqrt only for a node that may enter the list) and the k best live one per lane, kept sorted by ballot +
// shuffle insertion under (distance, index).
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "kdtree.cuh"

namespace prgpu {
namespace {
// Leaf capacity = 32 * PPL slots (PPL = slots per lane, a template parameter of the search kernels): kd_build takes
// the PPL in {2, 3} that leaves the fewest empty slots for the node count at hand -- a balanced tree has 2^L leaves of
// n / 2^L nodes, so one fixed capacity would leave up to half of every leaf (traffic and arithmetic) empty.
constexpr int KD_PPL_MIN = 2, KD_PPL_MAX = 3;
#ifndef KD_WARPS_N
#define KD_WARPS_N 8
#endif
constexpr int KD_WARPS = KD_WARPS_N;  // queries per block

__device__ __forceinline__ uint32_t ord32(float f) { // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord32(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__host__ __device__ __forceinline__ uint64_t kd_bound(uint64_t s, uint64_t n, int level) { return (s * n) >> level; }
__device__ __forceinline__ uint32_t kd_segment(uint64_t p, uint64_t n, int level) {
  // the s with bound(s) <= p < bound(s + 1)
  return (uint32_t)((((p + 1) << level) + n - 1) / n - 1);
}

// global bounding box (fp32, rounded outward) by warp reduction + ordered atomics
__global__ void kd_extent_kernel(const double *__restrict__ pts, uint64_t stride, uint32_t n, int dim,
                                 uint32_t *__restrict__ lo_ord, uint32_t *__restrict__ hi_ord) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int d = 0; d < dim; ++d) {
    float lo = INFINITY, hi = -INFINITY;
    if (i < n) {
      const double x = pts[(size_t)d * stride + i];
      lo = __double2float_rd(x);
      hi = __double2float_ru(x);
    }
    for (int o = 16; o; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(lo_ord + d, ord32(lo));
      atomicMax(hi_ord + d, ord32(hi));
    }
  }
}
__global__ void kd_root_cell_kernel(const uint32_t *lo_ord, const uint32_t *hi_ord, int dim, float *cell) {
  const int d = threadIdx.x;
  if (d < dim) {
    cell[d] = unord32(lo_ord[d]);
    cell[8 + d] = unord32(hi_ord[d]);
  }
}
// per segment of level l: split dimension = widest side of its cell (16 floats: lo[8], hi[8])
__global__ void kd_choose_dim_kernel(const float *__restrict__ cells, uint32_t nseg, int dim, uint32_t first_node,
                                     KdNode *__restrict__ nodes) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg)
    return;
  const float *c = cells + (size_t)s * 16;
  int best = 0;
  float bw = -1.f;
  for (int d = 0; d < dim; ++d) {
    const float w = c[8 + d] - c[d];
    if (w > bw) {
      bw = w;
      best = d;
    }
  }
  KdNode nd;
  nd.dim = (uint32_t)best;
  nd.cut = nd.cut_lo = nd.cut_hi = 0.f;
  nd.pad[0] = nd.pad[1] = 0.f;
  nd.lo = c[best];
  nd.hi = c[8 + best];
  nodes[first_node + s] = nd;
}
__global__ void kd_keys_kernel(const double *__restrict__ pts, uint64_t stride, uint32_t n, int level,
                               const uint32_t *__restrict__ perm, const KdNode *__restrict__ nodes, uint32_t first_node,
                               uint64_t *__restrict__ keys) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n)
    return;
  const uint32_t s = kd_segment(p, n, level);
  const uint32_t d = nodes[first_node + s].dim;
  const float x = __double2float_rn(pts[(size_t)d * stride + perm[p]]);
  keys[p] = ((uint64_t)s << 32) | ord32(x);
}
__global__ void kd_iota_kernel(uint32_t *perm, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    perm[i] = i;
}
// after the sort of level l: cut value of every segment and the cells of its two children
__global__ void kd_cut_kernel(const uint64_t *__restrict__ keys, uint32_t n, int level, uint32_t nseg,
                              uint32_t first_node, KdNode *__restrict__ nodes, const float *__restrict__ cells,
                              float *__restrict__ child_cells) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg)
    return;
  const uint64_t mid = kd_bound(2 * (uint64_t)s + 1, n, level + 1); // first position of the right half
  const float cut = unord32((uint32_t)(keys[mid] & 0xFFFFFFFFull));
  KdNode nd = nodes[first_node + s];
  nd.cut = cut;
  nd.cut_lo = nextafterf(cut, -INFINITY);
  nd.cut_hi = nextafterf(cut, INFINITY);
  nodes[first_node + s] = nd;
  const float *c = cells + (size_t)s * 16;
  float *l = child_cells + (size_t)(2 * s) * 16, *r = l + 16;
  for (int i = 0; i < 16; ++i) {
    l[i] = c[i];
    r[i] = c[i];
  }
  // the sort key is the fp32 rounding of the coordinate: left nodes lie below the next fp32 above the cut,
  // right nodes above the next fp32 below it
  l[8 + nd.dim] = fminf(c[8 + nd.dim], nd.cut_hi);
  r[nd.dim] = fmaxf(c[nd.dim], nd.cut_lo);
}
__global__ void kd_leaves_kernel(const double *__restrict__ pts, uint64_t stride, uint32_t n, int dim, int levels, int ppl,
                                 const uint32_t *__restrict__ perm, double *__restrict__ leaf_pts,
                                 float *__restrict__ leaf_f32, uint32_t *__restrict__ leaf_idx) {
  const uint32_t leaf = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (leaf >= (1u << levels))
    return;
  const int KD_LEAF = 32 * ppl;
  const uint64_t b = kd_bound(leaf, n, levels), e = kd_bound((uint64_t)leaf + 1, n, levels);
  for (int j = 0; j < ppl; ++j) {
    const uint32_t slot = j * 32 + lane;
    const bool have = b + slot < e;
    const uint32_t src = have ? perm[b + slot] : 0u;
    leaf_idx[(size_t)leaf * KD_LEAF + slot] = have ? src : PRGPU_INVALID_INDEX;
    for (int d = 0; d < dim; ++d) {
      const double x = have ? pts[(size_t)d * stride + src] : (double)INFINITY;
      leaf_pts[((size_t)leaf * dim + d) * KD_LEAF + slot] = x;
      leaf_f32[((size_t)leaf * dim + d) * KD_LEAF + slot] = __double2float_rn(x); // the filter's copy (see kd_T32)
    }
  }
}

// ---- search ----------------------------------------------------------------------------------------------
// A treelet = a node and its descendants down to depth 3 (15 nodes), loaded by lanes 0..14 in ONE round
// trip: lane t = 2^v - 1 + o holds the o-th node of depth v below the treelet's root.  Right after the load
// every lane evaluates ITS node against the query (which side is nearer; by how much entering the far side
// raises the squared cell distance), so a step of the descent is two shuffles, one add and one compare.
struct KdTreelet {
  double delta; // pd^2 - off^2: what entering the far child adds to the squared cell distance
  bool left;    // the query lies on the left of the cut
};
__device__ __forceinline__ KdTreelet kd_load_treelet(const KdNode *__restrict__ nodes, uint32_t root,
                                                     uint32_t first_leaf, int lane, const double *__restrict__ xq) {
  KdTreelet t;
  t.delta = 0.0;
  t.left = true;
  if (lane < 15) {
    const int v = 31 - __clz(lane + 1);                  // depth below the root: 0..3
    const uint32_t o = (uint32_t)(lane + 1) - (1u << v); // position inside that depth
    const uint64_t idx = (((uint64_t)root + 1) << v) - 1 + o;
    if (idx < first_leaf) {
      const KdNode nd = nodes[idx];
      const double x = xq[nd.dim];
      t.left = x < (double)nd.cut;
      // distance to the far half-space (planes widened by one fp32 ulp, see kd_cut_kernel)
      const double pp = (double)(t.left ? nd.cut_lo : nd.cut_hi) - x;
      double pd = t.left ? pp : -pp;
      pd = pd > 0.0 ? pd : 0.0;
      // term of this dimension already inside the cell distance: distance to the cell's own bounds
      const double o1 = (double)nd.lo - x, o2 = x - (double)nd.hi;
      double off = o1 > o2 ? o1 : o2;
      off = off > 0.0 ? off : 0.0;
      t.delta = pd * pd - off * off;
    }
  }
  return t;
}

#ifndef KD_MINB
#define KD_MINB 4
#endif
template <int DIM, int KD_PPL_>
__global__ void __launch_bounds__(KD_WARPS * 32, KD_MINB)
kd_search_kernel(const KdNode *__restrict__ nodes, const double *__restrict__ leaf_pts,
                 const uint32_t *__restrict__ leaf_idx, const float *__restrict__ root_cell, int levels,
                 uint32_t index_base,
                 const double *__restrict__ queries, uint32_t nq, int k, uint32_t *__restrict__ idx_out,
                 double *__restrict__ dist_out, unsigned long long *__restrict__ visited) {
  constexpr int KD_LEAF = 32 * KD_PPL_;
  __shared__ double sq[KD_WARPS][8];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * KD_WARPS + w;
  if (q >= nq)
    return;
  double qx[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d)
    qx[d] = queries[(size_t)q * DIM + d];
  if (lane < DIM)
    sq[w][lane] = queries[(size_t)q * DIM + lane];
  __syncwarp();
  const double *xq = sq[w];
  // the k best so far, sorted ascending by (distance, index): lane j holds the j-th
  double best_d = INFINITY;
  uint32_t best_i = PRGPU_INVALID_INDEX;
  double r2hi = INFINITY; // largest squared sum that can still enter the list
  const uint32_t first_leaf = (1u << levels) - 1u;
  // squared distance from the query to the root cell (0 inside the bounding box of the nodes)
  double rd = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const double o = fmax(0.0, fmax((double)root_cell[d] - qx[d], qx[d] - (double)root_cell[8 + d]));
    rd += o * o;
  }
  // the stack lives in registers: lane i holds entry i (depth <= 27 < 32)
  uint32_t stk_node = 0;
  double stk_rd = 0.0;
  int sp = 0;
  uint32_t node = 0;
  KdTreelet tl = kd_load_treelet(nodes, node, first_leaf, lane, xq);
  int tv = 0;        // depth of `node` inside the held treelet
  uint32_t to = 0;   // its position inside that depth
  unsigned long long nvis = 0;
  bool alive = true;
  while (alive) {
    // ---- descend from `node` to a leaf, nearer child first, far children onto the stack ----
    while (node < first_leaf) {
      if (tv > 3) { // below the held treelet: fetch the next one (one round trip per three levels)
        tl = kd_load_treelet(nodes, node, first_leaf, lane, xq);
        tv = 0;
        to = 0;
      }
      const int src = (1 << tv) - 1 + (int)to;
      const bool left = __shfl_sync(0xFFFFFFFFu, (int)tl.left, src) != 0;
      const double rd_far = rd + __shfl_sync(0xFFFFFFFFu, tl.delta, src);
      if (rd_far * (1.0 - 1e-12) <= r2hi) {
        const uint32_t far = 2 * node + (left ? 2u : 1u);
        if (lane == sp) {
          stk_node = far;
          stk_rd = rd_far;
        }
        ++sp;
#ifndef KD_NO_PREFETCH
        // the subtree will be popped later: start moving what the pop will need (a far LEAF's rows to L2;
        // an internal node's treelet rows towards L1)
        if (far >= first_leaf) {
          const char *base = reinterpret_cast<const char *>(leaf_pts + (size_t)(far - first_leaf) * DIM * KD_LEAF);
          if (lane * 128 < DIM * KD_LEAF * 8)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + lane * 128));
        } else if (lane < 4) {
          const uint64_t idx = (((uint64_t)far + 1) << lane) - 1;
          if (idx < first_leaf)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + idx));
        }
#endif
      }
      node = 2 * node + (left ? 1u : 2u);
      ++tv;
      to = 2 * to + (left ? 0u : 1u);
    }
    // ---- leaf: KD_PPL slots per lane (independent loads), reference arithmetic ----
    const uint32_t leaf = node - first_leaf;
    const double *lp = leaf_pts + (size_t)leaf * DIM * KD_LEAF + lane;
    double c0[KD_PPL_][DIM];
    uint32_t my_i[KD_PPL_];
#pragma unroll
    for (int j = 0; j < KD_PPL_; ++j) {
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        c0[j][d] = lp[d * KD_LEAF + j * 32];
      my_i[j] = leaf_idx[(size_t)leaf * KD_LEAF + j * 32 + lane];
    }
    // while those loads are in flight: the treelet of the subtree the stack says comes next (dropped if
    // this leaf shrinks the bound below that subtree's distance)
    uint32_t nx_node = 0;
    double nx_rd = 0.0;
    KdTreelet nx_tl = tl;
    if (sp > 0) {
      nx_node = __shfl_sync(0xFFFFFFFFu, stk_node, sp - 1);
      nx_rd = __shfl_sync(0xFFFFFFFFu, stk_rd, sp - 1);
      if (nx_node < first_leaf)
        nx_tl = kd_load_treelet(nodes, nx_node, first_leaf, lane, xq);
    }
    double s[KD_PPL_];
#pragma unroll
    for (int j = 0; j < KD_PPL_; ++j) {
      double a = __dsub_rn(c0[j][0], qx[0]);
      s[j] = __dmul_rn(a, a);
#pragma unroll
      for (int d = 1; d < DIM; ++d) {
        a = __dsub_rn(c0[j][d], qx[d]);
        s[j] = __dadd_rn(s[j], __dmul_rn(a, a));
      }
    }
    nvis += 1;
#pragma unroll
    for (int j = 0; j < KD_PPL_; ++j) {
      unsigned cand = __ballot_sync(0xFFFFFFFFu, my_i[j] != PRGPU_INVALID_INDEX && s[j] <= r2hi);
      while (cand) {
        const int c = __ffs(cand) - 1;
        cand &= cand - 1;
        const double cs = __shfl_sync(0xFFFFFFFFu, s[j], c);
        const uint32_t ci = __shfl_sync(0xFFFFFFFFu, my_i[j], c);
        if (cs > r2hi)
          continue; // the bound shrank since the ballot
        const double cd = __dsqrt_rn(cs);
        // position = number of list entries that precede (cd, ci)
        const bool before = lane < k && (best_d < cd || (best_d == cd && best_i < ci));
        const int pos = __popc(__ballot_sync(0xFFFFFFFFu, before));
        if (pos >= k)
          continue;
        const double up_d = __shfl_up_sync(0xFFFFFFFFu, best_d, 1);
        const uint32_t up_i = __shfl_up_sync(0xFFFFFFFFu, best_i, 1);
        if (lane == pos) {
          best_d = cd;
          best_i = ci;
        } else if (lane > pos && lane < k) {
          best_d = up_d;
          best_i = up_i;
        }
        const double r = __shfl_sync(0xFFFFFFFFu, best_d, k - 1);
        if (r < INFINITY)
          r2hi = __dmul_ru(r, r) * (1.0 + 0x1p-50);
      }
    }
    // ---- next subtree: pop until one is still within the bound ----
    alive = false;
    bool first_pop = true;
    while (sp > 0) {
      --sp;
      const uint32_t pn = first_pop ? nx_node : __shfl_sync(0xFFFFFFFFu, stk_node, sp);
      const double prd = first_pop ? nx_rd : __shfl_sync(0xFFFFFFFFu, stk_rd, sp);
      if (prd * (1.0 - 1e-12) <= r2hi) {
        node = pn;
        rd = prd;
        if (pn < first_leaf)
          tl = first_pop ? nx_tl : kd_load_treelet(nodes, pn, first_leaf, lane, xq);
        tv = 0;
        to = 0;
        alive = true;
        break;
      }
      first_pop = false;
    }
  }
  if (lane < k) {
    idx_out[(size_t)q * k + lane] = best_i == PRGPU_INVALID_INDEX ? best_i : best_i + index_base;
    if (dist_out)
      dist_out[(size_t)q * k + lane] = best_d;
  }
  if (visited && lane == 0)
    atomicAdd(visited, nvis);
}

// ---- team search: the same traversal, with the warps of a block helping one another -----------------------
// A query is a dependent chain (descend -> leaf -> pop), ~2 us per leaf for a lone warp, so the slowest query of
// a batch bounds the kernel and a small batch cannot fill the machine.  Here a block of KD_WARPS warps owns QPB
// queries (QPB <= KD_WARPS, chosen by the host from the batch size).  Every warp is a WORKER: it explores one
// subtree of one query with a private register stack and a private sorted list, exactly as above.  A worker that
// runs out of work raises its bit in `idle`; a busy worker that sees an idle bit hands over the BOTTOM entry of
// its stack (the shallowest pending subtree) through the idle warp's mailbox.  The workers of a query share one
// pruning bound (the smallest k-th distance any of them holds -- k real nodes, hence an upper bound of the true
// k-th distance) and merge their lists into the query's list when their subtree is done; the last one writes the
// result.  Same distances, same (distance, index) order, so the output is bit-identical to the single-warp search.
struct KdSlot {                 // one query of the block
  double qx[8];
  double list_d[32];
  uint32_t list_i[32];
  unsigned long long bound;     // bits of the squared pruning bound (non-negative doubles order like their bits)
  uint32_t lock, workers, qid;
  uint32_t bound_t;             // bits of the fp32 leaf threshold that goes with `bound` (kd_T32)
  double err;                   // E of kd_T32: what rounding the query and the nodes to fp32 can move a distance by
  float qf[8];
};

// ---- the fp32 leaf filter ----------------------------------------------------------------------------------
// A leaf is first evaluated on its fp32 copy: s32 = sum_d (x32_d - q32_d)^2 in fp32 (FMA chain).  With a_d = x_d -
// q_d the exact differences and A = |a| the exact distance: x32_d = x_d (1 + al), q32_d = q_d (1 + be), the fp32
// subtraction adds a relative 2^-24, so a32 = a + z with |z| <= 2^-24 A + (1 + 2^-24) 2^-24 (|x| + |q|) =: 2^-24 A + E,
// and the FMA chain of 7 non-negative terms is within (1 + 2^-24)^8 of |a32|^2.  Hence
//   A >= (sqrt(s32) (1 - 4.2 * 2^-24) - E) / (1 + 2^-24),
// and a node with s32 > T32 := ((r (1 + 2^-21) + E)^2 (1 + 2^-19) + 1e-37, rounded UP to fp32, has A > r (1 + 2^-22):
// its fp64 distance as the reference computes it (relative error < 2^-49) is strictly above r, the k-th best so far,
// so it cannot enter the list whatever its index.  Everything else -- including NaN / inf from coordinates beyond
// fp32, and every node while the list is not full (T32 = +inf) -- is evaluated again from the fp64 block in the
// reference's own arithmetic.  The filter only rejects; the list never sees an fp32 number.
// E takes |x| <= M, the largest norm inside the root cell; 1e-40 / 1e-37 cover fp32 underflow.
__device__ __forceinline__ float kd_T32(double r, double E) {
  double t = r * (1.0 + 0x1p-21) + E;
  t = t * t * (1.0 + 0x1p-19) + 1e-37;
  return __double2float_ru(t); // +inf beyond fp32: everything is a candidate
}
struct KdMail {
  double rd;
  uint32_t node, slot;
  volatile uint32_t ready;
  uint32_t pad;
};

template <int DIM, bool F32, int KD_PPL_>
__global__ void __launch_bounds__(KD_WARPS * 32, KD_MINB)
kd_search_team_kernel(const KdNode *__restrict__ nodes, const double *__restrict__ leaf_pts,
                      const float *__restrict__ leaf_f32,
                      const uint32_t *__restrict__ leaf_idx, const float *__restrict__ root_cell,
                      const uint32_t first_leaf, // 2^levels - 1: heap index of the first leaf
                      uint32_t index_base, const double *__restrict__ queries, uint32_t nq, int k, int qpb,
                      uint32_t *__restrict__ idx_out, double *__restrict__ dist_out,
                      unsigned long long *__restrict__ visited, const KdPeerOut po) {
  constexpr int KD_LEAF = 32 * KD_PPL_;
  __shared__ KdSlot slots[KD_WARPS];
  __shared__ KdMail mail[KD_WARPS];
  __shared__ uint32_t idle; // bit w: warp w waits for work
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ALL = KD_WARPS >= 32 ? 0xFFFFFFFFu : (1u << (KD_WARPS & 31)) - 1u;
  // ---- set-up: slot w <- query blockIdx.x * qpb + w ----
  if (w < qpb) {
    const uint32_t q = blockIdx.x * (uint32_t)qpb + (uint32_t)w;
    KdSlot &sl = slots[w];
    const double x = (q < nq && lane < DIM) ? queries[(size_t)q * DIM + lane] : 0.0;
    if (lane < 8) {
      sl.qx[lane] = x;
      sl.qf[lane] = __double2float_rn(x);
    }
    sl.list_d[lane] = INFINITY;
    sl.list_i[lane] = PRGPU_INVALID_INDEX;
    // |q|^2 and M^2 (largest norm inside the root cell) for E
    double q2 = x * x, m2 = 0.0;
    if (lane < DIM) {
      const double m = fmax(fabs((double)root_cell[lane]), fabs((double)root_cell[8 + lane]));
      m2 = m * m;
    }
    for (int o = 16; o; o >>= 1) {
      q2 += __shfl_xor_sync(0xFFFFFFFFu, q2, o);
      m2 += __shfl_xor_sync(0xFFFFFFFFu, m2, o);
    }
    if (lane == 0) {
      sl.bound = (unsigned long long)__double_as_longlong((double)INFINITY);
      sl.bound_t = __float_as_uint(INFINITY);
      sl.err = (__dsqrt_ru(q2) + __dsqrt_ru(m2)) * (1.0 + 0x1p-40) * (0x1p-24 * (1.0 + 0x1p-23)) + 1e-40;
      sl.lock = 0;
      sl.workers = q < nq ? 1u : 0u;
      sl.qid = q;
    }
  }
  if (lane == 0) {
    mail[w].ready = 0;
    if (w == 0)
      idle = 0;
  }
  __syncthreads();
  bool have = w < qpb && slots[w].workers != 0;
  uint32_t slot = (uint32_t)w, node = 0;
  double rd = 0.0;
  if (have) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      const double x = slots[w].qx[d];
      const double o = fmax(0.0, fmax((double)root_cell[d] - x, x - (double)root_cell[8 + d]));
      rd += o * o;
    }
  }
  unsigned long long nvis = 0;
  while (true) {
    if (!have) {
      // ---- idle: advertise, then wait for a hand-over or for everybody to be idle ----
      if (lane == 0)
        atomicOr(&idle, 1u << w);
      bool got = false, quit = false;
      while (!got && !quit) {
        if (lane == 0) {
          if (mail[w].ready) {
            got = true;
          } else if (*(volatile uint32_t *)&idle == ALL) {
            quit = true;
          } else {
            __nanosleep(200);
          }
        }
        got = __shfl_sync(0xFFFFFFFFu, (int)got, 0) != 0;
        quit = __shfl_sync(0xFFFFFFFFu, (int)quit, 0) != 0;
      }
      if (quit)
        break;
      __threadfence_block();
      slot = mail[w].slot;
      node = mail[w].node;
      rd = mail[w].rd;
      __syncwarp();
      if (lane == 0)
        mail[w].ready = 0;
      have = true;
    }
    // ---- one work unit: the subtree `node` (cell distance rd) of query slot `slot` ----
    KdSlot &sl = slots[slot];
    const double *xq = sl.qx;
    double qx[F32 ? 1 : DIM]; // (the fp32 form reads the fp64 query from shared memory, for candidates only)
    float qf[F32 ? DIM : 1];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      if (F32)
        qf[d] = sl.qf[d];
      else
        qx[d] = xq[d];
    }
    const double err = sl.err;
    double best_d = INFINITY;
    uint32_t best_i = PRGPU_INVALID_INDEX;
    double r2loc = INFINITY; // from this worker's own list
    float t32loc = INFINITY; // the fp32 leaf threshold that goes with r2loc
    uint32_t stk_node = 0;
    double stk_rd = 0.0;
    int sp = 0;
    KdTreelet tl;
    tl.delta = 0.0;
    tl.left = true;
    if (node < first_leaf)
      tl = kd_load_treelet(nodes, node, first_leaf, lane, xq);
    int tv = 0;
    uint32_t to = 0;
    bool alive = true;
    while (alive) {
      double r2hi = fmin(r2loc, __longlong_as_double((long long)*(volatile unsigned long long *)&sl.bound));
      float t32 = fminf(t32loc, __uint_as_float(*(volatile uint32_t *)&sl.bound_t));
      while (node < first_leaf) {
        if (tv > 3) {
          tl = kd_load_treelet(nodes, node, first_leaf, lane, xq);
          tv = 0;
          to = 0;
        }
        const int src = (1 << tv) - 1 + (int)to;
        const bool left = __shfl_sync(0xFFFFFFFFu, (int)tl.left, src) != 0;
        const double rd_far = rd + __shfl_sync(0xFFFFFFFFu, tl.delta, src);
        if (rd_far * (1.0 - 1e-12) <= r2hi) {
          const uint32_t far = 2 * node + (left ? 2u : 1u);
          if (lane == sp) {
            stk_node = far;
            stk_rd = rd_far;
          }
          ++sp;
          if (far >= first_leaf) { // (the block the pop will read first -- the fp32 copy in the filtered form --
                                   //  towards L2: one bulk prefetch for the whole contiguous block)
            if (lane == 0) {
              const size_t row0 = (size_t)(far - first_leaf) * DIM * KD_LEAF;
              const char *base = F32 ? reinterpret_cast<const char *>(leaf_f32 + row0)
                                     : reinterpret_cast<const char *>(leaf_pts + row0);
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base), "r"(DIM * KD_LEAF * (F32 ? 4 : 8)));
            }
          } else if (lane < 4) {
            const uint64_t idx = (((uint64_t)far + 1) << lane) - 1;
            if (idx < first_leaf)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + idx));
          }
        }
        node = 2 * node + (left ? 1u : 2u);
        ++tv;
        to = 2 * to + (left ? 0u : 1u);
      }
      const uint32_t leaf = node - first_leaf;
      const double *lp = leaf_pts + (size_t)leaf * DIM * KD_LEAF + lane;
      double c0[F32 ? 1 : KD_PPL_][F32 ? 1 : DIM];
      float f0[F32 ? KD_PPL_ : 1][F32 ? DIM : 1];
      uint32_t my_i[KD_PPL_];
      if (F32) {
        const float *fp = leaf_f32 + (size_t)leaf * DIM * KD_LEAF + lane;
#pragma unroll
        for (int j = 0; j < KD_PPL_; ++j)
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            f0[j][d] = fp[d * KD_LEAF + j * 32];
      } else {
#pragma unroll
        for (int j = 0; j < KD_PPL_; ++j) {
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            c0[j][d] = lp[d * KD_LEAF + j * 32];
          my_i[j] = leaf_idx[(size_t)leaf * KD_LEAF + j * 32 + lane];
        }
      }
      // ---- while the leaf loads are in flight: hand the bottom of the stack to an idle warp ----
      while (sp >= 2 && *(volatile uint32_t *)&idle != 0) { // (every idle warp gets one, while the stack lasts)
        int target = -1;
        if (lane == 0) {
          const uint32_t m = *(volatile uint32_t *)&idle;
          if (m) {
            const int t = __ffs(m) - 1;
            if (atomicAnd(&idle, ~(1u << t)) & (1u << t))
              target = t;
          }
        }
        target = __shfl_sync(0xFFFFFFFFu, target, 0);
        if (target >= 0) {
          const uint32_t dn = __shfl_sync(0xFFFFFFFFu, stk_node, 0);
          const double dr = __shfl_sync(0xFFFFFFFFu, stk_rd, 0);
          if (lane == 0) {
            atomicAdd(&sl.workers, 1u);
            mail[target].slot = slot;
            mail[target].node = dn;
            mail[target].rd = dr;
            __threadfence_block();
            mail[target].ready = 1;
          }
          stk_node = __shfl_down_sync(0xFFFFFFFFu, stk_node, 1);
          stk_rd = __shfl_down_sync(0xFFFFFFFFu, stk_rd, 1);
          --sp;
        } else {
          break; // another donor was faster
        }
      }
      uint32_t nx_node = 0;
      double nx_rd = 0.0;
      KdTreelet nx_tl = tl;
      if (sp > 0) {
        nx_node = __shfl_sync(0xFFFFFFFFu, stk_node, sp - 1);
        nx_rd = __shfl_sync(0xFFFFFFFFu, stk_rd, sp - 1);
        if (nx_node < first_leaf)
          nx_tl = kd_load_treelet(nodes, nx_node, first_leaf, lane, xq);
      }
      double s[KD_PPL_];
      unsigned pass[KD_PPL_]; // fp32 form: the slots the filter could not reject
      if (F32) {
#pragma unroll
        for (int j = 0; j < KD_PPL_; ++j) {
          float a = f0[j][0] - qf[0];
          float s32 = a * a;
#pragma unroll
          for (int d = 1; d < DIM; ++d) {
            a = f0[j][d] - qf[d];
            s32 = fmaf(a, a, s32);
          }
          pass[j] = __ballot_sync(0xFFFFFFFFu, !(s32 > t32)); // (NaN passes)
        }
      } else {
#pragma unroll
        for (int j = 0; j < KD_PPL_; ++j) {
          double a = __dsub_rn(c0[j][0], qx[0]);
          s[j] = __dmul_rn(a, a);
#pragma unroll
          for (int d = 1; d < DIM; ++d) {
            a = __dsub_rn(c0[j][d], qx[d]);
            s[j] = __dadd_rn(s[j], __dmul_rn(a, a));
          }
          pass[j] = 0xFFFFFFFFu;
        }
      }
      nvis += 1;
      bool improved = false;
#pragma unroll
      for (int j = 0; j < KD_PPL_; ++j) {
        if (F32) {
          if (pass[j] == 0)
            continue;
          // the candidates again, from the fp64 block, in the reference's arithmetic
          s[j] = INFINITY;
          my_i[j] = PRGPU_INVALID_INDEX;
          if ((pass[j] >> lane) & 1u) {
            my_i[j] = leaf_idx[(size_t)leaf * KD_LEAF + j * 32 + lane];
            double a = __dsub_rn(lp[j * 32], xq[0]);
            double acc = __dmul_rn(a, a);
#pragma unroll
            for (int d = 1; d < DIM; ++d) {
              a = __dsub_rn(lp[d * KD_LEAF + j * 32], xq[d]);
              acc = __dadd_rn(acc, __dmul_rn(a, a));
            }
            s[j] = acc;
          }
        }
        unsigned cand = __ballot_sync(0xFFFFFFFFu, my_i[j] != PRGPU_INVALID_INDEX && s[j] <= r2hi);
        while (cand) {
          const int c = __ffs(cand) - 1;
          cand &= cand - 1;
          const double cs = __shfl_sync(0xFFFFFFFFu, s[j], c);
          const uint32_t ci = __shfl_sync(0xFFFFFFFFu, my_i[j], c);
          if (cs > r2hi)
            continue;
          const double cd = __dsqrt_rn(cs);
          const bool before = lane < k && (best_d < cd || (best_d == cd && best_i < ci));
          const int pos = __popc(__ballot_sync(0xFFFFFFFFu, before));
          if (pos >= k)
            continue;
          const double up_d = __shfl_up_sync(0xFFFFFFFFu, best_d, 1);
          const uint32_t up_i = __shfl_up_sync(0xFFFFFFFFu, best_i, 1);
          if (lane == pos) {
            best_d = cd;
            best_i = ci;
          } else if (lane > pos && lane < k) {
            best_d = up_d;
            best_i = up_i;
          }
          const double r = __shfl_sync(0xFFFFFFFFu, best_d, k - 1);
          if (r < INFINITY) {
            r2loc = __dmul_ru(r, r) * (1.0 + 0x1p-50);
            r2hi = fmin(r2hi, r2loc);
            improved = true;
          }
        }
      }
      if (improved) {
        if (F32) {
          t32loc = kd_T32(__shfl_sync(0xFFFFFFFFu, best_d, k - 1), err);
          t32 = fminf(t32, t32loc);
        }
        if (lane == 0) {
          atomicMin(&sl.bound, (unsigned long long)__double_as_longlong(r2loc));
          if (F32)
            atomicMin(&sl.bound_t, __float_as_uint(t32loc));
        }
      }
      r2hi = fmin(r2hi, __longlong_as_double((long long)*(volatile unsigned long long *)&sl.bound));
      alive = false;
      bool first_pop = true;
      while (sp > 0) {
        --sp;
        const uint32_t pn = first_pop ? nx_node : __shfl_sync(0xFFFFFFFFu, stk_node, sp);
        const double prd = first_pop ? nx_rd : __shfl_sync(0xFFFFFFFFu, stk_rd, sp);
        if (prd * (1.0 - 1e-12) <= r2hi) {
          node = pn;
          rd = prd;
          if (pn < first_leaf)
            tl = first_pop ? nx_tl : kd_load_treelet(nodes, pn, first_leaf, lane, xq);
          tv = 0;
          to = 0;
          alive = true;
          break;
        }
        first_pop = false;
      }
    }
    // ---- the unit is done: merge this worker's list into the query's (under the slot lock) ----
    if (lane == 0) {
      while (atomicCAS(&sl.lock, 0u, 1u) != 0u)
        __nanosleep(50);
    }
    __syncwarp();
    __threadfence_block();
    {
      double md = sl.list_d[lane];
      uint32_t mi = sl.list_i[lane];
      const int mine = __popc(__ballot_sync(0xFFFFFFFFu, lane < k && best_i != PRGPU_INVALID_INDEX));
      for (int c = 0; c < mine; ++c) { // ascending: once one entry does not enter, none of the rest does
        const double cd = __shfl_sync(0xFFFFFFFFu, best_d, c);
        const uint32_t ci = __shfl_sync(0xFFFFFFFFu, best_i, c);
        const bool before = lane < k && (md < cd || (md == cd && mi < ci));
        const int pos = __popc(__ballot_sync(0xFFFFFFFFu, before));
        if (pos >= k)
          break;
        const double up_d = __shfl_up_sync(0xFFFFFFFFu, md, 1);
        const uint32_t up_i = __shfl_up_sync(0xFFFFFFFFu, mi, 1);
        if (lane == pos) {
          md = cd;
          mi = ci;
        } else if (lane > pos && lane < k) {
          md = up_d;
          mi = up_i;
        }
      }
      sl.list_d[lane] = md;
      sl.list_i[lane] = mi;
      const double r = __shfl_sync(0xFFFFFFFFu, md, k - 1);
      uint32_t left = 0;
      if (lane == 0) {
        if (r < INFINITY) {
          atomicMin(&sl.bound, (unsigned long long)__double_as_longlong(__dmul_ru(r, r) * (1.0 + 0x1p-50)));
          if (F32)
            atomicMin(&sl.bound_t, __float_as_uint(kd_T32(r, err)));
        }
        left = atomicSub(&sl.workers, 1u) - 1u;
      }
      left = __shfl_sync(0xFFFFFFFFu, left, 0);
      if (left == 0 && lane < k) { // the query's last worker writes the result
        const uint32_t q = sl.qid;
        const uint32_t gi = mi == PRGPU_INVALID_INDEX ? mi : mi + index_base;
        if (po.nranks > 0) { // fused exchange: the row goes into every rank's window (NVLink stores)
          const size_t row = ((size_t)po.rank * nq + q) * k + lane;
          for (int r = 0; r < po.nranks; ++r) {
            reinterpret_cast<double *>(po.base[r])[row] = md;
            reinterpret_cast<uint32_t *>(po.base[r] + po.idx_off)[row] = gi;
          }
        } else {
          idx_out[(size_t)q * k + lane] = gi;
          if (dist_out)
            dist_out[(size_t)q * k + lane] = md;
        }
      }
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0)
      atomicExch(&sl.lock, 0u);
    have = false;
  }
  if (visited && lane == 0 && nvis)
    atomicAdd(visited, nvis);
}

} // namespace

void KdIndex::release() {
  for (void *p : {(void *)nodes, (void *)leaf_pts, (void *)leaf_f32, (void *)leaf_idx, (void *)visited, (void *)root_cell})
    if (p)
      cudaFree(p);
  nodes = nullptr;
  leaf_pts = nullptr;
  leaf_f32 = nullptr;
  leaf_idx = nullptr;
  visited = nullptr;
  root_cell = nullptr;
  n = 0;
  levels = 0;
}

// leaf capacity (32 * ppl slots) and tree depth for n nodes: the capacity that leaves the fewest empty slots (ties: the
// smaller leaf); PRGPU_KD_PPL forces one (tests)
void kd_leaf_capacity(uint32_t n, int *ppl_out, int *levels_out) {
  int L = 0, ppl = KD_PPL_MIN;
  double best_fill = -1.0;
  for (int c = KD_PPL_MIN; c <= KD_PPL_MAX && n >= 2u * 32u * (unsigned)c; ++c) {
    int l = 0;
    while ((((uint64_t)n + (1ull << l) - 1) >> l) > (uint64_t)(32 * c))
      ++l;
    const double fill = (double)n / ((double)(1ull << l) * 32.0 * c);
    if (fill > best_fill + 1e-9) {
      best_fill = fill;
      L = l;
      ppl = c;
    }
  }
  if (getenv("PRGPU_KD_PPL")) {
    const int c = atoi(getenv("PRGPU_KD_PPL"));
    if (c >= KD_PPL_MIN && c <= KD_PPL_MAX && n >= 2u * 32u * (unsigned)c) {
      ppl = c;
      L = 0;
      while ((((uint64_t)n + (1ull << L) - 1) >> L) > (uint64_t)(32 * c))
        ++L;
    }
  }
  *ppl_out = ppl;
  *levels_out = L;
}

int kd_build(KdIndex &kd, int dim, const double *pts_soa, uint64_t stride, uint32_t n, cudaStream_t stream) {
  kd.release();
  if (n < 2 * 32 * KD_PPL_MIN)
    return fail(PRGPU_EINVAL, "kd_build: too few nodes");
  int L = 0, ppl = KD_PPL_MIN;
  kd_leaf_capacity(n, &ppl, &L);
  const int KD_LEAF = 32 * ppl;
  const uint32_t nleaf = 1u << L;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, stream);
  uint32_t *perm_a = nullptr, *perm_b = nullptr, *ext = nullptr;
  uint64_t *keys_a = nullptr, *keys_b = nullptr;
  float *cells_a = nullptr, *cells_b = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  auto cleanup = [&]() {
    for (void *p : {(void *)perm_a, (void *)perm_b, (void *)ext, (void *)keys_a, (void *)keys_b, (void *)cells_a,
                    (void *)cells_b, tmp})
      if (p)
        cudaFree(p);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  };
#define KD_TRY(expr)                                                                               \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      cleanup();                                                                                   \
      kd.release();                                                                                \
      return fail(_e == cudaErrorMemoryAllocation ? PRGPU_ENOMEM : PRGPU_ECUDA,                    \
                  std::string("kd_build: ") + #expr + ": " + cudaGetErrorString(_e));              \
    }                                                                                              \
  } while (0)
  KD_TRY(cudaMalloc(&kd.nodes, sizeof(KdNode) * (size_t)std::max<uint32_t>(nleaf - 1, 1)));
  KD_TRY(cudaMalloc(&kd.leaf_pts, sizeof(double) * (size_t)nleaf * dim * KD_LEAF));
  KD_TRY(cudaMalloc(&kd.leaf_f32, sizeof(float) * (size_t)nleaf * dim * KD_LEAF));
  KD_TRY(cudaMalloc(&kd.leaf_idx, sizeof(uint32_t) * (size_t)nleaf * KD_LEAF));
  KD_TRY(cudaMalloc(&kd.visited, sizeof(unsigned long long)));
  KD_TRY(cudaMalloc(&kd.root_cell, sizeof(float) * 16));
  KD_TRY(cudaMemsetAsync(kd.visited, 0, sizeof(unsigned long long), stream));
  KD_TRY(cudaMalloc(&perm_a, sizeof(uint32_t) * (size_t)n));
  KD_TRY(cudaMalloc(&perm_b, sizeof(uint32_t) * (size_t)n));
  KD_TRY(cudaMalloc(&keys_a, sizeof(uint64_t) * (size_t)n));
  KD_TRY(cudaMalloc(&keys_b, sizeof(uint64_t) * (size_t)n));
  KD_TRY(cudaMalloc(&cells_a, sizeof(float) * 16 * (size_t)nleaf));
  KD_TRY(cudaMalloc(&cells_b, sizeof(float) * 16 * (size_t)nleaf));
  KD_TRY(cudaMalloc(&ext, sizeof(uint32_t) * 16));
  KD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, perm_a, perm_b, (int)n, 0, 64, stream));
  KD_TRY(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16)));
  const unsigned nb = (unsigned)((n + 255) / 256);
  KD_TRY(cudaMemsetAsync(ext, 0xFF, sizeof(uint32_t) * 8, stream));     // lo: max ordered value
  KD_TRY(cudaMemsetAsync(ext + 8, 0x00, sizeof(uint32_t) * 8, stream)); // hi: min ordered value
  kd_extent_kernel<<<nb, 256, 0, stream>>>(pts_soa, stride, n, dim, ext, ext + 8); PRGPU_COUNT_LAUNCH(1);
  KD_TRY(cudaMemsetAsync(cells_a, 0, sizeof(float) * 16, stream));
  kd_root_cell_kernel<<<1, 32, 0, stream>>>(ext, ext + 8, dim, cells_a); PRGPU_COUNT_LAUNCH(1);
  KD_TRY(cudaMemcpyAsync(kd.root_cell, cells_a, sizeof(float) * 16, cudaMemcpyDeviceToDevice, stream));
  kd_iota_kernel<<<nb, 256, 0, stream>>>(perm_a, n); PRGPU_COUNT_LAUNCH(1);
  for (int l = 0; l < L; ++l) {
    const uint32_t nseg = 1u << l, first = nseg - 1u;
    kd_choose_dim_kernel<<<(nseg + 127) / 128, 128, 0, stream>>>(cells_a, nseg, dim, first, kd.nodes); PRGPU_COUNT_LAUNCH(1);
    kd_keys_kernel<<<nb, 256, 0, stream>>>(pts_soa, stride, n, l, perm_a, kd.nodes, first, keys_a); PRGPU_COUNT_LAUNCH(1);
    KD_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_b, perm_a, perm_b, (int)n, 0, 32 + l, stream));
    PRGPU_COUNT_LAUNCH(5);
    kd_cut_kernel<<<(nseg + 127) / 128, 128, 0, stream>>>(keys_b, n, l, nseg, first, kd.nodes, cells_a, cells_b); PRGPU_COUNT_LAUNCH(1);
    std::swap(perm_a, perm_b);
    std::swap(cells_a, cells_b);
  }
  kd_leaves_kernel<<<(nleaf + 7) / 8, 256, 0, stream>>>(pts_soa, stride, n, dim, L, ppl, perm_a, kd.leaf_pts, kd.leaf_f32, kd.leaf_idx); PRGPU_COUNT_LAUNCH(1);
  KD_TRY(cudaGetLastError());
  cudaEventRecord(e1, stream);
  KD_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  kd.build_ms = ms;
  kd.n = n;
  kd.levels = L;
  kd.ppl = ppl;
  kd.dim = dim;
  ++kd.builds;
  cleanup();
  return PRGPU_OK;
#undef KD_TRY
}

namespace {
template <int D, bool F32, int PPL>
void kd_launch_team_one(const KdIndex &kd, unsigned tblocks, uint32_t index_base, const double *queries_dev, uint32_t nq,
                        int k, int qpb, uint32_t *idx_dev, double *dist_dev, const KdPeerOut &po, cudaStream_t stream) {
  kd_search_team_kernel<D, F32, PPL><<<tblocks, KD_WARPS * 32, 0, stream>>>(
      kd.nodes, kd.leaf_pts, kd.leaf_f32, kd.leaf_idx, kd.root_cell, (1u << kd.levels) - 1u, index_base, queries_dev, nq, k,
      qpb, idx_dev, dist_dev, kd.count_visits ? kd.visited : nullptr, po);
}
template <int D>
void kd_launch_team_dim(const KdIndex &kd, bool f32, unsigned tblocks, uint32_t index_base, const double *queries_dev,
                        uint32_t nq, int k, int qpb, uint32_t *idx_dev, double *dist_dev, const KdPeerOut &po,
                        cudaStream_t stream) {
#define KD_GO(F, P) kd_launch_team_one<D, F, P>(kd, tblocks, index_base, queries_dev, nq, k, qpb, idx_dev, dist_dev, po, stream)
#define KD_GO2(P)                                                                                  \
  case P:                                                                                          \
    if (f32)                                                                                       \
      KD_GO(true, P);                                                                              \
    else                                                                                           \
      KD_GO(false, P);                                                                             \
    break;
  switch (kd.ppl) {
    KD_GO2(2)
    KD_GO2(3)
  }
#undef KD_GO2
#undef KD_GO
}
bool kd_launch_team(const KdIndex &kd, bool f32, unsigned tblocks, uint32_t index_base, const double *queries_dev,
                    uint32_t nq, int k, int qpb, uint32_t *idx_dev, double *dist_dev, const KdPeerOut &po,
                    cudaStream_t stream) {
  switch (kd.dim) {
#define KD_CASE(D)                                                                                                  \
  case D:                                                                                                           \
    kd_launch_team_dim<D>(kd, f32, tblocks, index_base, queries_dev, nq, k, qpb, idx_dev, dist_dev, po, stream);    \
    return true;
    KD_CASE(1)
    KD_CASE(2)
    KD_CASE(3)
    KD_CASE(4)
    KD_CASE(5)
    KD_CASE(6)
    KD_CASE(7)
    KD_CASE(8)
#undef KD_CASE
  }
  return false;
}
bool kd_launch_single(const KdIndex &kd, unsigned blocks, uint32_t index_base, const double *queries_dev, uint32_t nq,
                      int k, uint32_t *idx_dev, double *dist_dev, cudaStream_t stream) {
#define KD_S(D, P)                                                                                                  \
  case P:                                                                                                           \
    kd_search_kernel<D, P><<<blocks, KD_WARPS * 32, 0, stream>>>(kd.nodes, kd.leaf_pts, kd.leaf_idx, kd.root_cell,   \
                                                                 kd.levels, index_base, queries_dev, nq, k, idx_dev, \
                                                                 dist_dev, kd.count_visits ? kd.visited : nullptr); \
    break;
  switch (kd.dim) {
#define KD_CASE(D)                                                                                                  \
  case D:                                                                                                           \
    switch (kd.ppl) {                                                                                               \
      KD_S(D, 2) KD_S(D, 3)                                                                                         \
    }                                                                                                               \
    return true;
    KD_CASE(1)
    KD_CASE(2)
    KD_CASE(3)
    KD_CASE(4)
    KD_CASE(5)
    KD_CASE(6)
    KD_CASE(7)
    KD_CASE(8)
#undef KD_CASE
  }
  return false;
}
} // namespace

int kd_search(const KdIndex &kd, uint32_t index_base, const double *queries_dev, uint32_t nq, int k, uint32_t *idx_dev,
              double *dist_dev, cudaStream_t stream, KernelTimer *timer, const KdPeerOut *peers) {
  KdPeerOut po;
  std::memset(&po, 0, sizeof(po));
  if (peers)
    po = *peers;
  if (k < 1 || k > 32)
    return fail(PRGPU_ELIMIT, "kd_search: k must be in [1,32]");
  if (nq == 0)
    return PRGPU_OK;
  static const int team_env = getenv("PRGPU_KD_TEAM") ? atoi(getenv("PRGPU_KD_TEAM")) : -1;
  if (peers && (team_env == 0 || KD_WARPS < 2))
    return fail(PRGPU_EINVAL, "kd_search: the fused exchange needs the team kernel");
  if (team_env != 0 && KD_WARPS >= 2) {
    // queries per block: all KD_WARPS warps own a query when the batch alone fills the machine; smaller batches
    // get several warps per query from the start (PRGPU_KD_TEAM = queries per block, 0 = the single-warp kernel)
    // ... so that the blocks of the batch fill every block slot of the machine once: ceil(nq / slots)
    static int slots = 0;
    if (!slots) {
      int dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      slots = std::max(1, sms) * KD_MINB;
    }
    int qpb = (int)std::min<uint64_t>(KD_WARPS, ((uint64_t)nq + slots - 1) / slots);
    if (team_env > 0)
      qpb = std::min(team_env, (int)KD_WARPS);
    const unsigned tblocks = (nq + qpb - 1) / qpb;
    // leaves filtered on their fp32 copy (default; PRGPU_KD_F32=0 evaluates every leaf from the fp64 block -- same
    // results, kept for A/B tests)
    const bool f32 = !(getenv("PRGPU_KD_F32") && atoi(getenv("PRGPU_KD_F32")) == 0);
    if (timer)
      timer->begin(stream);
    if (!kd_launch_team(kd, f32, tblocks, index_base, queries_dev, nq, k, qpb, idx_dev, dist_dev, po, stream))
      return fail(PRGPU_ELIMIT, "kd_search: dim must be in [1,8]");
    PRGPU_COUNT_LAUNCH(1);
    if (timer)
      timer->end(stream);
    PRGPU_CUDA_TRY(cudaGetLastError());
    return PRGPU_OK;
  }
  const unsigned blocks = (nq + KD_WARPS - 1) / KD_WARPS;
  if (timer)
    timer->begin(stream);
  if (!kd_launch_single(kd, blocks, index_base, queries_dev, nq, k, idx_dev, dist_dev, stream))
    return fail(PRGPU_ELIMIT, "kd_search: dim must be in [1,8]");
  PRGPU_COUNT_LAUNCH(1);
  if (timer)
    timer->end(stream);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}
} // namespace prgpu

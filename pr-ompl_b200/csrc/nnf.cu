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
#include <cub/device/device_scan.cuh>
#include "nnf.cuh"
#include "edge_plan.cuh"

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
  if (!wp) // host-callback form: the user's FK runs on the host
    return;
  const DeviceWorld &w = *wp;
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

// node weight f[r,s]: the host's table (user metric, frechet/include/NNFrechet.hpp:312) or the translation metric
__device__ __forceinline__ double node_weight(const double *__restrict__ ftab, int R, const double *__restrict__ pref,
                                              const double *__restrict__ pnn, uint32_t r, uint32_t s) {
  return ftab ? ftab[(size_t)s * R + r] : task_distance(pref + 3 * (size_t)r, pnn + 3 * (size_t)s);
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
          const double fv = node_weight(g.ftab, R, g.pref, g.pnn, r, s);
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
    const double fv = node_weight(g.ftab, g.R, g.pref, g.pnn, r, s);
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

// ------------------------------ banded column-wise dynamic program ------------------------------

// lower bound on any path's bottleneck: every path visits every row, so cost >= max_r min_n f[r,n]
__global__ void nnf_rowmin_kernel(int R, uint32_t V, const double *__restrict__ pnn, const double *__restrict__ pref,
                                  const double *__restrict__ ftab, unsigned long long *__restrict__ rowmin_bits) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t r = blockIdx.y;
  double f = NNF_INF;
  if (s < V)
    f = node_weight(ftab, R, pref, pnn, r, s);
  // non-negative doubles order like their bit patterns
  unsigned long long b = (unsigned long long)__double_as_longlong(f);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, b, off);
    b = o < b ? o : b;
  }
  if ((threadIdx.x & 31) == 0)
    atomicMin(rowmin_bits + r, b);
}

__global__ void nnf_band_kernel(int R, uint32_t V, const double *__restrict__ pnn, const double *__restrict__ pref,
                                const double *__restrict__ ftab, double tau, uint32_t *__restrict__ band_lo,
                                unsigned long long *__restrict__ band_w64) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= V)
    return;
  int first = -1, last = -1;
  for (int r = 0; r < R; ++r) {
    const double f = node_weight(ftab, R, pref, pnn, (uint32_t)r, s);
    if (f <= tau) {
      if (first < 0)
        first = r;
      last = r;
    }
  }
  band_lo[s] = first < 0 ? 0u : (uint32_t)first;
  band_w64[s] = first < 0 ? 0ull : (unsigned long long)(last - first + 1);
}

__global__ void nnf_band_pack_kernel(uint32_t V, const uint32_t *__restrict__ band_lo,
                                     const unsigned long long *__restrict__ band_w64,
                                     const unsigned long long *__restrict__ col_off,
                                     const uint32_t *__restrict__ pred_off, const uint32_t *__restrict__ pred_pos,
                                     const uint32_t *__restrict__ pred_eid, const uint8_t *__restrict__ blocked,
                                     NnfVertexMeta *__restrict__ meta, NnfEdgeRec *__restrict__ edge) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= V)
    return;
  NnfVertexMeta m;
  m.col_off = col_off[s];
  m.band_lo = band_lo[s];
  m.band_w = (uint32_t)band_w64[s];
  m.pred_begin = pred_off[s];
  m.pred_end = pred_off[s + 1];
  m.pad0 = m.pad1 = 0;
  meta[s] = m;
  for (uint32_t e = m.pred_begin; e < m.pred_end; ++e) {
    const uint32_t p = pred_pos[e];
    NnfEdgeRec rec;
    rec.col_off = col_off[p];
    rec.band_lo = band_lo[p];
    rec.band_w_blocked = (uint32_t)band_w64[p] | (blocked[pred_eid[e]] ? 0x80000000u : 0u);
    edge[e] = rec;
  }
}

// clamp function g(x) = max(a, min(x, b)), a <= b; (g2 o g1) is again a clamp
struct Clamp {
  double a, b;
};
__device__ __forceinline__ Clamp compose(const Clamp &g2, const Clamp &g1) { // g2 after g1
  Clamp c;
  c.a = fmax(g2.a, fmin(g1.a, g2.b));
  c.b = fmax(g2.a, fmin(g1.b, g2.b));
  return c;
}

constexpr int BAND_THREADS = 1024;

// One thread-block cluster sweeps the NN-graph levels; a warp owns a vertex, its lanes the rows of
// the vertex's band.  Row recurrence x_r = max(f_r, min(x_{r-1}, m_r)) is a composition of clamps,
// evaluated with a warp-shuffle scan, so a level costs one round of gathers plus one barrier.
__global__ void __launch_bounds__(BAND_THREADS) nnf_band_dp_kernel(const NnfBandDev g, uint32_t first_level) {
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = cluster.block_rank() * (BAND_THREADS / 32) + (threadIdx.x >> 5);
  const uint32_t nwarps = cluster.num_blocks() * (BAND_THREADS / 32);
  for (uint32_t level = first_level; level < g.Lv; ++level) {
    const uint32_t ls = g.lvl_start[level], le = g.lvl_start[level + 1];
    for (uint32_t s = ls + warp; s < le; s += nwarps) {
      const NnfVertexMeta m = g.meta[s];
      if (m.band_w == 0)
        continue;
      double px = 0.0, py = 0.0, pz = 0.0;
      if (!g.ftab) {
        px = g.pnn[3 * (size_t)s];
        py = g.pnn[3 * (size_t)s + 1];
        pz = g.pnn[3 * (size_t)s + 2];
      }
      double carry = NNF_INF; // x just below the band
      for (uint32_t c0 = 0; c0 < m.band_w; c0 += 32) {
        const bool active = c0 + lane < m.band_w;
        const uint32_t r = m.band_lo + c0 + lane;
        Clamp gfn;
        gfn.a = -NNF_INF; // identity for inactive lanes
        gfn.b = NNF_INF;
        if (active) {
          double mm = NNF_INF;
          for (uint32_t e = m.pred_begin; e < m.pred_end; ++e) {
            const NnfEdgeRec rec = g.edge[e];
            if (rec.band_w_blocked & 0x80000000u)
              continue;
            const uint32_t pw = rec.band_w_blocked;
            const uint32_t i0 = r - rec.band_lo; // row r of the predecessor (unsigned wrap = outside)
            if (i0 < pw)
              mm = fmin(mm, __ldcg(g.Bc + rec.col_off + i0));
            const uint32_t i1 = i0 - 1u; // row r-1
            if (r > 0 && i1 < pw)
              mm = fmin(mm, __ldcg(g.Bc + rec.col_off + i1));
          }
          double f;
          if (g.ftab) {
            f = g.ftab[(size_t)s * g.R + r];
          } else {
            const double dx = __dsub_rn(g.pref[3 * (size_t)r], px), dy = __dsub_rn(g.pref[3 * (size_t)r + 1], py),
                         dz = __dsub_rn(g.pref[3 * (size_t)r + 2], pz);
            f = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
          }
          gfn.a = f;
          gfn.b = fmax(f, mm);
          if (s == g.start_pos && r == 0)
            gfn.a = gfn.b = 0.0; // rhs(start) = 0
        }
        // inclusive scan of compositions: lane i ends with g_i o ... o g_0
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          Clamp lower;
          lower.a = __shfl_up_sync(0xffffffffu, gfn.a, off);
          lower.b = __shfl_up_sync(0xffffffffu, gfn.b, off);
          if ((int)lane >= off)
            gfn = compose(gfn, lower);
        }
        const double x = fmax(gfn.a, fmin(carry, gfn.b));
        if (active)
          __stcg(g.Bc + m.col_off + c0 + lane, x);
        const uint32_t last = min(31u, m.band_w - c0 - 1u);
        carry = __shfl_sync(0xffffffffu, x, last);
      }
    }
    cluster.sync();
  }
}

// Dataflow form of the same banded sweep: no barriers at all.  Warps of one cooperative (co-resident)
// grid take the vertices in level-sorted order; a warp waits only for the "done" flags of its own
// vertex's predecessors (which sit earlier in that order, hence belong to warps that are running),
// computes the column, publishes it and raises the vertex's flag.  The critical path is the chain of
// dependent vertices (one flag + one gather round trip per NN-graph level) instead of one grid- or
// cluster-wide barrier per level, and every SM takes part.
constexpr int FLOW_THREADS = 256;

// One vertex's column of the banded table, by one warp (lanes = rows, 32 at a time).  Row recurrence
// x_r = max(f_r, min(x_{r-1}, m_r)), m_r = min over the predecessors' cells (r, p) and (r-1, p), evaluated as a
// shuffle scan of clamp compositions.  The records of up to 32 predecessors are fetched by different lanes and
// handed round by shuffles, the cells of four predecessors are in flight at a time.  While gathering, every lane
// also notes which predecessor cell the fixed-order walk of followBackpointers' restatement would take from its
// row: the first cell with value <= f_r if there is one, else the first that attains m_r; composed with "stays on
// the vertex while x_{r-1} <= max(f_r, m_r)" this gives the hop word of every cell (nnf.cuh).
// COMPARE: returns whether any stored value changed (warp-uniform); else returns true.
template <bool COMPARE>
__device__ __forceinline__ bool band_column(const NnfBandDev &g, uint32_t s, const NnfVertexMeta &m, uint32_t lane) {
  double px = 0.0, py = 0.0, pz = 0.0;
  if (!g.ftab) {
    px = g.pnn[3 * (size_t)s];
    py = g.pnn[3 * (size_t)s + 1];
    pz = g.pnn[3 * (size_t)s + 2];
  }
  double carry = NNF_INF;
  unsigned long long carry_hop = NNF_HOP_NONE;
  bool changed = false;
  for (uint32_t c0 = 0; c0 < m.band_w; c0 += 32) {
    const bool active = c0 + lane < m.band_w;
    const uint32_t r = m.band_lo + c0 + lane;
    double f = 0.0, old = 0.0;
    if (COMPARE && active)
      old = __ldcg(g.Bc + m.col_off + c0 + lane); // in flight under the gather (the repair wave is a latency chain)
    if (active) {
      if (g.ftab) {
        f = g.ftab[(size_t)s * g.R + r];
      } else {
        const double dx = __dsub_rn(g.pref[3 * (size_t)r], px), dy = __dsub_rn(g.pref[3 * (size_t)r + 1], py),
                     dz = __dsub_rn(g.pref[3 * (size_t)r + 2], pz);
        f = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      }
    }
    double mm = NNF_INF;
    uint32_t first_pp = 0; // lane j: sorted position of predecessor j (the first 32 of the list)
    uint32_t amin = 0xFFFFFFFFu, fle = 0xFFFFFFFFu; // (predecessor index in the list << 1) | (row r-1 ? 1 : 0)
    for (uint32_t e0 = m.pred_begin; e0 < m.pred_end; e0 += 32) {
      const uint32_t ne = min(32u, m.pred_end - e0);
      NnfEdgeRec mine;
      mine.col_off = 0;
      mine.band_lo = 0;
      mine.band_w_blocked = 0x80000000u;
      if (lane < ne) {
        mine = g.edge[e0 + lane];
        if (e0 == m.pred_begin)
          first_pp = g.pred_pos[e0 + lane]; // for the hop word, fetched with the records instead of after the scan
      }
      for (uint32_t j0 = 0; j0 < ne; j0 += 4) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int src = (int)min(j0 + u, 31u);
          const uint64_t co = __shfl_sync(0xffffffffu, mine.col_off, src);
          const uint32_t bl = __shfl_sync(0xffffffffu, mine.band_lo, src);
          uint32_t bw = __shfl_sync(0xffffffffu, mine.band_w_blocked, src);
          if (j0 + u >= ne)
            bw = 0x80000000u;
          v[2 * u] = NNF_INF;
          v[2 * u + 1] = NNF_INF;
          if (active && !(bw & 0x80000000u)) {
            const uint32_t i0 = r - bl;
            if (i0 < bw)
              v[2 * u] = __ldcg(g.Bc + co + i0);
            const uint32_t i1 = i0 - 1u;
            if (r > 0 && i1 < bw)
              v[2 * u + 1] = __ldcg(g.Bc + co + i1);
          }
        }
        const uint32_t code0 = (e0 - m.pred_begin + j0) << 1;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (v[u] < mm) {
            mm = v[u];
            amin = code0 + (uint32_t)u;
          }
          if (fle == 0xFFFFFFFFu && v[u] <= f)
            fle = code0 + (uint32_t)u;
        }
      }
    }
    const bool is_start = s == g.start_pos && r == 0;
    Clamp gfn;
    gfn.a = -NNF_INF;
    gfn.b = NNF_INF;
    if (active) {
      gfn.a = f;
      gfn.b = fmax(f, mm);
      if (is_start)
        gfn.a = gfn.b = 0.0;
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      Clamp lower;
      lower.a = __shfl_up_sync(0xffffffffu, gfn.a, off);
      lower.b = __shfl_up_sync(0xffffffffu, gfn.b, off);
      if ((int)lane >= off)
        gfn = compose(gfn, lower);
    }
    const double x = fmax(gfn.a, fmin(carry, gfn.b));
    const uint32_t last = min(31u, m.band_w - c0 - 1u);
    unsigned long long h = NNF_HOP_NONE;
    if (g.hop) {
      double xprev = __shfl_up_sync(0xffffffffu, x, 1);
      if (lane == 0)
        xprev = carry;
      // the walk stays on the vertex: (r-1, s) explains the cell (inactive lanes inherit, they are never stored)
      const bool stays = !active || (!is_start && xprev != NNF_INF && xprev <= fmax(f, mm));
      unsigned long long own = NNF_HOP_NONE;
      const bool need = active && !stays && !is_start && x != NNF_INF;
      const uint32_t code = fle != 0xFFFFFFFFu ? fle : amin;
      // the chosen predecessor's position: handed round from the lane that fetched it with the records
      const uint32_t pj = need ? (code >> 1) : 0u;
      uint32_t pp = __shfl_sync(0xffffffffu, first_pp, (int)(pj & 31u));
      if (active && !stays) {
        if (is_start) {
          own = NNF_HOP_START;
        } else if (x != NNF_INF) {
          const uint32_t e = m.pred_begin + pj;
          if (pj >= 32u)
            pp = g.pred_pos[e];
          const NnfEdgeRec rec = g.edge[e];
          const unsigned long long cell = rec.col_off + (r - rec.band_lo) - (code & 1u);
          own = (cell << 28) | ((unsigned long long)pp << 1) | (code & 1u);
        }
      }
      const unsigned nonstay = __ballot_sync(0xffffffffu, !stays) & ((2u << lane) - 1u);
      const int src = nonstay ? 31 - __clz(nonstay) : 0;
      h = __shfl_sync(0xffffffffu, own, src);
      if (!nonstay)
        h = carry_hop;
      carry_hop = __shfl_sync(0xffffffffu, h, last);
    }
    if (active) {
      if (COMPARE)
        changed |= old != x;
      __stcg(g.Bc + m.col_off + c0 + lane, x);
      if (g.hop)
        __stcg(g.hop + m.col_off + c0 + lane, h);
    }
    carry = __shfl_sync(0xffffffffu, x, last);
  }
  return COMPARE ? __any_sync(0xffffffffu, changed) : true;
}

__global__ void __launch_bounds__(FLOW_THREADS)
nnf_flow_dp_kernel(const NnfBandDev g, uint32_t first_pos, uint32_t epoch, volatile uint32_t *done) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = blockIdx.x * (FLOW_THREADS / 32) + (threadIdx.x >> 5);
  const uint32_t nwarps = gridDim.x * (FLOW_THREADS / 32);
  for (uint32_t s = first_pos + warp; s < g.V; s += nwarps) {
    const NnfVertexMeta m = g.meta[s];
    if (m.band_w != 0) {
      // wait until every (non-blocked) predecessor that is being recomputed has published its column
      for (uint32_t e0 = m.pred_begin; e0 < m.pred_end; e0 += 32) {
        const uint32_t e = e0 + lane;
        if (e < m.pred_end && !(g.edge[e].band_w_blocked & 0x80000000u)) {
          const uint32_t p = g.pred_pos[e];
          if (p >= first_pos) {
            unsigned ns = 32;
            while (done[p] != epoch) {
              __nanosleep(ns);
              if (ns < 256)
                ns *= 2;
            }
          }
        }
      }
      __threadfence(); // acquire: the columns behind the flags
      __syncwarp();
      band_column<false>(g, s, m, lane);
      __threadfence(); // release: the column before the flag
    }
    __syncwarp();
    if (lane == 0)
      done[s] = epoch;
  }
}

// ---- device-resident LazySP loop (nnf.cuh: NnfLazyDev) -----------------------------------------------------
constexpr int LAZY_THREADS = 512;
constexpr uint32_t LAZY_WAVE_BUDGET = 2048; // columns one block repairs per knock-out before it hands the rest
                                            // of the wave to the whole-GPU sweep (about the cost of that sweep)

constexpr uint32_t LAZY_WL_CAP = 4096;   // dirty vertices of one wave kept in shared memory
constexpr uint32_t LAZY_HASH = 8192;     // open-addressing set of the wave's vertices (de-duplication)
constexpr uint32_t LAZY_DONE = 0xFFFFFFFFu;

template <int SP>
__global__ void __launch_bounds__(LAZY_THREADS)
nnf_lazysp_kernel(const DeviceWorld *__restrict__ wp, const NnfBandDev g, const NnfLazyDev z, uint32_t tables_bytes) {
  const DeviceWorld &w = *wp;
  extern __shared__ float4 s_obs[];
  const ObstacleTables tabs = load_tables(w, s_obs);
  // the repair wave's bookkeeping lives in shared memory: the level starts (level of a sorted position by binary
  // search), the wave's dirty vertices with their levels, a hash set against duplicates, the current level's list
  uint32_t *s_lvl_start = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(s_obs) + tables_bytes);
  uint32_t *wl_pos = s_lvl_start + ((g.Lv + 1 + 3) & ~3u);
  uint32_t *wl_level = wl_pos + LAZY_WL_CAP;
  uint32_t *hset = wl_level + LAZY_WL_CAP;
  uint32_t *cur = hset + LAZY_HASH;
  __shared__ uint32_t s_status, s_n, s_level, s_target, s_wl_n, s_overflow, s_par, s_cur_n2[2], s_next2[2];
  __shared__ unsigned long long s_err_bits;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t NW = LAZY_THREADS / 32;
  for (uint32_t i = tid; i <= g.Lv; i += LAZY_THREADS)
    s_lvl_start[i] = g.lvl_start[i];
  __syncthreads();
  auto level_of = [&](uint32_t pos) -> uint32_t { // largest L with lvl_start[L] <= pos
    uint32_t lo = 0, hi = g.Lv;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_lvl_start[mid] <= pos)
        lo = mid;
      else
        hi = mid;
    }
    return lo;
  };
  auto push = [&](uint32_t pos, uint32_t pb) { // one thread: add a dirty vertex to the wave (once)
    uint32_t hslot = (pos * 2654435761u) & (LAZY_HASH - 1u);
    while (true) {
      const uint32_t old = atomicCAS(hset + hslot, 0u, pos + 1u);
      if (old == pos + 1u)
        return; // already in the wave
      if (old == 0u)
        break;
      hslot = (hslot + 1u) & (LAZY_HASH - 1u);
    }
    const uint32_t i = atomicAdd(&s_wl_n, 1u);
    if (i >= LAZY_WL_CAP) {
      s_overflow = 1;
      return;
    }
    const uint32_t lv = level_of(pos);
    wl_pos[i] = pos;
    wl_level[i] = lv;
    atomicMin(&s_next2[s_par], lv);
    // the warp that takes it starts with these: the vertex record, its first predecessor records, its position
    asm volatile("prefetch.global.L1 [%0];" ::"l"(g.meta + pos));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(g.edge + pb));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(g.edge + pb + 4));
    if (g.pnn)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(g.pnn + 3 * (size_t)pos));
  };
  const int dof = z.dof;
  const uint32_t R = (uint32_t)g.R;
  uint32_t iters = 0, epoch = z.epoch, n_eval = 0, n_coll = 0, repaired = 0; // thread 0's copies are the truth
  double goal_cost = 0.0, final_error = -1.0; // final_error < 0: no path extracted by this launch
  uint32_t path_len = 0;
  unsigned long long cyc[5] = {0, 0, 0, 0, 0};
  long long t_mark = clock64();
  auto lap = [&](int phase) { // thread 0's phase clock
    const long long now = clock64();
    cyc[phase] += (unsigned long long)(now - t_mark);
    t_mark = now;
  };
  while (true) {
    // ---- search result: the goal's cost, then the hops of the optimum, goal -> start ----
    if (tid == 0) {
      uint32_t status = NNF_LAZY_RUNNING;
      s_err_bits = 0ull;
      if (iters >= z.max_iters) {
        status = NNF_LAZY_CONTINUE;
      } else {
        const NnfVertexMeta mg = g.meta[g.goal_pos];
        const uint32_t gi = (R - 1u) - mg.band_lo;
        const double cost = gi < mg.band_w ? __ldcg(g.Bc + mg.col_off + gi) : NNF_INF;
        if (!(cost <= z.tau_ub) && !z.band_full) {
          goal_cost = cost; // the band may not contain the optimum: the host widens it (NNF_LAZY_NEED_REBUILD)
          status = NNF_LAZY_NEED_REBUILD;
        } else {
          if (z.trace)
            z.trace[iters] = cost;
          ++iters;
          goal_cost = cost;
          if (cost == NNF_INF) {
            status = NNF_LAZY_NO_PATH; // the search found no path (NNFrechet.cpp:112-113)
          } else {
            uint32_t n = 0, pos = g.goal_pos;
            unsigned long long cell = mg.col_off + gi;
            while (true) {
              if (n >= z.cap_hops) {
                status = NNF_LAZY_ERROR;
                break;
              }
              z.path_pos[n] = pos;
              z.path_cell[n] = cell;
              ++n;
              const unsigned long long h = __ldcg(g.hop + cell);
              if (h == NNF_HOP_START)
                break;
              if (h == NNF_HOP_NONE) {
                status = NNF_LAZY_ERROR;
                break;
              }
              cell = h >> 28;
              pos = (uint32_t)(h >> 1) & (NNF_HOP_MAX_V - 1u);
            }
            s_n = n;
          }
        }
      }
      s_status = status;
      lap(0);
    }
    __syncthreads();
    if (s_status != NNF_LAZY_RUNNING)
      break;
    const uint32_t n = s_n; // hop 0 = c1 (goal), hop n-1 = c0 (start)
    // ---- per hop, a warp each: the NN edge it crossed, and max node weight over the rows it stayed (mFinalError
    //      = max over the NON-dummy path nodes, NNFrechet.cpp:629-640) ----
    for (uint32_t i = warp; i < n; i += NW) {
      const uint32_t sp = z.path_pos[i];
      const NnfVertexMeta m = g.meta[sp];
      const unsigned long long cell = z.path_cell[i];
      const uint32_t r_hi = m.band_lo + (uint32_t)(cell - m.col_off);
      const unsigned long long h = __ldcg(g.hop + cell); // (written with st.cg by the repair wave)
      uint32_t r_lo = 0;
      if (h != NNF_HOP_START) {
        const uint32_t pp = (uint32_t)(h >> 1) & (NNF_HOP_MAX_V - 1u);
        const NnfVertexMeta mp = g.meta[pp];
        r_lo = mp.band_lo + (uint32_t)((h >> 28) - mp.col_off) + (uint32_t)(h & 1ull);
        // CSR slot of the NN edge pp -> sp
        uint32_t slot = 0xFFFFFFFFu;
        for (uint32_t e = m.pred_begin + lane; e < m.pred_end; e += 32)
          if (g.pred_pos[e] == pp)
            slot = e;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
          slot = min(slot, __shfl_xor_sync(0xffffffffu, slot, off));
        if (lane == 0)
          z.path_slot[i] = slot;
      }
      if (sp != g.start_pos && sp != g.goal_pos) {
        double fm = 0.0;
        for (uint32_t r = r_lo + lane; r <= r_hi; r += 32)
          fm = fmax(fm, node_weight(g.ftab, g.R, g.pref, g.pnn, r, sp));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
          fm = fmax(fm, __shfl_xor_sync(0xffffffffu, fm, off));
        if (lane == 0)
          atomicMax(&s_err_bits, (unsigned long long)__double_as_longlong(fm)); // non-negative doubles order like their bits
      }
    }
    __syncthreads();
    if (tid == 0)
      lap(1);
    // ---- evaluateEdge for the path's NN edges whose verdict is not known yet (a warp per edge) ----
    // NN path, start -> goal: hops n-2 .. 1; the edge INTO hop i comes from hop i+1: edges i = n-3 .. 1
    for (uint32_t i = 1 + warp; i + 2 < n; i += NW) {
      const uint32_t eid = z.pred_eid[z.path_slot[i]];
      if (z.known[eid] >= 0)
        continue;
      const uint32_t vu = z.vertex_of_pos[z.path_pos[i + 1]], vv = z.vertex_of_pos[z.path_pos[i]];
      double a[WORLD_MAX_DOF], b[WORLD_MAX_DOF];
#pragma unroll
      for (int k = 0; k < WORLD_MAX_DOF; ++k) {
        a[k] = k < dof ? z.states[(size_t)vu * dof + k] : 0.0;
        b[k] = k < dof ? z.states[(size_t)vv * dof + k] : 0.0;
      }
      const EdgePlan pl = plan_edge(a, b, z.check_resolution, PRGPU_MODE_NNF_ALPHA, 0);
      bool ok = true;
      for (uint32_t p0 = 0; p0 < pl.np && ok; p0 += 32) {
        bool good = true;
        if (p0 + lane < pl.np) {
          double q[WORLD_MAX_DOF];
          edge_state(pl, a, b, PRGPU_MODE_NNF_ALPHA, p0 + lane, q);
          good = state_valid_call<SP>(wp, &tabs, q);
        }
        ok = !__any_sync(0xffffffffu, !good);
      }
      if (lane == 0)
        z.known[eid] = ok ? 1 : 0;
    }
    __syncthreads();
    // ---- the reference's bookkeeping, in path order (NNFrechet.cpp:120-151) ----
    if (tid == 0) {
      lap(2);
      final_error = __longlong_as_double((long long)s_err_bits);
      bool collision_free = true;
      for (uint32_t i = n >= 3 ? n - 3 : 0; i >= 1 && n >= 3; --i) {
        const uint32_t slot = z.path_slot[i], eid = z.pred_eid[slot];
        if (!z.evaluated[eid]) {
          z.evaluated[eid] = 1;
          ++n_eval;
          if (z.known[eid] == 0) { // markEdgeInCollision (:131-136, :685-702)
            z.blocked[eid] = 1;
            z.edge_rw[slot].band_w_blocked |= 0x80000000u;
            ++n_coll;
            collision_free = false;
            s_target = z.path_pos[i];
            break;
          }
        }
      }
      if (collision_free) {
        path_len = n >= 2 ? n - 2 : 0;
        for (uint32_t k = 0; k < path_len; ++k)
          z.out_path[k] = z.vertex_of_pos[z.path_pos[n - 2 - k]];
        s_status = NNF_LAZY_SOLVED;
      } else {
        // repair: the knocked-out edge's target is the first dirty vertex
        ++epoch;
        s_wl_n = 0;
        s_overflow = 0;
        s_par = 1; // the first push (the knocked-out edge's target) reports its level through pair 1 ...
        s_cur_n2[0] = 0;
        s_next2[0] = LAZY_DONE; // ... and the level loop starts on a fresh pair 0
        s_next2[1] = LAZY_DONE;
      }
      lap(3);
    }
    __syncthreads();
    if (s_status != NNF_LAZY_RUNNING)
      break;
    // ---- local repair, level by level: recompute the dirty vertices' columns; a changed column makes the
    //      vertex's successors dirty (they sit at later levels).  Mostly a long, thin wave (one or two vertices
    //      per level along the old optimum): a latency chain, so its bookkeeping stays in shared memory ----
    for (uint32_t i = tid; i < LAZY_HASH; i += LAZY_THREADS)
      hset[i] = 0;
    __syncthreads();
    if (tid == 0)
      push(s_target, g.meta[s_target].pred_begin);
    __syncthreads();
    auto process = [&](uint32_t sp) { // one warp: recompute a dirty vertex; a changed column dirties its successors
      const NnfVertexMeta m = g.meta[sp];
      const uint32_t so0 = z.succ_off[sp], so1 = z.succ_off[sp + 1]; // in flight while the column is computed
      const uint32_t first_succ = so0 + lane < so1 ? z.succ_pos[so0 + lane] : 0u;
      const uint32_t first_pb = so0 + lane < so1 ? z.succ_pb[so0 + lane] : 0u;
      if (m.band_w != 0 && band_column<true>(g, sp, m, lane))
        for (uint32_t e = so0 + lane; e < so1; e += 32) {
          if (e == so0 + lane)
            push(first_succ, first_pb);
          else
            push(z.succ_pos[e], z.succ_pb[e]);
        }
    };
    // block-wide, level by level; two barriers per level (the per-level counters are double-buffered: the other
    // parity's pair is reset while this level is processed).  One warp carrying a thin wave alone (no barriers at
    // all) was tried and is slower (0.85 vs 0.59 ms per iteration): a column is a ~3 us dependent chain and the
    // waves average 1.5 vertices per level, which the block overlaps.
    uint32_t L = s_next2[1], wave = 0, par = 0;
    bool bailed = false;
    while (L != LAZY_DONE) {
      // the entries of level L -> cur[]; the smallest later level -> s_next2[par] (pushes below add to it)
      const uint32_t nwl = min(s_wl_n, LAZY_WL_CAP);
      for (uint32_t i = tid; i < nwl; i += LAZY_THREADS) {
        const uint32_t lv = wl_level[i];
        if (lv == L) {
          const uint32_t j = atomicAdd(&s_cur_n2[par], 1u);
          if (j < LAZY_WAVE_BUDGET)
            cur[j] = wl_pos[i];
          wl_level[i] = LAZY_DONE;
        } else if (lv != LAZY_DONE) {
          atomicMin(&s_next2[par], lv);
        }
      }
      __syncthreads();
      const uint32_t cnt = s_cur_n2[par];
      if (tid == 0) { // the next level's pair
        s_cur_n2[par ^ 1u] = 0;
        s_next2[par ^ 1u] = LAZY_DONE;
      }
      s_par = par; // (every thread writes the same value; push() reads it)
      wave += cnt;
      if (wave > z.wave_budget || s_overflow) { // everything before level L is consistent; the host re-sweeps
        bailed = true;                             // levels >= L with the whole GPU
        break;
      }
      for (uint32_t i = warp; i < cnt; i += NW)
        process(cur[i]);
      __syncthreads();
      if (tid == 0)
        repaired += cnt;
      if (s_overflow) { // a successor did not fit: it sits at a level > L, re-sweep from L + 1
        bailed = true;
        ++L;
        break;
      }
      L = s_next2[par];
      par ^= 1u;
    }
    if (tid == 0)
      lap(4);
    if (bailed) {
      if (tid == 0) {
        s_status = NNF_LAZY_NEED_FLOW;
        s_level = L;
      }
      __syncthreads();
      break;
    }
  }
  if (tid == 0) {
    NnfLazyOut o;
    for (int i = 0; i < 5; ++i)
      o.cycles[i] = cyc[i];
    o.status = s_status;
    o.iters = iters;
    o.n_evaluated = n_eval;
    o.n_collisions = n_coll;
    o.path_len = path_len;
    o.epoch = epoch;
    o.repaired_vertices = repaired;
    o.pad = s_status == NNF_LAZY_NEED_FLOW ? s_level : 0;
    o.goal_cost = goal_cost;
    o.final_error = final_error;
    *z.out = o;
  }
}

__device__ __forceinline__ double band_get(const NnfBandDev &g, uint32_t r, uint32_t s) {
  const NnfVertexMeta m = g.meta[s];
  const uint32_t i = r - m.band_lo;
  return i < m.band_w ? g.Bc[m.col_off + i] : NNF_INF;
}

// One warp walks the path goal -> start.  A step has to find, in the fixed order of the full-table walk
// -- (r-1,s), then per predecessor p in list order (r,p), (r-1,p) -- the first cell whose value explains
// the current one.  (r-1,s) is tried alone (most steps stay on the vertex); otherwise the predecessors'
// cells are independent loads: lane 1+2j / 2+2j take the two cells of predecessor j, a ballot picks the
// lowest lane that matches (chunks of 15 predecessors keep the order), so such a step costs one dependent
// round trip instead of one per rejected candidate.
__global__ void nnf_band_backtrack_kernel(const NnfBandDev g, uint32_t *__restrict__ out_path, uint32_t cap_pairs,
                                          uint32_t *__restrict__ out_n, double *__restrict__ out_cost) {
  if (blockIdx.x != 0 || threadIdx.x >= 32)
    return;
  const int lane = threadIdx.x;
  uint32_t r = (uint32_t)g.R - 1, s = g.goal_pos, n = 0;
  double target = band_get(g, r, s);
  if (lane == 0)
    *out_cost = target;
  if (target == NNF_INF) {
    if (lane == 0)
      *out_n = 0;
    return;
  }
  while (true) {
    if (lane == 0 && n < cap_pairs) {
      out_path[2 * n] = r;
      out_path[2 * n + 1] = s;
    }
    ++n;
    if (r == 0 && s == g.start_pos)
      break;
    const double fv = node_weight(g.ftab, g.R, g.pref, g.pnn, r, s);
    const NnfVertexMeta m = g.meta[s];
    bool found = false;
    uint32_t br = 0, bs = 0;
    double bval = 0.0;
    // (r-1, s) first, on its own: most steps stay on the vertex, and its value sits next to the current one
    if (r > 0) {
      const uint32_t i = r - 1 - m.band_lo;
      const double bu = i < m.band_w ? g.Bc[m.col_off + i] : NNF_INF;
      if (bu != NNF_INF && fmax(bu, fv) == target) {
        found = true;
        br = r - 1;
        bs = s;
        bval = bu;
      }
    }
    for (uint32_t e0 = m.pred_begin; !found && e0 < m.pred_end; e0 += 15) {
      // lanes 1..30: predecessors e0 .. e0+14, two cells each
      double v = NNF_INF;
      uint32_t cr = 0, cs = 0;
      if (lane >= 1 && lane <= 30) {
        const uint32_t e = e0 + (uint32_t)(lane - 1) / 2u;
        if (e < m.pred_end) {
          const NnfEdgeRec rec = g.edge[e];
          if (!(rec.band_w_blocked & 0x80000000u)) {
            const uint32_t pw = rec.band_w_blocked;
            const uint32_t i0 = r - rec.band_lo;
            if ((lane - 1) & 1) { // (r-1, p)
              const uint32_t i1 = i0 - 1u;
              if (r > 0 && i1 < pw)
                v = g.Bc[rec.col_off + i1];
              cr = r - 1;
            } else { // (r, p)
              if (i0 < pw)
                v = g.Bc[rec.col_off + i0];
              cr = r;
            }
            cs = g.pred_pos[e];
          }
        }
      }
      const bool hit = v != NNF_INF && fmax(v, fv) == target;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        const int src = __ffs(bal) - 1;
        br = __shfl_sync(0xffffffffu, cr, src);
        bs = __shfl_sync(0xffffffffu, cs, src);
        bval = __shfl_sync(0xffffffffu, v, src);
        found = true;
      }
    }
    if (!found) { // cannot happen on a consistent table
      n = 0xFFFFFFFFu;
      break;
    }
    r = br;
    s = bs;
    target = bval;
  }
  if (lane == 0)
    *out_n = n;
}

} // namespace

int nnf_subsample_launch(const DeviceWorld *w_dev, int dof, uint32_t n_new, uint32_t first_vertex,
                         const uint32_t *src, const uint32_t *dst, const uint32_t *step, int D, double *states,
                         double *pose12, cudaStream_t stream) {
  if (n_new == 0)
    return PRGPU_OK;
  nnf_subsample_kernel<<<(n_new + 127) / 128, 128, 0, stream>>>(w_dev, dof, n_new, first_vertex, src, dst, step, D,
                                                               states, pose12); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int nnf_translations_launch(const double *pose12, const uint32_t *vertex_of_pos, uint32_t n, double *out3,
                            cudaStream_t stream) {
  if (n == 0)
    return PRGPU_OK;
  nnf_translations_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pose12, vertex_of_pos, n, out3); PRGPU_COUNT_LAUNCH(1);
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
  PRGPU_CUDA_TRY(cudaLaunchCooperativeKernel((void *)nnf_dp_kernel, grid, block, args, 0, stream)); PRGPU_COUNT_LAUNCH(1);
  if (timer)
    timer->end(stream);
  return PRGPU_OK;
}

int nnf_rowmin_launch(int R, uint32_t V, const double *pnn, const double *pref, const double *ftab,
                      unsigned long long *rowmin_bits, cudaStream_t stream) {
  PRGPU_CUDA_TRY(cudaMemsetAsync(rowmin_bits, 0xFF, (size_t)R * sizeof(unsigned long long), stream));
  nnf_rowmin_kernel<<<dim3((V + 255) / 256, R), 256, 0, stream>>>(R, V, pnn, pref, ftab, rowmin_bits); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int nnf_band_launch(int R, uint32_t V, const double *pnn, const double *pref, const double *ftab, double tau,
                    uint32_t *band_lo, unsigned long long *band_w64, cudaStream_t stream) {
  nnf_band_kernel<<<(V + 127) / 128, 128, 0, stream>>>(R, V, pnn, pref, ftab, tau, band_lo, band_w64); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int nnf_band_pack_launch(uint32_t V, uint32_t E, const uint32_t *band_lo, const unsigned long long *band_w64,
                         const unsigned long long *col_off, const uint32_t *pred_off, const uint32_t *pred_pos,
                         const uint32_t *pred_eid, const uint8_t *blocked, NnfVertexMeta *meta, NnfEdgeRec *edge,
                         cudaStream_t stream) {
  (void)E;
  nnf_band_pack_kernel<<<(V + 127) / 128, 128, 0, stream>>>(V, band_lo, band_w64, col_off, pred_off, pred_pos,
                                                            pred_eid, blocked, meta, edge); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int nnf_exclusive_scan_u64(const unsigned long long *in, unsigned long long *out, uint32_t n, void *temp,
                           size_t &temp_bytes, cudaStream_t stream) {
  PRGPU_CUDA_TRY(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, (int)n, stream)); if (temp) PRGPU_COUNT_LAUNCH(2);
  return PRGPU_OK;
}

int nnf_band_dp_launch(const NnfBandDev &g, uint32_t first_level, cudaStream_t stream, KernelTimer *timer) {
  static int cluster_size = 0;
  if (cluster_size == 0) {
    // 16 CTAs per cluster where the device allows it (non-portable size), else the portable 8
    cluster_size = 8;
    if (cudaFuncSetAttribute(nnf_band_dp_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t probe = {};
      probe.gridDim = dim3(16);
      probe.blockDim = dim3(BAND_THREADS);
      cudaLaunchAttribute pa[1];
      pa[0].id = cudaLaunchAttributeClusterDimension;
      pa[0].val.clusterDim.x = 16;
      pa[0].val.clusterDim.y = 1;
      pa[0].val.clusterDim.z = 1;
      probe.attrs = pa;
      probe.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, nnf_band_dp_kernel, &probe) == cudaSuccess && nclusters >= 1)
        cluster_size = 16;
    }
    cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster_size);
  cfg.blockDim = dim3(BAND_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (timer)
    timer->begin(stream);
  PRGPU_CUDA_TRY(cudaLaunchKernelEx(&cfg, nnf_band_dp_kernel, g, first_level)); PRGPU_COUNT_LAUNCH(1);
  if (timer)
    timer->end(stream);
  return PRGPU_OK;
}

int nnf_flow_dp_launch(const NnfBandDev &g, uint32_t first_pos, uint32_t epoch, uint32_t *done, int sm_count,
                       cudaStream_t stream, KernelTimer *timer) {
  int per_sm = 0;
  PRGPU_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nnf_flow_dp_kernel, FLOW_THREADS, 0));
  if (per_sm < 1)
    return fail(PRGPU_ECUDA, "nnf_flow_dp_kernel cannot be made resident");
  if (per_sm > 4)
    per_sm = 4;
  dim3 grid((unsigned)(per_sm * sm_count)), block(FLOW_THREADS);
  NnfBandDev arg = g;
  void *args[] = {&arg, &first_pos, &epoch, &done};
  if (timer)
    timer->begin(stream);
  // cooperative launch: guarantees that all blocks are co-resident (warps wait on one another's flags)
  PRGPU_CUDA_TRY(cudaLaunchCooperativeKernel((void *)nnf_flow_dp_kernel, grid, block, args, 0, stream)); PRGPU_COUNT_LAUNCH(1);
  if (timer)
    timer->end(stream);
  return PRGPU_OK;
}

int nnf_band_backtrack_launch(const NnfBandDev &g, uint32_t *out_path, uint32_t cap_pairs, uint32_t *out_n,
                              double *out_cost, cudaStream_t stream) {
  nnf_band_backtrack_kernel<<<1, 32, 0, stream>>>(g, out_path, cap_pairs, out_n, out_cost); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int nnf_backtrack_launch(const NnfGraphDev &g, uint32_t *out_path, uint32_t cap_pairs, uint32_t *out_n,
                         double *out_cost, cudaStream_t stream) {
  nnf_backtrack_kernel<<<1, 32, 0, stream>>>(g, out_path, cap_pairs, out_n, out_cost); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int nnf_lazysp_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const NnfBandDev &g, const NnfLazyDev &z,
                      cudaStream_t stream) {
  const size_t tables = (obstacle_tables_bytes(w) + 15) & ~(size_t)15;
  const size_t smem = tables + (((size_t)g.Lv + 1 + 3) & ~(size_t)3) * sizeof(uint32_t) +
                      (2 * (size_t)LAZY_WL_CAP + LAZY_HASH + LAZY_WAVE_BUDGET) * sizeof(uint32_t);
  if (smem > 200 * 1024)
    return fail(PRGPU_ELIMIT, "nnf_lazysp: obstacle tables + level counters exceed shared memory");
#define NNF_LAZY_CASE(SPV)                                                                         \
  do {                                                                                             \
    auto kern = nnf_lazysp_kernel<SPV>;                                                            \
    PRGPU_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<1, LAZY_THREADS, smem, stream>>>(w_dev, g, z, (uint32_t)tables);                        \
  } while (0)
  if (w.grid_n[0] > 0)
    NNF_LAZY_CASE(0);
  else if (w.S <= 8)
    NNF_LAZY_CASE(8);
  else if (w.S <= 16)
    NNF_LAZY_CASE(16);
  else if (w.S <= 24)
    NNF_LAZY_CASE(24);
  else
    NNF_LAZY_CASE(32);
#undef NNF_LAZY_CASE
  PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

} // namespace prgpu

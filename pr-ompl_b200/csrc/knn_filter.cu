// K1 fast path: TF32 tensor-core scan + fp32 / fp64 re-rank with a proof of exactness.
//
// The exact kernel (knn.cu) spends 20 FP64 instructions per (query, node) pair because the
// reference's distance forbids FMA and runs on the half-rate FP64 pipe.  For the batched shapes
// (>= 128 queries per tree pass, k <= 16) almost all of that work only establishes that a node is
// NOT among the k nearest.  This path does that part at reduced precision with rigorous error bounds:
//   key(q,x) = |x|^2 - 2 q.x  (= |q-x|^2 - |q|^2).
//   1. knn_sample_kernel (fp32 FMA, fp32 mirror of the tree) on every 64th tile: T_q = 8-th smallest
//      key of the sample.  About 8*64 nodes of the whole tree lie below it; the chance that fewer
//      than 16 do is that of a Gamma(8) variate below 1/8, 1.5e-12.
//   2. knn_collect_kernel: one pass over the TF32 mirror on the tensor cores (mma.sync m16n8k8),
//      D = key - T_q for 16 queries x 8 nodes per instruction; nodes with the sign bit of D set are
//      appended to the query's candidate buffer.  No per-query state is updated in the loop.
//   3. knn_select_kernel: one block per query cuts the candidates by their TF32 keys, then by
//      recomputed fp32 keys, recomputes the reference's exact fp64 distance of the few survivors,
//      selects the k first under (distance, index) -- and PROVES that no discarded node can precede
//      them: a discarded node has D >= +0 and D is within eps_tc of the real value, so its exact squared
//      distance is >= T_q - eps_tc + |q|^2.  If the k-th exact squared distance is not strictly below
//      that bound (ties at the cut, heavy duplication, clusters far tighter than the TF32 error) or a
//      buffer overflowed, the query is flagged.
//   4. flagged queries are re-run through the exact kernel.  Results are therefore bit-identical to
//      knn_search_exact for every input.
#include <math_constants.h>
#include <vector>
#include "knn.cuh"

namespace prgpu {
namespace {

constexpr int FT_TILE = 128;     // nodes per tile (32 bytes each)
constexpr int FT_THREADS = 128;  // queries per block


// Pass 1, fp32: the KP smallest keys of every query over a strided sub-sample of the tree (the 32x
// coarser pre-sample of every call, and the sample proper of trees below 4M nodes, seeded with the
// pre-sample's threshold; larger trees take their sample on the tensor cores, see
// knn_sample_threshold_kernel).  One thread =
// one query; the keys live in registers as a sorted array updated by a branch-free insertion network
// (new[e] = min(old[e], max(old[e-1], x))), entered only when some lane of the warp has a key below
// its current KP-th -- no per-lane divergent code, no shared-memory lists.  Node chunks (blockIdx.y)
// emit their sorted keys as partial lists, merged like any other partial top-k tables.
constexpr int PS_KP = KNN_FILTER_SAMPLE_RANK;
template <int DIM>
__global__ void __launch_bounds__(FT_THREADS)
knn_sample_kernel(const float *__restrict__ ptsf, uint32_t n, const double *__restrict__ queries, uint32_t nq,
                  int kp, const uint32_t *__restrict__ lo, uint32_t tile_stride, uint32_t tiles_per_chunk,
                  const float *__restrict__ seed, double *__restrict__ out_key, uint32_t *__restrict__ out_idx) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *tile = reinterpret_cast<float4 *>(smem_raw);                     // [2][FT_TILE][2]
  uint64_t *bars = reinterpret_cast<uint64_t *>(tile + 2 * FT_TILE * 2);   // [2]
  const int tid = threadIdx.x;
  const uint32_t q = blockIdx.x * FT_THREADS + tid;
  const bool act = q < nq;
  float q2[7];
#pragma unroll
  for (int d = 0; d < 7; ++d)
    q2[d] = (act && d < DIM) ? (float)(-2.0 * queries[(size_t)q * DIM + d]) : 0.f;
  const uint32_t my_lo = (act && lo) ? lo[q] : 0u;
  // seeding every slot with a known upper bound of the kp-th smallest key keeps the result an upper
  // bound and spares the warm-up flood
  const float s0 = (act && seed) ? seed[q] : CUDART_INF_F;
  float keys[PS_KP];
#pragma unroll
  for (int e = 0; e < PS_KP; ++e)
    keys[e] = s0;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t ntiles = (n + FT_TILE - 1) / FT_TILE;
  const uint32_t nsamp = (ntiles + tile_stride - 1) / tile_stride;
  const uint32_t t0 = blockIdx.y * tiles_per_chunk;
  const uint32_t t1 = min(t0 + tiles_per_chunk, nsamp);
  auto issue = [&](uint32_t t, int buf) {
    mbar_expect_tx(&bars[buf], FT_TILE * 32);
    bulk_g2s(tile + buf * FT_TILE * 2, ptsf + (size_t)t * tile_stride * FT_TILE * 8, FT_TILE * 32, &bars[buf]);
  };
  if (tid == 0 && t0 < t1)
    issue(t0, 0);
  for (uint32_t t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    if (tid == 0 && t + 1 < t1)
      issue(t + 1, buf ^ 1);
    mbar_wait(&bars[buf], ((t - t0) >> 1) & 1);
    const float4 *T = tile + buf * FT_TILE * 2;
    const uint32_t base = t * tile_stride * FT_TILE;
    const uint32_t jmax = min((uint32_t)FT_TILE, n - base);
    for (uint32_t j = 0; j < jmax; ++j) {
      const float4 a = T[2 * j], b = T[2 * j + 1];
      float x = b.w;
      x = fmaf(q2[0], a.x, x);
      if (DIM > 1) x = fmaf(q2[1], a.y, x);
      if (DIM > 2) x = fmaf(q2[2], a.z, x);
      if (DIM > 3) x = fmaf(q2[3], a.w, x);
      if (DIM > 4) x = fmaf(q2[4], b.x, x);
      if (DIM > 5) x = fmaf(q2[5], b.y, x);
      if (DIM > 6) x = fmaf(q2[6], b.z, x);
      if (!act || base + j < my_lo)
        x = CUDART_INF_F;
      if (__any_sync(0xffffffffu, x < keys[PS_KP - 1])) {
#pragma unroll
        for (int e = PS_KP - 1; e >= 1; --e)
          keys[e] = fminf(keys[e], fmaxf(keys[e - 1], x));
        keys[0] = fminf(keys[0], x);
      }
    }
    __syncthreads();
  }
  if (act) { // partial list (chunk, query): kp keys ascending; the index only has to be unique
    double *od = out_key + ((size_t)blockIdx.y * nq + q) * kp;
    uint32_t *oi = out_idx + ((size_t)blockIdx.y * nq + q) * kp;
#pragma unroll
    for (int e = 0; e < PS_KP; ++e)
      if (e < kp) {
        od[e] = keys[e] < CUDART_INF_F ? (double)keys[e] : CUDART_INF;
        oi[e] = keys[e] < CUDART_INF_F ? blockIdx.y * (uint32_t)kp + (uint32_t)e : PRGPU_INVALID_INDEX;
      }
  }
}

// Pass 2: full scan on the tensor cores.  For query q and node x the scan evaluates
//     D(q, x) = |x|^2 - 2 q.x - T_q
// as one 16x8x8 TF32 mma.sync per 16 queries x 8 nodes: A = 16 queries x 8 slots (-2 q_d rounded to
// TF32 in slots 0..6, -T_q in slot 7), B = 8 slots x 8 nodes (the TF32 mirror record, 1.0 in slot 7),
// C = |x|^2 in fp32.  A node is a candidate of q iff the sign bit of D is set and node >= lo_q; the
// candidate's D is stored as its key (same order as the key itself for one query).  The hot loop is
// 2 LDS.64 + 4 MMA + 8 LOP3 + 1 branch per 64 queries x 8 nodes and holds no per-query state besides
// the A fragments.  The contraction is only 7 long, so this is not a GEMM worth the tcgen05/TMEM path:
// the output (one value per pair, consumed in place) is the bottleneck, TMEM read-out runs at 16
// values/clk/SM, while the warp-level MMA leaves D in registers at 60 values/clk/SM (measured,
// scripts/ubench/mma_tf32.cu) against 18/clk/SM for 7 FFMA per pair.
//
// Error of D against the real |x|^2 - 2 q.x - T_q (eps_tc in knn_select_kernel): both factors of every
// product are within 2^-11 (1 + 2^-12) relative of the fp64 values (fp32 then TF32 rounding), products
// of TF32 numbers are exact, so the dot product is off by at most 2^-9 (1 + 2^-10) |q| |x|
// (Cauchy-Schwarz on 2 |q_d| |x_d| 2^-10); the accumulation of 8 products and C in the tensor core
// (>= fp32 precision, truncating) is bounded generously by 2^-18 ((|x| + |q|)^2 + |T_q|).
#ifndef CT_TILE_N
#define CT_TILE_N 256
#endif
#ifndef CT_STAGES_N
#define CT_STAGES_N 2
#endif
constexpr int CT_TILE = CT_TILE_N;      // nodes per tile of the collect kernel (32 groups of 8: one history word)
constexpr int CT_STAGES = CT_STAGES_N;      // tiles in flight per block
constexpr int CT_MT = 4;          // 16-query MMA row blocks per warp
constexpr int CT_QPB = (FT_THREADS / 32) * CT_MT * 16; // queries per block (256)
#ifndef CT_MINB
#define CT_MINB 8 // resident blocks per SM the collect kernel is compiled for (64 registers)
#endif
constexpr uint32_t CT_WAVES = 32; // grid size of the collect kernel in units of resident blocks
constexpr uint32_t CT_MIN_TILES = 12; // tiles per block below which the block prologue shows

__device__ __forceinline__ uint32_t tf32_bits(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return u;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], float2 b, const float4 &c) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)), "f"(c.x),
        "f"(c.y), "f"(c.z), "f"(c.w));
}

template <int DIM>
__global__ void __launch_bounds__(FT_THREADS, CT_MINB)
knn_collect_kernel(const float *__restrict__ ptst, const float *__restrict__ n2f, uint32_t n,
                   const double *__restrict__ queries, uint32_t nq, const uint32_t *__restrict__ lo,
                   const float *__restrict__ thresh, uint32_t tile_stride, uint32_t tiles_per_chunk, uint32_t cap,
                   uint32_t *__restrict__ cand_count, uint32_t *__restrict__ cand_idx, float *__restrict__ cand_key) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *recs = reinterpret_cast<float *>(smem_raw);                        // [CT_STAGES][CT_TILE][8]   TF32 records
  float *n2s = recs + CT_STAGES * CT_TILE * 8;                              // [CT_STAGES][CT_TILE/2][4] |x|^2 pairs
  uint64_t *full = reinterpret_cast<uint64_t *>(n2s + CT_STAGES * CT_TILE * 2); // [CT_STAGES] tile landed
  uint64_t *empty = full + CT_STAGES;                                       // [CT_STAGES] tile consumed by every warp
  __shared__ uint32_t s_lo_min;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const uint32_t qbase = blockIdx.x * CT_QPB + warp * (CT_MT * 16);
  // A fragments.  MMA slot s <-> coordinate: slot t <-> 2t, slot t+4 <-> 2t+1, so that the B fragment of
  // a lane (slots t and t+4 of node g) is one 8-byte word of the node record.
  auto aval = [&](uint32_t row, int d) -> uint32_t {
    if (row >= nq)
      return 0u; // D = |x|^2 >= +0: never a candidate
    if (d < DIM)
      return tf32_bits((float)(-2.0 * queries[(size_t)row * DIM + d]));
    if (d == 7)
      return __float_as_uint(-thresh[row]); // already a TF32 number (knn_threshold_kernel)
    return 0u;
  };
  uint32_t a[CT_MT][4];
#pragma unroll
  for (int mt = 0; mt < CT_MT; ++mt) {
    const uint32_t r0 = qbase + mt * 16 + g, r1 = r0 + 8;
    a[mt][0] = aval(r0, 2 * t4);
    a[mt][1] = aval(r1, 2 * t4);
    a[mt][2] = aval(r0, 2 * t4 + 1);
    a[mt][3] = aval(r1, 2 * t4 + 1);
  }
  if (tid == 0) {
    s_lo_min = lo ? 0xFFFFFFFFu : 0u;
    for (int i = 0; i < CT_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], FT_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (lo) {
    for (uint32_t r = blockIdx.x * CT_QPB + tid; r < min(nq, (blockIdx.x + 1) * CT_QPB); r += FT_THREADS)
      atomicMin(&s_lo_min, lo[r]);
    __syncthreads();
  }
  // the scan visits every tile_stride-th tile (1: the whole tree; > 1: the sample pass)
  const uint32_t ntiles = ((n + CT_TILE - 1) / CT_TILE + tile_stride - 1) / tile_stride;
  uint32_t t0 = blockIdx.y * tiles_per_chunk;
  const uint32_t t1 = min(t0 + tiles_per_chunk, ntiles);
  if (s_lo_min != 0xFFFFFFFFu)
    t0 = max(t0, s_lo_min / CT_TILE / tile_stride);
  const uint32_t ntl = t1 > t0 ? t1 - t0 : 0;
  // A ring of CT_STAGES tiles with full/empty barriers and no block-wide barrier: the warps of a block
  // drift up to CT_STAGES - 1 tiles apart (the rare emit paths make their pace uneven).
  auto issue = [&](uint32_t i) { // local tile i -> slot i % CT_STAGES
    const uint32_t slot = i % CT_STAGES, use = i / CT_STAGES;
    if (use > 0)
      mbar_wait(&empty[slot], (use - 1) & 1);
    mbar_expect_tx(&full[slot], CT_TILE * 40);
    const size_t tile = (size_t)(t0 + i) * tile_stride;
    bulk_g2s(recs + slot * CT_TILE * 8, ptst + tile * CT_TILE * 8, CT_TILE * 32, &full[slot]);
    bulk_g2s(n2s + slot * CT_TILE * 2, n2f + tile * CT_TILE * 2, CT_TILE * 8, &full[slot]);
  };
  auto emit = [&](uint32_t row, uint32_t gi, float dv) {
    if (gi < n && row < nq && (!lo || gi >= lo[row])) {
      const uint32_t slot = atomicAdd(cand_count + row, 1u);
      if (slot < cap) {
        cand_idx[(size_t)row * cap + slot] = gi;
        cand_key[(size_t)row * cap + slot] = dv;
      }
    }
  };
  if (tid == 0)
    for (uint32_t i = 0; i + 1 < CT_STAGES && i < ntl; ++i)
      issue(i);
  for (uint32_t i = 0; i < ntl; ++i) {
    if (tid == 0 && i + CT_STAGES - 1 < ntl)
      issue(i + CT_STAGES - 1);
    const uint32_t slot = i % CT_STAGES;
    mbar_wait(&full[slot], (i / CT_STAGES) & 1);
    const float2 *R = reinterpret_cast<const float2 *>(recs + slot * CT_TILE * 8) + g * 4 + t4;
    const float4 *N2 = reinterpret_cast<const float4 *>(n2s + slot * CT_TILE * 2) + t4;
    const uint32_t base = (t0 + i) * tile_stride * CT_TILE;
    auto load_mma = [&](int nt, float (&d)[CT_MT][4]) {
      const float2 b = R[nt * 32];  // node nt*8+g, slots (t, t+4)
      const float4 c = N2[nt * 4];  // |x|^2 of nodes nt*8+2t, nt*8+2t+1 (the lane's two output columns), twice
#pragma unroll
      for (int mt = 0; mt < CT_MT; ++mt)
        mma_tf32_16x8x8(d[mt], a[mt], b, c);
    };
    // hot loop, straight-line: the sign bits of the 16 outputs of a lane are OR-ed and shifted into a
    // 32-bit history (one bit per 8-node group); nothing else is decided here
    uint32_t hist = 0;
#pragma unroll
    for (int nt = 0; nt < CT_TILE / 8; ++nt) {
      float d[CT_MT][4];
      load_mma(nt, d);
      uint32_t m = 0;
#pragma unroll
      for (int mt = 0; mt < CT_MT; ++mt)
        m |= __float_as_uint(d[mt][0]) | __float_as_uint(d[mt][1]) | __float_as_uint(d[mt][2]) |
             __float_as_uint(d[mt][3]);
      hist = __funnelshift_l(m, hist, 1); // (hist << 1) | sign(m): group nt ends at bit 31 - nt
    }
    // rare path: the 8-node groups in which some lane saw a candidate (about 3 %) are evaluated again
    // and their candidates appended
    uint32_t todo = __reduce_or_sync(0xffffffffu, hist);
    while (todo) {
      const int nt = __clz(todo);
      todo &= ~(0x80000000u >> nt);
      float d[CT_MT][4];
      load_mma(nt, d);
#pragma unroll
      for (int mt = 0; mt < CT_MT; ++mt)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (__float_as_uint(d[mt][j]) & 0x80000000u)
            emit(qbase + mt * 16 + g + (j >> 1) * 8, base + nt * 8 + 2 * t4 + (j & 1), d[mt][j]);
    }
    __syncwarp();
    if (lane == 0)
      mbar_arrive(&empty[slot]);
  }
}

// smallest TF32 number >= v (v finite or +inf)
__device__ __forceinline__ float tf32_round_up(float v) {
  uint32_t b = __float_as_uint(v);
  if ((b & 0x7F800000u) == 0x7F800000u)
    return v;
  b = (b & 0x80000000u) ? (b & ~0x1FFFu) : ((b + 0x1FFFu) & ~0x1FFFu);
  return __uint_as_float(b);
}

// threshold of every query = kp-th smallest key of the sample (last entry of its merged list)
__global__ void knn_threshold_kernel(const double *__restrict__ cand_key, const uint32_t *__restrict__ cand_idx,
                                     uint32_t nq, int kp, const float *coarse, float *thresh, uint32_t *counts) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq)
    return;
  const bool full = cand_idx[(size_t)q * kp + kp - 1] != PRGPU_INVALID_INDEX;
  // list not full: fewer than kp sampled nodes lie below the coarser bound (or no bound: keep all)
  const float t = full ? (float)cand_key[(size_t)q * kp + kp - 1] : (coarse ? coarse[q] : CUDART_INF_F);
  thresh[q] = tf32_round_up(t); // the scan multiplies it on the tensor cores: must be a TF32 number
  if (counts)
    counts[q] = 0;
}

// Sample pass on the tensor cores, second half: the scan over the sample tiles (knn_collect_kernel with a
// tile stride and the pre-sample's coarse threshold T_c) left D = key - T_c of the sample nodes below T_c
// in the candidate buffers; one warp per query takes the kp-th smallest of them: T_q = T_c + D_kp.
// Any value is a legal threshold (exactness never depends on it): with fewer than kp entries T_c stays.
__global__ void knn_sample_threshold_kernel(uint32_t nq, int kp, uint32_t cap, uint32_t *__restrict__ counts,
                                            const float *__restrict__ ckey, float *__restrict__ thresh) {
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= nq)
    return;
  const uint32_t cnt = min(counts[q], cap);
  const float *v = ckey + (size_t)q * cap;
  float pv = -CUDART_INF_F; // (value, position) of the previous round's pick
  uint32_t pc = 0;
  bool first = true;
  for (int r = 0; r < kp && (uint32_t)r < cnt; ++r) {
    float bv = CUDART_INF_F;
    uint32_t bc = 0xFFFFFFFFu;
    for (uint32_t c = lane; c < cnt; c += 32) {
      const float x = v[c];
      const bool after = first || x > pv || (x == pv && c > pc);
      if (after && (x < bv || (x == bv && c < bc))) {
        bv = x;
        bc = c;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const uint32_t oc = __shfl_xor_sync(0xffffffffu, bc, off);
      if (ov < bv || (ov == bv && oc < bc)) {
        bv = ov;
        bc = oc;
      }
    }
    pv = bv;
    pc = bc;
    first = false;
  }
  __syncwarp();
  if (lane == 0) {
    if (cnt >= (uint32_t)kp && pv < CUDART_INF_F && pv > -CUDART_INF_F)
      thresh[q] = tf32_round_up(thresh[q] + pv);
    counts[q] = 0;
  }
}

// Pass 3: one block per query, three levels.  Selections are done by RANK (every thread counts the entries
// that precede its own; no rounds of block-wide arg-min, two barriers per level).
//  (a) on the stored TF32 keys D: kk = the k-th smallest of the 128 per-thread minima, an upper bound of the
//      k-th smallest D.  Every D is within eps_tc of the real value, so a candidate with D > kk + 2 eps_tc
//      has a real key above that of each of k other candidates and cannot be among the k nearest: the others
//      (a few dozen) are the level-1 survivors;
//  (b) their keys are recomputed with the fp32 FMA chain of the sample pass (error eps_32, a thousand
//      times smaller) and cut again at (k-th smallest) + 2 eps_32: usually k plus a few remain;
//  (c) exact fp64 distance (the reference's operation order) of those; the survivor of rank r under
//      (distance, index) is output r; then the proof against the nodes the scan discarded (D >= +0): their
//      exact squared distance is >= T_q - eps_tc + |q|^2, which must lie strictly above the k-th selected.
// Unproven queries, overflowed candidate buffers and over-long survivor lists are flagged for the exact
// kernel.
// the two error bounds the selection rests on (derivation: DESIGN.md §4 K1 step 3; pinned on the CPU by
// tests/test_filter_bounds.py and on the device by prgpu_diag_knn_scan / tests/test_gpu_bounds.py)
__device__ __forceinline__ void filter_error_bounds(double xmax2, double qn2, float T, double &eps_tc,
                                                             double &eps32) {
  const double xn = sqrt(xmax2) * (1.0 + 1e-12), qn = sqrt(qn2) * (1.0 + 1e-12);
  const double u24 = 5.9604644775390625e-08;
  eps32 = 16.0 * u24 * (xn + qn) * (xn + qn) * (1.0 + 1e-6) + 1e-30;
  const double at = T < CUDART_INF_F ? fabs((double)T) : 0.0;
  eps_tc = 1.01 * 0.001953125 * qn * xn + 3.814697265625e-06 * ((xn + qn) * (xn + qn) + at) + 1e-30;
}
// fp32 key |x|^2 - 2 q.x of node i by the FMA chain of the sample pass (q2[d] = fp32(-2 q_d))
template <int DIM> __device__ __forceinline__ float key32_chain(const float *__restrict__ ptsf, uint32_t i, const float (&q2)[7]) {
  const float4 *rec = reinterpret_cast<const float4 *>(ptsf + (size_t)i * 8);
  const float4 a = rec[0], b = rec[1];
  float x = b.w;
  x = fmaf(q2[0], a.x, x);
  if (DIM > 1) x = fmaf(q2[1], a.y, x);
  if (DIM > 2) x = fmaf(q2[2], a.z, x);
  if (DIM > 3) x = fmaf(q2[3], a.w, x);
  if (DIM > 4) x = fmaf(q2[4], b.x, x);
  if (DIM > 5) x = fmaf(q2[5], b.y, x);
  if (DIM > 6) x = fmaf(q2[6], b.z, x);
  return x;
}

constexpr int SEL_THREADS = 128;
constexpr int SEL_MAX_L1 = 512;        // level-1 survivors
constexpr int SEL_MAX_SURVIVORS = 64;  // level-2 survivors
template <int DIM>
__global__ void __launch_bounds__(SEL_THREADS)
knn_select_kernel(const double *__restrict__ pts, const float *__restrict__ ptsf, uint64_t stride, uint32_t index_base,
                  const double *__restrict__ xmax2, const double *__restrict__ queries, const float *__restrict__ thresh,
                  uint32_t cap, const uint32_t *__restrict__ cand_count, const uint32_t *__restrict__ cand_idx,
                  const float *__restrict__ cand_key, int k, uint32_t *__restrict__ idx_out,
                  double *__restrict__ dist_out, uint32_t *__restrict__ flags) {
  __shared__ float tmin[SEL_THREADS];
  __shared__ uint32_t l1_i[SEL_MAX_L1];
  __shared__ float l1_k[SEL_MAX_L1];
  __shared__ double sv_d[SEL_MAX_SURVIVORS];
  __shared__ uint32_t sv_i[SEL_MAX_SURVIVORS];
  __shared__ uint32_t n_l1, n_sv;
  __shared__ float s_kk;
  __shared__ double s_eps_tc, s_eps32, s_qn2, s_dk;
  const uint32_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const uint32_t cnt = cand_count[q];
  auto flag_me = [&]() {
    if (tid == 0) {
      const uint32_t slot = atomicAdd(flags, 1u);
      flags[1 + slot] = q;
    }
  };
  if (cnt > cap) { // buffer overflow: cannot be decided here
    flag_me();
    return;
  }
  double qx[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d)
    qx[d] = queries[(size_t)q * DIM + d];
  const float T = thresh[q];
  if (tid == 0) {
    double qn2 = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
      qn2 += qx[d] * qx[d];
    double etc, e32;
    filter_error_bounds(*xmax2, qn2, T, etc, e32);
    s_eps_tc = etc;
    s_eps32 = e32;
    s_qn2 = qn2;
    n_l1 = 0;
    n_sv = 0;
    s_kk = CUDART_INF_F;
    s_dk = CUDART_INF;
  }
  const float *ckey = cand_key + (size_t)q * cap;
  const uint32_t *cidx = cand_idx + (size_t)q * cap;
  // (a) an upper bound of the k-th smallest stored key: the k-th smallest of the per-thread minima
  {
    float mine = CUDART_INF_F;
    for (uint32_t c = tid; c < cnt; c += SEL_THREADS)
      mine = fminf(mine, ckey[c]);
    tmin[tid] = mine;
  }
  __syncthreads();
  const uint32_t nmin = min(cnt, (uint32_t)SEL_THREADS); // threads that hold a candidate
  if (cnt > (uint32_t)k && (uint32_t)tid < nmin) {
    const float mine = tmin[tid];
    uint32_t rank = 0;
    for (uint32_t u = 0; u < nmin; ++u) {
      const float v = tmin[u];
      rank += (v < mine || (v == mine && u < (uint32_t)tid)) ? 1u : 0u;
    }
    if (rank == (uint32_t)k - 1u)
      s_kk = mine;
  }
  __syncthreads();
  // level-1 survivors: everything whose key is within 2 eps_tc of that bound (all, if no more than k)
  const float cut1 =
      cnt <= (uint32_t)k ? CUDART_INF_F : (float)((double)s_kk + 2.0 * s_eps_tc) * (1.0f + 1e-6f) + 1e-30f;
  for (uint32_t c = tid; c < cnt; c += SEL_THREADS)
    if (ckey[c] <= cut1) {
      const uint32_t slot = atomicAdd(&n_l1, 1u);
      if (slot < SEL_MAX_L1)
        l1_i[slot] = cidx[c];
    }
  __syncthreads();
  const uint32_t n1 = n_l1;
  if (n1 > SEL_MAX_L1) {
    flag_me();
    return;
  }
  // (b) fp32 keys of the level-1 survivors (the chain of knn_sample_kernel), cut again
  {
    float q2[7];
#pragma unroll
    for (int d = 0; d < 7; ++d)
      q2[d] = d < DIM ? (float)(-2.0 * qx[d < DIM ? d : 0]) : 0.f;
    for (uint32_t c = tid; c < n1; c += SEL_THREADS)
      l1_k[c] = key32_chain<DIM>(ptsf, l1_i[c], q2);
  }
  if (tid == 0)
    s_kk = CUDART_INF_F;
  __syncthreads();
  if (n1 > (uint32_t)k)
    for (uint32_t c = tid; c < n1; c += SEL_THREADS) { // the entry of rank k-1 is the k-th smallest fp32 key
      const float mine = l1_k[c];
      uint32_t rank = 0;
      for (uint32_t u = 0; u < n1; ++u) {
        const float v = l1_k[u];
        rank += (v < mine || (v == mine && u < c)) ? 1u : 0u;
      }
      if (rank == (uint32_t)k - 1u)
        s_kk = mine;
    }
  __syncthreads();
  const float cut2 =
      n1 <= (uint32_t)k ? CUDART_INF_F : (float)((double)s_kk + 2.0 * s_eps32) * (1.0f + 1e-6f) + 1e-30f;
  for (uint32_t c = tid; c < n1; c += SEL_THREADS)
    if (l1_k[c] <= cut2) {
      const uint32_t slot = atomicAdd(&n_sv, 1u);
      if (slot < SEL_MAX_SURVIVORS)
        sv_i[slot] = l1_i[c];
    }
  __syncthreads();
  const uint32_t ns = n_sv;
  if (ns > SEL_MAX_SURVIVORS) { // massive ties around the k-th key
    flag_me();
    return;
  }
  // (c) exact distances of the survivors; output r = the survivor of rank r under (distance, index)
  if ((uint32_t)tid < ns) {
    const uint32_t i = sv_i[tid];
    double x[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
      x[d] = pts[(size_t)d * stride + i];
    sv_d[tid] = __dsqrt_rn(sqdist_exact<DIM>(x, qx)); // RealVectorStateSpace::distance(node, query)
  }
  __syncthreads();
  if ((uint32_t)tid < ns) {
    const double md = sv_d[tid];
    const uint32_t mi = sv_i[tid];
    uint32_t rank = 0;
    for (uint32_t u = 0; u < ns; ++u) {
      const double d = sv_d[u];
      const uint32_t i = sv_i[u];
      rank += (d < md || (d == md && i < mi)) ? 1u : 0u;
    }
    if (rank < (uint32_t)k) {
      dist_out[(size_t)q * k + rank] = md;
      idx_out[(size_t)q * k + rank] = mi + index_base;
      if (rank == (uint32_t)k - 1u)
        s_dk = md;
    }
  }
  for (uint32_t r = ns + tid; r < (uint32_t)k; r += SEL_THREADS) { // fewer than k eligible nodes: padding
    dist_out[(size_t)q * k + r] = CUDART_INF;
    idx_out[(size_t)q * k + r] = PRGPU_INVALID_INDEX;
  }
  __syncthreads();
  if (tid == 0 && T < CUDART_INF_F) { // nodes were discarded by the scan: prove they cannot matter
    const double dk = s_dk;
    const double bound = (double)T - s_eps_tc + s_qn2; // lower bound of any discarded node's exact squared distance
    const double sk = dk * dk * (1.0 + 1e-12) + 1e-12 * (s_qn2 + 1.0);
    if (!(sk < bound)) {
      const uint32_t slot = atomicAdd(flags, 1u);
      flags[1 + slot] = q;
    }
  }
}

__global__ void gather_queries_kernel(const double *__restrict__ queries, const uint32_t *__restrict__ lo, int dim,
                                      const uint32_t *__restrict__ ids, uint32_t n, double *__restrict__ out_q,
                                      uint32_t *__restrict__ out_lo) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t q = ids[i];
  for (int d = 0; d < dim; ++d)
    out_q[(size_t)i * dim + d] = queries[(size_t)q * dim + d];
  if (out_lo)
    out_lo[i] = lo ? lo[q] : 0u;
}
__global__ void scatter_rows_kernel(const uint32_t *__restrict__ ids, uint32_t n, int k, const uint32_t *__restrict__ si,
                                    const double *__restrict__ sd, uint32_t *__restrict__ idx_out,
                                    double *__restrict__ dist_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * (uint32_t)k)
    return;
  const uint32_t row = i / k, e = i % k, q = ids[row];
  idx_out[(size_t)q * k + e] = si[i];
  dist_out[(size_t)q * k + e] = sd[i];
}

constexpr uint32_t FT_SAMPLE_STRIDE = 64; // the sample pass visits every 64th tile (every 32nd of a tree below 4M nodes)

// re-runs the `nflag` unproven queries listed in sc.flags[1..] with the exact kernel
template <int DIM>
int filter_fallback(const KnnTreeView &t, const double *queries, uint32_t nq, int k, const uint32_t *lo,
                    uint32_t *idx_out, double *dist_out, KnnScratch &sc, int sm_count, cudaStream_t stream,
                    uint32_t nflag) {
  (void)nq;
  if (nflag == 0)
    return PRGPU_OK;
  sc.fallback_queries += nflag;
  if (sc.fb_capacity < nflag) {
    for (void *p : {(void *)sc.fb_q, (void *)sc.fb_dist, (void *)sc.fb_idx, (void *)sc.fb_lo})
      if (p)
        cudaFree(p);
    sc.fb_q = sc.fb_dist = nullptr;
    sc.fb_idx = sc.fb_lo = nullptr;
    sc.fb_capacity = 0;
    const size_t cap = (size_t)nflag * 2;
    PRGPU_CUDA_TRY(cudaMalloc(&sc.fb_q, cap * DIM * sizeof(double)));
    PRGPU_CUDA_TRY(cudaMalloc(&sc.fb_dist, cap * KNN_FILTER_MAX_K * sizeof(double)));
    PRGPU_CUDA_TRY(cudaMalloc(&sc.fb_idx, cap * KNN_FILTER_MAX_K * sizeof(uint32_t)));
    PRGPU_CUDA_TRY(cudaMalloc(&sc.fb_lo, cap * sizeof(uint32_t)));
    sc.fb_capacity = cap;
  }
  gather_queries_kernel<<<(nflag + 127) / 128, 128, 0, stream>>>(queries, lo, DIM, sc.flags + 1, nflag, sc.fb_q,
                                                                sc.fb_lo); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  int rc = knn_search_exact(DIM, t, sc.fb_q, nflag, k, lo ? sc.fb_lo : nullptr, sc.fb_idx, sc.fb_dist, sc, sm_count,
                            stream, nullptr);
  if (rc != PRGPU_OK)
    return rc;
  scatter_rows_kernel<<<(nflag * k + 127) / 128, 128, 0, stream>>>(sc.flags + 1, nflag, k, sc.fb_idx, sc.fb_dist,
                                                                  idx_out, dist_out); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

template <int DIM>
int filter_path(const KnnTreeView &t, const double *queries, uint32_t nq, int k, const uint32_t *lo,
                uint32_t *idx_out, double *dist_out, KnnScratch &sc, int sm_count, cudaStream_t stream,
                KernelTimer *timer, bool *deferred) {
  const int kp = KNN_FILTER_SAMPLE_RANK;
  const uint32_t ntiles = (t.n + FT_TILE - 1) / FT_TILE;
  const uint32_t qblocks = (nq + CT_QPB - 1) / CT_QPB;
  // candidate buffer per query: the sample's kp-th smallest key admits ~kp*stride nodes of the full tree
  uint32_t cap = 1;
  // A candidate costs the scan a replayed 8-node group, an atomic and two stores, and the select pass a
  // key: about kp * stride of them per query whatever the tree size, so small trees (and the shards of a
  // multi-GPU sweep) take a denser sample -- its cost shrinks with the tree.  P[fewer than 16 nodes
  // below T_q] = P[Gamma(8) < 16/stride]: 1.5e-12 at stride 64, 1e-7 at 32.
  const uint32_t sample_stride = t.n >= (1u << 22) ? FT_SAMPLE_STRIDE : FT_SAMPLE_STRIDE / 2;
  while (cap < 3u * (uint32_t)kp * sample_stride)
    cap <<= 1;
  if (cap > 4096)
    cap = 4096;

  // ---- pass 1: thresholds from strided sub-samples (pre-sample, then the sample seeded with it) ----
  const size_t smem1 = (size_t)2 * FT_TILE * 32 + 2 * sizeof(uint64_t);
  auto samp = knn_sample_kernel<DIM>;
  const uint32_t qblocks1 = (nq + FT_THREADS - 1) / FT_THREADS;
  // enough node chunks to fill the machine with one-thread-per-query blocks (the partial lists are merged)
  uint32_t chunks_max = (uint32_t)(16 * sm_count) / qblocks1;
  chunks_max = chunks_max < 1 ? 1 : (chunks_max > 128 ? 128 : chunks_max);
  int rc = scratch_reserve(sc.cand_dist, sc.cand_idx, sc.cand_capacity, (size_t)nq * kp);
  if (rc != PRGPU_OK)
    return rc;
  rc = scratch_reserve(sc.part_dist, sc.part_idx, sc.part_capacity, (size_t)chunks_max * nq * kp);
  if (rc != PRGPU_OK)
    return rc;
  // flags | thresholds | counts | candidate indices | candidate keys
  const size_t need_words = ((size_t)nq + 1) + (size_t)nq + (size_t)nq + 2 * (size_t)nq * cap;
  if (sc.flags_capacity < need_words) {
    if (sc.flags)
      cudaFree(sc.flags);
    sc.flags = nullptr;
    sc.flags_capacity = 0;
    PRGPU_CUDA_TRY(cudaMalloc(&sc.flags, need_words * sizeof(uint32_t)));
    sc.flags_capacity = need_words;
  }
  float *thresh = reinterpret_cast<float *>(sc.flags + (size_t)nq + 1);
  uint32_t *counts = sc.flags + 2 * (size_t)nq + 1;
  uint32_t *cidx = sc.flags + 3 * (size_t)nq + 1;
  float *ckey = reinterpret_cast<float *>(cidx + (size_t)nq * cap);
  PRGPU_CUDA_TRY(cudaMemsetAsync(sc.flags, 0, sizeof(uint32_t), stream));
  // one sampling pass over every `stride`-th tile: kp-th smallest key of each query -> `out`
  auto sample_pass = [&](uint32_t stride, const float *seed, float *out, uint32_t *zero_counts) -> int {
    const uint32_t nsamp = (ntiles + stride - 1) / stride;
    uint32_t ch = chunks_max > nsamp ? nsamp : chunks_max;
    const uint32_t tpc = (nsamp + ch - 1) / ch;
    ch = (nsamp + tpc - 1) / tpc;
    samp<<<dim3(qblocks1, ch), FT_THREADS, smem1, stream>>>(t.ptsf, t.n, queries, nq, kp, lo, stride, tpc, seed,
                                                         ch > 1 ? sc.part_dist : sc.cand_dist,
                                                         ch > 1 ? sc.part_idx : sc.cand_idx); PRGPU_COUNT_LAUNCH(1);
    PRGPU_CUDA_TRY(cudaGetLastError());
    if (ch > 1) {
      const int mrc = merge_passes(sc, (int)ch, nq, kp, sc.cand_idx, sc.cand_dist, stream);
      if (mrc != PRGPU_OK)
        return mrc;
    }
    knn_threshold_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(sc.cand_dist, sc.cand_idx, nq, kp, seed, out, zero_counts); PRGPU_COUNT_LAUNCH(1);
    PRGPU_CUDA_TRY(cudaGetLastError());
    return PRGPU_OK;
  };
  // pre-sample (fp32 FMA, every (32 * stride)-th 128-node tile -- a subset of the sample below): its kp-th
  // smallest key T_c admits about 32 * kp nodes of the sample
  rc = sample_pass(sample_stride * 32, nullptr, thresh, counts);
  if (rc != PRGPU_OK)
    return rc;

  // ---- the scan kernel, used twice: over every sample_stride-th tile with T_c (sample pass), then over
  //      the whole tree with T_q (pass 2) ----
  const size_t smem2 = (size_t)CT_STAGES * CT_TILE * 40 + 2 * CT_STAGES * sizeof(uint64_t);
  auto coll = knn_collect_kernel<DIM>;
  int per_sm = 0;
  PRGPU_CUDA_TRY(cudaFuncSetAttribute(coll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  PRGPU_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coll, FT_THREADS, smem2));
  per_sm = per_sm > 0 ? per_sm : 1;
  auto scan = [&](uint32_t tile_stride, KernelTimer *tm) -> int {
    const uint32_t ntiles2 = ((t.n + CT_TILE - 1) / CT_TILE + tile_stride - 1) / tile_stride;
    uint32_t chunks = (uint32_t)(per_sm * sm_count) / qblocks;
    chunks = chunks < 1 ? 1 : chunks;
    // many short blocks instead of one wave of long ones: the blocks resident on an SM do not advance at
    // the same rate, and a one-wave grid ran its last third at 4.4 of 6 warps per scheduler (ncu r01d)
    const uint32_t one_wave = chunks;
    chunks *= CT_WAVES;
    // ... but a block's prologue (64 queries' A fragments per warp) has to be paid off: at least
    // CT_MIN_TILES tiles per block while that still leaves a full wave
    if (chunks > ntiles2 / CT_MIN_TILES)
      chunks = ntiles2 / CT_MIN_TILES > one_wave ? ntiles2 / CT_MIN_TILES : one_wave;
    if (chunks > ntiles2)
      chunks = ntiles2;
    const uint32_t tiles_per_chunk = (ntiles2 + chunks - 1) / chunks;
    chunks = (ntiles2 + tiles_per_chunk - 1) / tiles_per_chunk;
    if (tm)
      tm->begin(stream);
    coll<<<dim3(qblocks, chunks), FT_THREADS, smem2, stream>>>(t.ptsf + 8 * t.stride, t.ptsf + 16 * t.stride, t.n, queries,
                                                             nq, lo, thresh, tile_stride, tiles_per_chunk, cap, counts,
                                                             cidx, ckey); PRGPU_COUNT_LAUNCH(1);
    if (tm)
      tm->end(stream);
    PRGPU_CUDA_TRY(cudaGetLastError());
    return PRGPU_OK;
  };
  if (t.n >= (1u << 22)) {
    // sample pass on the tensor cores: D = key - T_c of the sample nodes below T_c, then
    // T_q = T_c + (kp-th smallest D)
    if ((rc = scan(sample_stride, nullptr)) != PRGPU_OK)
      return rc;
    knn_sample_threshold_kernel<<<(nq * 32 + 255) / 256, 256, 0, stream>>>(nq, kp, cap, counts, ckey, thresh); PRGPU_COUNT_LAUNCH(1);
    PRGPU_CUDA_TRY(cudaGetLastError());
  } else {
    // small trees: the fp32 sampler seeded with T_c is cheaper than a second strided scan launch
    if ((rc = sample_pass(sample_stride, thresh, thresh, counts)) != PRGPU_OK)
      return rc;
  }
  // ---- pass 2: full scan collecting every node below the threshold ----
  if ((rc = scan(1, timer)) != PRGPU_OK)
    return rc;

  // ---- pass 3: exact re-rank of the candidates + proof ----
  auto sel = knn_select_kernel<DIM>;
  sel<<<nq, SEL_THREADS, 0, stream>>>(t.pts, t.ptsf, t.stride, t.index_base, t.xmax2, queries, thresh, cap, counts,
                                     cidx, ckey, k, idx_out, dist_out, sc.flags); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  ++sc.filter_calls;

  // unproven queries (ties at the cut, heavy duplicates) take the exact kernel
  if (deferred) { // the caller reads flags[0] itself, after whatever it enqueues behind this search
    *deferred = true;
    return PRGPU_OK;
  }
  uint32_t nflag = 0;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(&nflag, sc.flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(stream));
  return filter_fallback<DIM>(t, queries, nq, k, lo, idx_out, dist_out, sc, sm_count, stream, nflag);
}

// ---- measurement aid: the scan's own numbers (prgpu_diag_knn_scan) ----------------------------------------
__global__ void diag_round_thresholds_kernel(float *thresh, uint32_t nq, uint32_t *counts) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq) {
    thresh[q] = tf32_round_up(thresh[q]);
    counts[q] = 0;
  }
}
template <int DIM>
__global__ void diag_keys_kernel(const float *__restrict__ ptsf, const double *__restrict__ xmax2,
                                 const double *__restrict__ queries, const float *__restrict__ thresh, uint32_t nq,
                                 uint32_t cap, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ cidx,
                                 float *__restrict__ key32, double *__restrict__ eps) {
  const uint32_t q = blockIdx.x;
  float q2[7];
  double qn2 = 0.0;
#pragma unroll
  for (int d = 0; d < 7; ++d) {
    const double x = d < DIM ? queries[(size_t)q * DIM + (d < DIM ? d : 0)] : 0.0;
    q2[d] = d < DIM ? (float)(-2.0 * x) : 0.f;
    qn2 += x * x;
  }
  const uint32_t cnt = min(counts[q], cap);
  for (uint32_t c = threadIdx.x; c < cnt; c += blockDim.x)
    key32[(size_t)q * cap + c] = key32_chain<DIM>(ptsf, cidx[(size_t)q * cap + c], q2);
  if (threadIdx.x == 0) {
    double etc, e32;
    filter_error_bounds(*xmax2, qn2, thresh[q], etc, e32);
    eps[2 * (size_t)q] = etc;
    eps[2 * (size_t)q + 1] = e32;
  }
}
template <int DIM>
int diag_scan(const KnnTreeView &t, const double *queries, uint32_t nq, float *thresh, uint32_t cap, uint32_t *counts,
              uint32_t *cidx, float *ckey, float *key32, double *eps, cudaStream_t stream) {
  const uint32_t qblocks = (nq + CT_QPB - 1) / CT_QPB;
  const uint32_t ntiles = (t.n + CT_TILE - 1) / CT_TILE;
  diag_round_thresholds_kernel<<<(nq + 127) / 128, 128, 0, stream>>>(thresh, nq, counts); PRGPU_COUNT_LAUNCH(1);
  const size_t smem2 = (size_t)CT_STAGES * CT_TILE * 40 + 2 * CT_STAGES * sizeof(uint64_t);
  auto coll = knn_collect_kernel<DIM>;
  PRGPU_CUDA_TRY(cudaFuncSetAttribute(coll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  const uint32_t chunks = std::min<uint32_t>(ntiles, 64u);
  const uint32_t tpc = (ntiles + chunks - 1) / chunks;
  coll<<<dim3(qblocks, (ntiles + tpc - 1) / tpc), FT_THREADS, smem2, stream>>>(
      t.ptsf + 8 * t.stride, t.ptsf + 16 * t.stride, t.n, queries, nq, nullptr, thresh, 1u, tpc, cap, counts, cidx, ckey); PRGPU_COUNT_LAUNCH(1);
  diag_keys_kernel<DIM><<<nq, 128, 0, stream>>>(t.ptsf, t.xmax2, queries, thresh, nq, cap, counts, cidx, key32, eps); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}
} // namespace

int knn_diag_scan(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, float *thresh_dev, uint32_t cap,
                  uint32_t *counts_dev, uint32_t *cidx_dev, float *ckey_dev, float *key32_dev, double *eps_dev,
                  cudaStream_t stream) {
  if (!t.ptsf || dim > 7)
    return fail(PRGPU_EINVAL, "diag_knn_scan: the tree has no TF32 mirror (dim <= 7 only)");
  switch (dim) {
#define PRGPU_DIM_CASE(D)                                                                          \
  case D:                                                                                          \
    return diag_scan<D>(t, queries_dev, nq, thresh_dev, cap, counts_dev, cidx_dev, ckey_dev, key32_dev, eps_dev, stream);
    PRGPU_DIM_CASE(1)
    PRGPU_DIM_CASE(2)
    PRGPU_DIM_CASE(3)
    PRGPU_DIM_CASE(4)
    PRGPU_DIM_CASE(5)
    PRGPU_DIM_CASE(6)
    PRGPU_DIM_CASE(7)
#undef PRGPU_DIM_CASE
  }
  return fail(PRGPU_ELIMIT, "diag_knn_scan: dim must be in [1,7]");
}

static bool filter_eligible(int dim, const KnnTreeView &t, uint32_t nq, int k, int force_exact) {
  return !force_exact && t.ptsf != nullptr && dim <= 7 && k >= 1 && k <= KNN_FILTER_MAX_K &&
         nq >= (uint32_t)FT_THREADS && t.n >= 8192;
}

int knn_search_deferred(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, int k,
                        const uint32_t *lo_dev, uint32_t *idx_dev, double *dist_dev, KnnScratch &scratch, int sm_count,
                        cudaStream_t stream, KernelTimer *timer, int force_exact, bool *pending) {
  if (pending)
    *pending = false;
  if (!filter_eligible(dim, t, nq, k, force_exact))
    return knn_search_exact(dim, t, queries_dev, nq, k, lo_dev, idx_dev, dist_dev, scratch, sm_count, stream, timer);
  switch (dim) {
#define PRGPU_DIM_CASE(D)                                                                          \
  case D:                                                                                          \
    return filter_path<D>(t, queries_dev, nq, k, lo_dev, idx_dev, dist_dev, scratch, sm_count, stream, timer, pending);
    PRGPU_DIM_CASE(1)
    PRGPU_DIM_CASE(2)
    PRGPU_DIM_CASE(3)
    PRGPU_DIM_CASE(4)
    PRGPU_DIM_CASE(5)
    PRGPU_DIM_CASE(6)
    PRGPU_DIM_CASE(7)
#undef PRGPU_DIM_CASE
  }
  return fail(PRGPU_ELIMIT, "knn: dim must be in [1,PRGPU_MAX_DIM]");
}

int knn_search_finish(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, int k,
                      const uint32_t *lo_dev, uint32_t *idx_dev, double *dist_dev, KnnScratch &scratch, int sm_count,
                      cudaStream_t stream, uint32_t nflag) {
  switch (dim) {
#define PRGPU_DIM_CASE(D)                                                                          \
  case D:                                                                                          \
    return filter_fallback<D>(t, queries_dev, nq, k, lo_dev, idx_dev, dist_dev, scratch, sm_count, stream, nflag);
    PRGPU_DIM_CASE(1)
    PRGPU_DIM_CASE(2)
    PRGPU_DIM_CASE(3)
    PRGPU_DIM_CASE(4)
    PRGPU_DIM_CASE(5)
    PRGPU_DIM_CASE(6)
    PRGPU_DIM_CASE(7)
#undef PRGPU_DIM_CASE
  }
  return fail(PRGPU_ELIMIT, "knn: dim must be in [1,PRGPU_MAX_DIM]");
}

int knn_search(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, int k, const uint32_t *lo_dev,
               uint32_t *idx_dev, double *dist_dev, KnnScratch &scratch, int sm_count, cudaStream_t stream,
               KernelTimer *timer, int force_exact) {
  return knn_search_deferred(dim, t, queries_dev, nq, k, lo_dev, idx_dev, dist_dev, scratch, sm_count, stream, timer,
                             force_exact, nullptr);
}

} // namespace prgpu

// K1: exact brute-force k-nearest-neighbour search over a tree's configuration buffer.
//
// Replaces ompl::NearestNeighbors<Motion*>::nearest / nearestK as used by the reference
// (pr_ompl/src/RRTConnect.cpp:181) and findKnnNodes (frechet/src/util.cpp:34-58).
//
// Layout: nodes are stored SoA in double (coordinate d of node i at pts[d*stride + i]), so a
// tile of KNN_TILE consecutive nodes is DIM contiguous 2 KB runs, moved to shared memory by
// DIM 1-D bulk async copies (TMA engine) into a double buffer.  One thread owns one
// (query, split) pair: its query sits in registers, every node of the tile is read from shared
// memory (a broadcast when split == 1) and the squared distance is accumulated in the
// reference's operation order with no FMA (sqdist is compared after an IEEE sqrt, so results
// are bit-identical to RealVectorStateSpace::distance).  The running k best are kept per
// thread in shared memory (registers for k == 1) behind a register-resident bound, so the
// common case is 20 DADD/DMUL + one compare per pair.  Partial lists (splits inside a block,
// node chunks across blocks, shards across GPUs) are combined by a warp-shuffle merge under
// the (distance, index) order.
#include <cfloat>
#include <math_constants.h>
#include "knn.cuh"

namespace prgpu {
namespace {

constexpr int MERGE_LISTS_PER_LANE = 4; // up to 128 partial lists per query

__device__ __forceinline__ double bound_of(double d) {
  // smallest convenient U with: sum >= U  =>  sqrt_rn(sum) >= d   (so `sum < U` is a safe filter)
  return __dmul_ru(__dmul_ru(d, d), 1.0000000000000009);
}

__device__ __forceinline__ bool pair_less(double da, uint32_t ia, double db, uint32_t ib) {
  return da < db || (da == db && ia < ib);
}

// Merge P ordered lists of length k (disjoint node sets) into the k smallest under
// (distance, index).  One warp; `get(l, e, d, i)` reads entry e of list l.
template <class Get>
__device__ __forceinline__ void warp_merge(Get get, int P, int k, double *out_d, uint32_t *out_i) {
  const int lane = threadIdx.x & 31;
  int head[MERGE_LISTS_PER_LANE];
  double hd[MERGE_LISTS_PER_LANE];
  uint32_t hi[MERGE_LISTS_PER_LANE];
#pragma unroll
  for (int m = 0; m < MERGE_LISTS_PER_LANE; ++m) {
    int l = lane + 32 * m;
    head[m] = 0;
    hd[m] = CUDART_INF;
    hi[m] = PRGPU_INVALID_INDEX;
    if (l < P)
      get(l, 0, hd[m], hi[m]);
  }
  for (int r = 0; r < k; ++r) {
    double bd = hd[0];
    uint32_t bi = hi[0];
#pragma unroll
    for (int m = 1; m < MERGE_LISTS_PER_LANE; ++m)
      if (pair_less(hd[m], hi[m], bd, bi)) {
        bd = hd[m];
        bi = hi[m];
      }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double od = __shfl_xor_sync(0xffffffffu, bd, off);
      uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (pair_less(od, oi, bd, bi)) {
        bd = od;
        bi = oi;
      }
    }
    if (lane == 0) {
      out_d[r] = bd;
      out_i[r] = bi;
    }
    if (bi != PRGPU_INVALID_INDEX) {
#pragma unroll
      for (int m = 0; m < MERGE_LISTS_PER_LANE; ++m)
        if (hi[m] == bi) {
          int l = lane + 32 * m;
          head[m]++;
          hd[m] = CUDART_INF;
          hi[m] = PRGPU_INVALID_INDEX;
          if (head[m] < k)
            get(l, head[m], hd[m], hi[m]);
        }
    }
  }
}

template <int DIM, bool KONE, bool BCAST>
__global__ void __launch_bounds__(KNN_THREADS)
knn_tile_kernel(const double *__restrict__ pts, uint64_t stride, uint32_t n, uint32_t index_base,
                const double *__restrict__ queries, uint32_t nq, int k, int nsplit_log2,
                const uint32_t *__restrict__ lo, uint32_t tiles_per_chunk, double *__restrict__ out_dist,
                uint32_t *__restrict__ out_idx) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tile = reinterpret_cast<double *>(smem_raw);                       // [2][DIM][TILE]
  uint64_t *bars = reinterpret_cast<uint64_t *>(tile + 2 * DIM * KNN_TILE);  // [2]
  double *ldist = reinterpret_cast<double *>(bars + 2);                      // [k][THREADS]
  uint32_t *lidx = reinterpret_cast<uint32_t *>(ldist + (size_t)k * KNN_THREADS);
  __shared__ uint32_t s_lo_min;

  const int tid = threadIdx.x;
  const uint32_t nsplit = 1u << nsplit_log2;
  const uint32_t vq = blockIdx.x * KNN_THREADS + tid;
  const uint32_t q = vq >> nsplit_log2;
  const uint32_t s = vq & (nsplit - 1);
  const bool active = q < nq;

  double qx[DIM];
  uint32_t my_lo = 0;
#pragma unroll
  for (int d = 0; d < DIM; ++d)
    qx[d] = active ? queries[(size_t)q * DIM + d] : 0.0;
  if (active && lo)
    my_lo = lo[q];

  if (tid == 0) {
    s_lo_min = 0xFFFFFFFFu;
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  for (int e = 0; e < k; ++e) {
    ldist[e * KNN_THREADS + tid] = CUDART_INF;
    lidx[e * KNN_THREADS + tid] = PRGPU_INVALID_INDEX;
  }
  __syncthreads();
  if (active)
    atomicMin(&s_lo_min, my_lo);
  __syncthreads();

  const uint32_t ntiles = (n + KNN_TILE - 1) / KNN_TILE;
  uint32_t t0 = blockIdx.y * tiles_per_chunk;
  const uint32_t t1 = min(t0 + tiles_per_chunk, ntiles);
  if (s_lo_min != 0xFFFFFFFFu)
    t0 = max(t0, s_lo_min / KNN_TILE);

  auto issue = [&](uint32_t t, int buf) {
    mbar_expect_tx(&bars[buf], DIM * KNN_TILE * sizeof(double));
#pragma unroll
    for (int d = 0; d < DIM; ++d)
      bulk_g2s(tile + (buf * DIM + d) * KNN_TILE, pts + (size_t)d * stride + (size_t)t * KNN_TILE,
               KNN_TILE * sizeof(double), &bars[buf]);
  };

  double best_d = CUDART_INF, U = CUDART_INF; // KONE: best; else: current k-th best
  uint32_t best_i = PRGPU_INVALID_INDEX;

  auto offer = [&](double sum, uint32_t gi) {
    if (sum < U) {
      const double dist = __dsqrt_rn(sum);
      if (dist < best_d) {
        if (KONE) {
          best_d = dist;
          best_i = gi;
        } else {
          // insert after every entry <= dist (larger index loses ties: nodes arrive in index order)
          int p = k - 1;
          while (p > 0 && ldist[(p - 1) * KNN_THREADS + tid] > dist) {
            ldist[p * KNN_THREADS + tid] = ldist[(p - 1) * KNN_THREADS + tid];
            lidx[p * KNN_THREADS + tid] = lidx[(p - 1) * KNN_THREADS + tid];
            --p;
          }
          ldist[p * KNN_THREADS + tid] = dist;
          lidx[p * KNN_THREADS + tid] = gi;
          best_d = ldist[(k - 1) * KNN_THREADS + tid];
        }
        U = bound_of(best_d);
      }
    }
  };

  if (tid == 0 && t0 < t1)
    issue(t0, 0);

  for (uint32_t t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    const uint32_t parity = ((t - t0) >> 1) & 1;
    if (tid == 0 && t + 1 < t1)
      issue(t + 1, buf ^ 1);
    mbar_wait(&bars[buf], parity);

    if (active) {
      const double *T = tile + buf * DIM * KNN_TILE;
      const uint32_t base = t * KNN_TILE;
      const uint32_t jmax = min((uint32_t)KNN_TILE, n - base);
      if (BCAST) {
        // every thread visits every node: 4 nodes (two 16-byte broadcast loads per dimension)
        // per step, i.e. 4 independent accumulation chains in flight
        uint32_t j = (my_lo > base) ? ((my_lo - base) & ~3u) : 0u;
        for (; j + 3 < jmax; j += 4) {
          double s0, s1, s2, s3;
          {
            const double2 x = *reinterpret_cast<const double2 *>(T + j);
            const double2 y = *reinterpret_cast<const double2 *>(T + j + 2);
            const double a = __dsub_rn(qx[0], x.x), b = __dsub_rn(qx[0], x.y);
            const double c = __dsub_rn(qx[0], y.x), e = __dsub_rn(qx[0], y.y);
            s0 = __dmul_rn(a, a);
            s1 = __dmul_rn(b, b);
            s2 = __dmul_rn(c, c);
            s3 = __dmul_rn(e, e);
          }
#pragma unroll
          for (int d = 1; d < DIM; ++d) {
            const double2 x = *reinterpret_cast<const double2 *>(T + d * KNN_TILE + j);
            const double2 y = *reinterpret_cast<const double2 *>(T + d * KNN_TILE + j + 2);
            const double a = __dsub_rn(qx[d], x.x), b = __dsub_rn(qx[d], x.y);
            const double c = __dsub_rn(qx[d], y.x), e = __dsub_rn(qx[d], y.y);
            s0 = __dadd_rn(s0, __dmul_rn(a, a));
            s1 = __dadd_rn(s1, __dmul_rn(b, b));
            s2 = __dadd_rn(s2, __dmul_rn(c, c));
            s3 = __dadd_rn(s3, __dmul_rn(e, e));
          }
          const double smin = fmin(fmin(s0, s1), fmin(s2, s3));
          if (smin < U) { // rare: at least one of the four may enter the list
            if (base + j >= my_lo)
              offer(s0, base + j);
            if (base + j + 1 >= my_lo)
              offer(s1, base + j + 1);
            if (base + j + 2 >= my_lo)
              offer(s2, base + j + 2);
            offer(s3, base + j + 3);
          }
        }
        for (; j < jmax; ++j) {
          if (base + j < my_lo)
            continue;
          double a = __dsub_rn(qx[0], T[j]);
          double s0 = __dmul_rn(a, a);
#pragma unroll
          for (int d = 1; d < DIM; ++d) {
            a = __dsub_rn(qx[d], T[d * KNN_TILE + j]);
            s0 = __dadd_rn(s0, __dmul_rn(a, a));
          }
          offer(s0, base + j);
        }
      } else {
        for (uint32_t j = s; j < jmax; j += nsplit) {
          if (base + j < my_lo)
            continue;
          double a = __dsub_rn(qx[0], T[j]);
          double s0 = __dmul_rn(a, a);
#pragma unroll
          for (int d = 1; d < DIM; ++d) {
            a = __dsub_rn(qx[d], T[d * KNN_TILE + j]);
            s0 = __dadd_rn(s0, __dmul_rn(a, a));
          }
          offer(s0, base + j);
        }
      }
    }
    __syncthreads(); // everyone is done with `buf` before it is refilled at t+2
  }

  if (KONE) {
    ldist[tid] = best_d;
    lidx[tid] = best_i;
  }
  __syncthreads();

  // merge the nsplit lists of each query of this block and emit the (query, chunk) partial list
  const int warp = tid >> 5;
  const uint32_t q_per_block = KNN_THREADS >> nsplit_log2;
  for (uint32_t ql = warp; ql < q_per_block; ql += KNN_THREADS / 32) {
    const uint32_t qq = blockIdx.x * q_per_block + ql;
    if (qq >= nq)
      break;
    const uint32_t first = ql << nsplit_log2;
    double *od = out_dist + ((size_t)blockIdx.y * nq + qq) * k;
    uint32_t *oi = out_idx + ((size_t)blockIdx.y * nq + qq) * k;
    auto get = [&](int l, int e, double &d, uint32_t &i) {
      d = ldist[e * KNN_THREADS + first + l];
      i = lidx[e * KNN_THREADS + first + l];
      if (i != PRGPU_INVALID_INDEX)
        i += index_base;
    };
    warp_merge(get, (int)nsplit, k, od, oi);
  }
}

// lists [g*group, min((g+1)*group, nparts)) of every query are merged into list g of the output
__global__ void __launch_bounds__(128)
topk_merge_kernel(const uint32_t *__restrict__ idx_parts, const double *__restrict__ dist_parts, int nparts,
                  int group, uint32_t nq, int k, uint32_t *__restrict__ idx_out,
                  double *__restrict__ dist_out, size_t idx_part_stride, size_t dist_part_stride) {
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq)
    return;
  const int g = blockIdx.y;
  const int first = g * group;
  const int P = min(group, nparts - first);
  auto get = [&](int l, int e, double &d, uint32_t &i) {
    const size_t at = (size_t)q * k + e; // part strides in elements: nq*k for a dense table, the chunk
                                         // size for the [dist | idx] records ranks all-gather
    d = dist_parts[(size_t)(first + l) * dist_part_stride + at];
    i = idx_parts[(size_t)(first + l) * idx_part_stride + at];
  };
  warp_merge(get, P, k, dist_out + ((size_t)g * nq + q) * k, idx_out + ((size_t)g * nq + q) * k);
}

__device__ __forceinline__ float to_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
// row-major (n x dim) -> fp64 SoA at `offset`, plus the fp32 mirror (8 floats per node, |x|^2 in slot 7)
// and the running maximum of |x|^2
__global__ void transpose_append_kernel(const double *__restrict__ aos, uint64_t n, int dim,
                                        double *__restrict__ soa, float *__restrict__ ptsf,
                                        double *__restrict__ xmax2, uint64_t stride, uint64_t offset) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double n2 = 0.0;
  if (i < n) {
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int d = 0; d < dim; ++d) {
      const double v = aos[i * dim + d];
      soa[(uint64_t)d * stride + offset + i] = v;
      n2 += v * v;
      if (d < 7)
        f[d] = (float)v;
    }
    if (ptsf) {
      f[7] = (float)n2;
      float4 *o = reinterpret_cast<float4 *>(ptsf + (offset + i) * 8);
      o[0] = make_float4(f[0], f[1], f[2], f[3]);
      o[1] = make_float4(f[4], f[5], f[6], f[7]);
      // TF32 mirror for the tensor-core scan: coordinates rounded to nearest TF32, 1.0 in slot 7 (the
      // multiplier of the per-query threshold), |x|^2 kept in fp32 in its own (pair-duplicated) array
      float4 *ot = reinterpret_cast<float4 *>(ptsf + stride * 8 + (offset + i) * 8);
      ot[0] = make_float4(to_tf32(f[0]), to_tf32(f[1]), to_tf32(f[2]), to_tf32(f[3]));
      ot[1] = make_float4(to_tf32(f[4]), to_tf32(f[5]), to_tf32(f[6]), 1.0f);
      float *nd = ptsf + stride * 16 + ((offset + i) >> 1) * 4 + ((offset + i) & 1); // (n2a, n2b, n2a, n2b) per node pair:
      nd[0] = f[7];                                                                  // the C fragment of one MMA lane
      nd[2] = f[7];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    n2 = fmax(n2, __shfl_xor_sync(0xffffffffu, n2, off));
  if ((threadIdx.x & 31) == 0 && n2 > 0.0) // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long *>(xmax2), (unsigned long long)__double_as_longlong(n2));
}
__global__ void gather_points_kernel(const double *__restrict__ soa, uint64_t stride, uint64_t first,
                                     uint64_t n, int dim, double *__restrict__ aos) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  for (int d = 0; d < dim; ++d)
    aos[i * dim + d] = soa[(uint64_t)d * stride + first + i];
}

template <int DIM, bool KONE, bool BCAST> size_t tile_smem(int k) {
  return (size_t)2 * DIM * KNN_TILE * sizeof(double) + 2 * sizeof(uint64_t) +
         (size_t)k * KNN_THREADS * (sizeof(double) + sizeof(uint32_t));
}

// resident blocks per SM of the kernel variant that (dim, k, nsplit) selects
template <int DIM> int tile_blocks_per_sm(int k, bool bcast) {
  const size_t smem = tile_smem<DIM, false, false>(k);
  int nb = 0;
#define PRGPU_KNN_OCC(KONE, BCAST)                                                                 \
  do {                                                                                             \
    auto kern = knn_tile_kernel<DIM, KONE, BCAST>;                                                 \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, KNN_THREADS, smem);                   \
  } while (0)
  if (k == 1) {
    if (bcast)
      PRGPU_KNN_OCC(true, true);
    else
      PRGPU_KNN_OCC(true, false);
  } else {
    if (bcast)
      PRGPU_KNN_OCC(false, true);
    else
      PRGPU_KNN_OCC(false, false);
  }
#undef PRGPU_KNN_OCC
  return nb > 0 ? nb : 1;
}

template <int DIM>
int launch_tile(const double *pts, uint64_t stride, uint32_t n, uint32_t index_base, const double *queries,
                uint32_t nq, int k, int nsplit_log2, const uint32_t *lo, uint32_t tiles_per_chunk,
                uint32_t chunks, double *od, uint32_t *oi, cudaStream_t stream) {
  const uint32_t nsplit = 1u << nsplit_log2;
  const uint32_t qblocks = (uint32_t)(((uint64_t)nq * nsplit + KNN_THREADS - 1) / KNN_THREADS);
  const size_t smem = (size_t)2 * DIM * KNN_TILE * sizeof(double) + 2 * sizeof(uint64_t) +
                      (size_t)k * KNN_THREADS * (sizeof(double) + sizeof(uint32_t));
  dim3 grid(qblocks, chunks);
#define PRGPU_KNN_LAUNCH(KONE, BCAST)                                                              \
  do {                                                                                             \
    auto kern = knn_tile_kernel<DIM, KONE, BCAST>;                                                 \
    PRGPU_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, KNN_THREADS, smem, stream>>>(pts, stride, n, index_base, queries, nq, k, nsplit_log2, \
                                              lo, tiles_per_chunk, od, oi); PRGPU_COUNT_LAUNCH(1);                        \
  } while (0)
  if (k == 1) {
    if (nsplit == 1)
      PRGPU_KNN_LAUNCH(true, true);
    else
      PRGPU_KNN_LAUNCH(true, false);
  } else {
    if (nsplit == 1)
      PRGPU_KNN_LAUNCH(false, true);
    else
      PRGPU_KNN_LAUNCH(false, false);
  }
#undef PRGPU_KNN_LAUNCH
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

} // namespace

int scratch_reserve(double *&d, uint32_t *&i, size_t &cap, size_t need) {
  if (cap >= need)
    return PRGPU_OK;
  if (d)
    cudaFree(d);
  if (i)
    cudaFree(i);
  d = nullptr;
  i = nullptr;
  cap = 0;
  PRGPU_CUDA_TRY(cudaMalloc(&d, need * sizeof(double)));
  PRGPU_CUDA_TRY(cudaMalloc(&i, need * sizeof(uint32_t)));
  cap = need;
  return PRGPU_OK;
}

constexpr int MERGE_GROUP = 32 * MERGE_LISTS_PER_LANE;

int topk_merge(const uint32_t *idx_parts, const double *dist_parts, int nparts, uint32_t nq, int k,
               uint32_t *idx_out, double *dist_out, cudaStream_t stream) {
  if (nparts < 1 || nparts > MERGE_GROUP)
    return fail(PRGPU_ELIMIT, "topk_merge: nparts must be in [1,128]");
  if (k < 1 || k > PRGPU_MAX_K)
    return fail(PRGPU_ELIMIT, "topk_merge: k must be in [1,PRGPU_MAX_K]");
  if (nq == 0)
    return PRGPU_OK;
  topk_merge_kernel<<<dim3((nq + 3) / 4, 1), 128, 0, stream>>>(idx_parts, dist_parts, nparts, MERGE_GROUP, nq,
                                                              k, idx_out, dist_out, (size_t)nq * k, (size_t)nq * k); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

int topk_merge_strided(const uint32_t *idx_parts, const double *dist_parts, size_t idx_part_stride,
                       size_t dist_part_stride, int nparts, uint32_t nq, int k, uint32_t *idx_out, double *dist_out,
                       cudaStream_t stream) {
  if (nparts < 1 || nparts > MERGE_GROUP)
    return fail(PRGPU_ELIMIT, "topk_merge: nparts must be in [1,128]");
  if (k < 1 || k > PRGPU_MAX_K)
    return fail(PRGPU_ELIMIT, "topk_merge: k must be in [1,PRGPU_MAX_K]");
  if (nq == 0)
    return PRGPU_OK;
  topk_merge_kernel<<<dim3((nq + 3) / 4, 1), 128, 0, stream>>>(idx_parts, dist_parts, nparts, MERGE_GROUP, nq, k,
                                                              idx_out, dist_out, idx_part_stride, dist_part_stride); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

// any number of partial lists: passes of 128-way merges (parts -> parts2 -> ... -> out)
int merge_passes(KnnScratch &sc, int nparts, uint32_t nq, int k, uint32_t *idx_out, double *dist_out,
                        cudaStream_t stream) {
  const uint32_t *si = sc.part_idx;
  const double *sd = sc.part_dist;
  bool into2 = true;
  while (nparts > MERGE_GROUP) {
    const int groups = (nparts + MERGE_GROUP - 1) / MERGE_GROUP;
    int rc = into2 ? scratch_reserve(sc.part2_dist, sc.part2_idx, sc.part2_capacity, (size_t)groups * nq * k)
                   : PRGPU_OK; // the first buffer is always large enough for a later pass
    if (rc != PRGPU_OK)
      return rc;
    uint32_t *di = into2 ? sc.part2_idx : sc.part_idx;
    double *dd = into2 ? sc.part2_dist : sc.part_dist;
    topk_merge_kernel<<<dim3((nq + 3) / 4, groups), 128, 0, stream>>>(si, sd, nparts, MERGE_GROUP, nq, k, di,
                                                                   dd, (size_t)nq * k, (size_t)nq * k); PRGPU_COUNT_LAUNCH(1);
    PRGPU_CUDA_TRY(cudaGetLastError());
    si = di;
    sd = dd;
    nparts = groups;
    into2 = !into2;
  }
  return topk_merge(si, sd, nparts, nq, k, idx_out, dist_out, stream);
}

int knn_search_exact(int dim, const KnnTreeView &t, const double *queries_dev, uint32_t nq, int k,
                     const uint32_t *lo_dev, uint32_t *idx_dev, double *dist_dev, KnnScratch &scratch, int sm_count,
                     cudaStream_t stream, KernelTimer *timer) {
  const double *pts_soa = t.pts;
  const uint64_t stride = t.stride;
  const uint32_t n = t.n, index_base = t.index_base;
  if (k < 1 || k > PRGPU_MAX_K)
    return fail(PRGPU_ELIMIT, "knn: k must be in [1,PRGPU_MAX_K]");
  if (nq == 0)
    return PRGPU_OK;
  // split a query over several threads only when there are too few queries to fill a block
  int nsplit_log2 = 0;
  while (((uint64_t)nq << nsplit_log2) < KNN_THREADS && nsplit_log2 < 7)
    ++nsplit_log2;
  const uint32_t qblocks = (uint32_t)((((uint64_t)nq << nsplit_log2) + KNN_THREADS - 1) / KNN_THREADS);
  const uint32_t ntiles = (n + KNN_TILE - 1) / KNN_TILE;
  // one wave: as many node chunks as fit in the resident block slots next to the query blocks
  // (a second, partial wave would double the run time of this compute-bound kernel)
  int per_sm = 4;
  switch (dim) {
#define PRGPU_DIM_CASE(D)                                                                          \
  case D:                                                                                          \
    per_sm = tile_blocks_per_sm<D>(k, nsplit_log2 == 0);                                           \
    break;
    PRGPU_DIM_CASE(1)
    PRGPU_DIM_CASE(2)
    PRGPU_DIM_CASE(3)
    PRGPU_DIM_CASE(4)
    PRGPU_DIM_CASE(5)
    PRGPU_DIM_CASE(6)
    PRGPU_DIM_CASE(7)
    PRGPU_DIM_CASE(8)
#undef PRGPU_DIM_CASE
  default:
    return fail(PRGPU_ELIMIT, "knn: dim must be in [1,PRGPU_MAX_DIM]");
  }
  const uint32_t slots = (uint32_t)per_sm * (uint32_t)sm_count;
  uint32_t chunks = slots / qblocks;
  chunks = chunks < 1 ? 1 : chunks;
  if (chunks > 4096)
    chunks = 4096;
  if (chunks > ntiles)
    chunks = ntiles ? ntiles : 1;
  uint32_t tiles_per_chunk = ntiles ? (ntiles + chunks - 1) / chunks : 1;
  chunks = ntiles ? (ntiles + tiles_per_chunk - 1) / tiles_per_chunk : 1;

  double *od = dist_dev;
  uint32_t *oi = idx_dev;
  if (chunks > 1) {
    int rcs = scratch_reserve(scratch.part_dist, scratch.part_idx, scratch.part_capacity,
                              (size_t)chunks * nq * k);
    if (rcs != PRGPU_OK)
      return rcs;
    od = scratch.part_dist;
    oi = scratch.part_idx;
  }
  int rc;
  if (timer)
    timer->begin(stream);
  switch (dim) {
#define PRGPU_DIM_CASE(D)                                                                          \
  case D:                                                                                          \
    rc = launch_tile<D>(pts_soa, stride, n, index_base, queries_dev, nq, k, nsplit_log2, lo_dev,   \
                        tiles_per_chunk, chunks, od, oi, stream);                                  \
    break;
    PRGPU_DIM_CASE(1)
    PRGPU_DIM_CASE(2)
    PRGPU_DIM_CASE(3)
    PRGPU_DIM_CASE(4)
    PRGPU_DIM_CASE(5)
    PRGPU_DIM_CASE(6)
    PRGPU_DIM_CASE(7)
    PRGPU_DIM_CASE(8)
#undef PRGPU_DIM_CASE
  default:
    return fail(PRGPU_ELIMIT, "knn: dim must be in [1,PRGPU_MAX_DIM]");
  }
  if (timer)
    timer->end(stream);
  if (rc != PRGPU_OK)
    return rc;
  if (chunks > 1)
    return merge_passes(scratch, (int)chunks, nq, k, idx_dev, dist_dev, stream);
  return PRGPU_OK;
}

int knn_transpose_append(int dim, const double *aos_dev, uint64_t n_new, double *pts_soa, float *ptsf,
                         double *xmax2, uint64_t stride, uint64_t offset, cudaStream_t stream) {
  if (n_new == 0)
    return PRGPU_OK;
  transpose_append_kernel<<<(unsigned)((n_new + 255) / 256), 256, 0, stream>>>(aos_dev, n_new, dim, pts_soa, ptsf,
                                                                            xmax2, stride, offset); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}
int knn_gather_points(int dim, const double *pts_soa, uint64_t stride, uint64_t first, uint64_t n,
                      double *aos_dev, cudaStream_t stream) {
  if (n == 0)
    return PRGPU_OK;
  gather_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pts_soa, stride, first, n, dim,
                                                                      aos_dev); PRGPU_COUNT_LAUNCH(1);
  PRGPU_CUDA_TRY(cudaGetLastError());
  return PRGPU_OK;
}

} // namespace prgpu

// C ABI (include/prgpu.h) over the kernels: handle bookkeeping, staging of host buffers,
// stream ownership.  No compute happens on the host; every entry point fails loudly when
// no sm_100 device is usable.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "capi_internal.cuh"
#include "rrtc_batch.cuh"

namespace prgpu {
static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

int select_device(int device, int *sm_count) {
  // cudaGetDeviceProperties costs milliseconds: query each device once
  static int cached_count = -1;
  static int cached_major[64], cached_minor[64], cached_sms[64];
  if (cached_count < 0) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
      cudaGetLastError();
      return fail(PRGPU_ENODEV, "no CUDA device available (prgpu has no CPU fallback)");
    }
    if (count > 64)
      count = 64;
    for (int d = 0; d < count; ++d) {
      cudaDeviceProp prop;
      PRGPU_CUDA_TRY(cudaGetDeviceProperties(&prop, d));
      cached_major[d] = prop.major;
      cached_minor[d] = prop.minor;
      cached_sms[d] = prop.multiProcessorCount;
    }
    cached_count = count;
  }
  if (device < 0 || device >= cached_count)
    return fail(PRGPU_EINVAL, "device ordinal out of range");
  if (cached_major[device] != 10)
    return fail(PRGPU_ENODEV, std::string("device is sm_") + std::to_string(cached_major[device] * 10 + cached_minor[device]) +
                                  "; prgpu kernels are built for sm_100a only");
  PRGPU_CUDA_TRY(cudaSetDevice(device));
  if (sm_count)
    *sm_count = cached_sms[device];
  return PRGPU_OK;
}
} // namespace prgpu

using namespace prgpu;

// ---- DevBuf's allocator: the device's stream-ordered pool, kept warm (capi_internal.cuh) ----
namespace {
struct PoolState {
  cudaStream_t stream = nullptr;
  int state = 0; // 0: untried, 1: in use, -1: unavailable (plain cudaMalloc)
};
PoolState g_pool[64];
std::mutex g_pool_mu;
PoolState *pool_for_current_device() {
  static const bool off = getenv("PRGPU_NO_POOL") != nullptr;
  int dev = -1;
  if (off || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return nullptr;
  std::lock_guard<std::mutex> lock(g_pool_mu);
  PoolState &ps = g_pool[dev];
  if (ps.state == 0) {
    ps.state = -1;
    int supported = 0;
    cudaMemPool_t pool = nullptr;
    unsigned long long never = ~0ull;
    if (cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && supported &&
        cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &never) == cudaSuccess &&
        cudaStreamCreateWithFlags(&ps.stream, cudaStreamNonBlocking) == cudaSuccess)
      ps.state = 1;
    else
      cudaGetLastError();
  }
  return ps.state == 1 ? &ps : nullptr;
}
} // namespace
namespace prgpu {
cudaError_t dev_alloc(void **p, size_t bytes) {
  if (PoolState *ps = pool_for_current_device()) {
    cudaError_t e = cudaMallocAsync(p, bytes, ps->stream);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(ps->stream); // usable on every stream from here on
    if (e == cudaSuccess || e == cudaErrorMemoryAllocation)
      return e;
    cudaGetLastError();
  }
  return cudaMalloc(p, bytes);
}
void dev_free(void *p) {
  if (!p)
    return;
  if (PoolState *ps = pool_for_current_device()) {
    cudaDeviceSynchronize(); // cudaFree's semantics: nothing on the device still uses the block
    if (cudaFreeAsync(p, ps->stream) == cudaSuccess)
      return;
    cudaGetLastError();
  }
  cudaFree(p);
}
} // namespace prgpu

extern "C" {

const char *prgpu_last_error(void) { return g_last_error.c_str(); }
int prgpu_version(void) { return 100; }
int prgpu_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

/* ---------------------------------------------- K1 ---------------------------------------------- */

} // extern "C"
int knn_settle(prgpu_knn *h) {
  if (h->ring_used) {
    PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->ring_used = 0;
  }
  return PRGPU_OK;
}
static bool host_window_on() { // PRGPU_NO_HOST_WINDOW: tiny calls take the copy path too (A/B tests)
  return getenv("PRGPU_NO_HOST_WINDOW") == nullptr;
}
extern "C" {

// Grows the tree buffers.  Callers may append on their own stream (prgpu_knn_add_device), so nothing here
// may be ordered only against h->stream: the first allocation zeroes xmax2 and waits for it; a regrow waits
// for the whole device (pending appends into the old buffers, on whatever stream they were issued) before
// copying, and for the copies before the old buffers are freed and the new ones handed out.  The handle is
// left untouched unless everything succeeded.
static int knn_reserve(prgpu_knn *h, uint64_t need) {
  if (need <= h->stride)
    return PRGPU_OK;
  uint64_t cap = std::max<uint64_t>(need, std::max<uint64_t>(2 * h->stride, 4096));
  cap = round_up(cap, KNN_TILE);
  double *np = nullptr, *nx = nullptr;
  float *nf = nullptr;
  cudaError_t e = cudaMalloc(&np, (size_t)h->dim * cap * sizeof(double));
  if (e == cudaSuccess && h->dim <= 7)
    e = cudaMalloc(&nf, (size_t)cap * 18 * sizeof(float)); // fp32 mirror | TF32 mirror | |x|^2 pairs
  if (e == cudaSuccess && !h->xmax2) {
    e = cudaMalloc(&nx, sizeof(double));
    if (e == cudaSuccess)
      e = cudaMemsetAsync(nx, 0, sizeof(double), h->stream);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(h->stream); // an append on a foreign stream may follow immediately
  }
  if (e == cudaSuccess && h->pts && h->n) {
    e = cudaDeviceSynchronize(); // appends still in flight on the caller's stream(s)
    for (int d = 0; d < h->dim && e == cudaSuccess; ++d)
      e = cudaMemcpyAsync(np + (size_t)d * cap, h->pts + (size_t)d * h->stride, h->n * sizeof(double),
                          cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess && nf && h->ptsf) {
      e = cudaMemcpyAsync(nf, h->ptsf, h->n * 8 * sizeof(float), cudaMemcpyDeviceToDevice, h->stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(nf + cap * 8, h->ptsf + h->stride * 8, h->n * 8 * sizeof(float), cudaMemcpyDeviceToDevice,
                            h->stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(nf + cap * 16, h->ptsf + h->stride * 16, (h->n + 1) / 2 * 4 * sizeof(float),
                            cudaMemcpyDeviceToDevice, h->stream);
    }
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(h->stream);
  }
  if (e != cudaSuccess) {
    cudaFree(np);
    cudaFree(nf);
    cudaFree(nx);
    cudaGetLastError();
    return fail(PRGPU_ECUDA, std::string("knn_reserve: ") + cudaGetErrorString(e));
  }
  if (h->pts) {
    if (!h->n)
      cudaDeviceSynchronize(); // nothing to copy, but a search may still be reading the old buffers
    cudaFree(h->pts);
    if (h->ptsf)
      cudaFree(h->ptsf);
  }
  if (nx)
    h->xmax2 = nx;
  h->pts = np;
  h->ptsf = nf;
  h->stride = cap;
  return PRGPU_OK;
}

int prgpu_knn_create(int device, int dim, uint64_t capacity_hint, prgpu_knn_t **out) {
  if (!out)
    return fail(PRGPU_EINVAL, "knn_create: out is NULL");
  *out = nullptr;
  if (dim < 1 || dim > PRGPU_MAX_DIM)
    return fail(PRGPU_ELIMIT, "knn_create: dim must be in [1,PRGPU_MAX_DIM]");
  int sm = 0;
  int rc = select_device(device, &sm);
  if (rc != PRGPU_OK)
    return rc;
  prgpu_knn *h = new prgpu_knn();
  h->device = device;
  h->dim = dim;
  h->sm_count = sm;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    return fail(PRGPU_ECUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  if (capacity_hint) {
    rc = knn_reserve(h, capacity_hint);
    if (rc != PRGPU_OK) {
      prgpu_knn_destroy(h);
      return rc;
    }
  }
  *out = h;
  return PRGPU_OK;
}

int prgpu_knn_destroy(prgpu_knn_t *h) {
  if (!h)
    return PRGPU_OK;
  cudaSetDevice(h->device);
  if (h->stream)
    cudaStreamSynchronize(h->stream);
  if (h->pts)
    cudaFree(h->pts);
  if (h->ptsf)
    cudaFree(h->ptsf);
  if (h->xmax2)
    cudaFree(h->xmax2);
  h->kd.release();
  for (cudaEvent_t ev : h->copy_events)
    if (ev)
      cudaEventDestroy(ev);
  if (h->copy_stream)
    cudaStreamDestroy(h->copy_stream);
  for (void *p : {(void *)h->scratch.cand_dist, (void *)h->scratch.cand_idx, (void *)h->scratch.flags,
                  (void *)h->scratch.fb_q, (void *)h->scratch.fb_dist, (void *)h->scratch.fb_idx,
                  (void *)h->scratch.fb_lo})
    if (p)
      cudaFree(p);
  if (h->scratch.part_dist)
    cudaFree(h->scratch.part_dist);
  if (h->scratch.part_idx)
    cudaFree(h->scratch.part_idx);
  if (h->scratch.part2_dist)
    cudaFree(h->scratch.part2_dist);
  if (h->scratch.part2_idx)
    cudaFree(h->scratch.part2_idx);
  h->stage.release();
  h->d_q.release();
  h->d_lo.release();
  h->d_idx.release();
  h->d_dist.release();
  h->win.release();
  h->timer.destroy();
  if (h->stream)
    cudaStreamDestroy(h->stream);
  delete h;
  return PRGPU_OK;
}

int prgpu_knn_clear(prgpu_knn_t *h) {
  if (!h)
    return fail(PRGPU_EINVAL, "knn_clear: NULL handle");
  h->n = 0;
  h->kd_valid = false;
  h->kd_regret_ms = 0.0;
  if (h->xmax2) {
    PRGPU_CUDA_TRY(cudaSetDevice(h->device));
    PRGPU_CUDA_TRY(cudaMemsetAsync(h->xmax2, 0, sizeof(double), h->stream));
    PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return PRGPU_OK;
}

int prgpu_knn_set_search_mode(prgpu_knn_t *h, int force_exact) {
  if (!h)
    return fail(PRGPU_EINVAL, "knn_set_search_mode: NULL handle");
  h->force_exact = force_exact == 1 ? 1 : 0;
  h->no_index = force_exact == 2 ? 1 : 0;
  h->force_index = force_exact == 3 ? 1 : 0;
  return PRGPU_OK;
}

int prgpu_knn_search_stats(prgpu_knn_t *h, uint64_t *filter_calls, uint64_t *fallback_queries) {
  if (!h)
    return fail(PRGPU_EINVAL, "knn_search_stats: NULL handle");
  if (filter_calls)
    *filter_calls = h->scratch.filter_calls;
  if (fallback_queries)
    *fallback_queries = h->scratch.fallback_queries;
  return PRGPU_OK;
}

int prgpu_knn_size(const prgpu_knn_t *h, uint64_t *n) {
  if (!h || !n)
    return fail(PRGPU_EINVAL, "knn_size: NULL argument");
  *n = h->n;
  return PRGPU_OK;
}

int prgpu_knn_set_index_base(prgpu_knn_t *h, uint32_t index_base) {
  if (!h)
    return fail(PRGPU_EINVAL, "knn_set_index_base: NULL handle");
  h->index_base = index_base;
  return PRGPU_OK;
}

int prgpu_knn_add_device(prgpu_knn_t *h, const double *pts_dev, uint64_t n, void *stream) {
  if (!h || (!pts_dev && n))
    return fail(PRGPU_EINVAL, "knn_add: NULL argument");
  if (h->n + n >= 0xFFFFFFFFull)
    return fail(PRGPU_ELIMIT, "knn_add: more than 2^32-2 points");
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if ((cudaStream_t)stream != h->stream && (rc = knn_settle(h)) != PRGPU_OK)
    return rc;
  rc = knn_reserve(h, h->n + n);
  if (rc != PRGPU_OK)
    return rc;
  rc = knn_transpose_append(h->dim, pts_dev, n, h->pts, h->ptsf, h->xmax2, h->stride, h->n, (cudaStream_t)stream);
  if (rc != PRGPU_OK)
    return rc;
  h->n += n;
  if (n) {
    h->kd_valid = false;
    h->kd_regret_ms = 0.0;
  }
  return PRGPU_OK;
}

int prgpu_knn_add(prgpu_knn_t *h, const double *pts, uint64_t n) {
  if (!h || (!pts && n))
    return fail(PRGPU_EINVAL, "knn_add: NULL argument");
  if (n == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  const size_t bytes = (size_t)n * h->dim * sizeof(double);
  int rc;
  if (n <= 8 && host_window_on() && h->win.ensure(KNN_WIN_BYTES)) {
    // a planner's add(): the points go into the window's ring and the append kernel reads them from there; no
    // copy call and no synchronisation -- the next search is ordered behind it on the same stream.  A slot is
    // reused only after a synchronisation of that stream (every host-buffer search is one).
    if (h->ring_used + n > KNN_WIN_RING_SLOTS && (rc = knn_settle(h)) != PRGPU_OK)
      return rc;
    const size_t off = KNN_WIN_RING + (size_t)h->ring_used * 8 * sizeof(double);
    std::memcpy(h->win.host + off, pts, bytes);
    h->ring_used += (uint32_t)n;
    return prgpu_knn_add_device(h, reinterpret_cast<const double *>(h->win.dev + off), n, h->stream);
  }
  rc = h->stage.reserve(bytes);
  if (rc != PRGPU_OK)
    return rc;
  const uint64_t CH = 1ull << 20; // nodes per pipeline stage
  if (n < 2 * CH) {
    PRGPU_CUDA_TRY(cudaMemcpyAsync(h->stage.p, pts, bytes, cudaMemcpyHostToDevice, h->stream));
    rc = prgpu_knn_add_device(h, h->stage.as<double>(), n, h->stream);
    if (rc != PRGPU_OK)
      return rc;
    PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->ring_used = 0;
    return PRGPU_OK;
  }
  // Bulk loads: the stages are copied on a second stream and transposed (fp64 SoA + the two mirrors) as
  // they land, so only the last stage's transpose is exposed after the upload.
  if (h->n + n >= 0xFFFFFFFFull)
    return fail(PRGPU_ELIMIT, "knn_add: more than 2^32-2 points");
  if ((rc = knn_reserve(h, h->n + n)) != PRGPU_OK)
    return rc;
  if (!h->copy_stream)
    PRGPU_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  const uint64_t stages = (n + CH - 1) / CH;
  if (h->copy_events.size() < stages) {
    const size_t old_n = h->copy_events.size();
    h->copy_events.resize(stages, nullptr);
    for (size_t i = old_n; i < stages; ++i)
      PRGPU_CUDA_TRY(cudaEventCreateWithFlags(&h->copy_events[i], cudaEventDisableTiming));
  }
  PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream)); // earlier users of the staging buffer
  for (uint64_t c = 0; c < stages; ++c) {
    const uint64_t i0 = c * CH, m = std::min<uint64_t>(CH, n - i0);
    PRGPU_CUDA_TRY(cudaMemcpyAsync(h->stage.as<double>() + i0 * h->dim, pts + i0 * h->dim, m * h->dim * sizeof(double),
                                   cudaMemcpyHostToDevice, h->copy_stream));
    PRGPU_CUDA_TRY(cudaEventRecord(h->copy_events[c], h->copy_stream));
  }
  for (uint64_t c = 0; c < stages; ++c) {
    const uint64_t i0 = c * CH, m = std::min<uint64_t>(CH, n - i0);
    PRGPU_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->copy_events[c], 0));
    rc = knn_transpose_append(h->dim, h->stage.as<double>() + i0 * h->dim, m, h->pts, h->ptsf, h->xmax2, h->stride,
                              h->n + i0, h->stream);
    if (rc != PRGPU_OK)
      return rc;
  }
  h->n += n;
  h->kd_valid = false;
  h->kd_regret_ms = 0.0;
  PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PRGPU_OK;
}

int prgpu_knn_get_points(prgpu_knn_t *h, uint64_t first, uint64_t n, double *pts) {
  if (!h || (!pts && n))
    return fail(PRGPU_EINVAL, "knn_get_points: NULL argument");
  if (first + n > h->n)
    return fail(PRGPU_EINVAL, "knn_get_points: range beyond size");
  if (n == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  const size_t bytes = (size_t)n * h->dim * sizeof(double);
  int rc = h->stage.reserve(bytes);
  if (rc != PRGPU_OK)
    return rc;
  rc = knn_gather_points(h->dim, h->pts, h->stride, first, n, h->stage.as<double>(), h->stream);
  if (rc != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(pts, h->stage.p, bytes, cudaMemcpyDeviceToHost, h->stream));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PRGPU_OK;
}

} // extern "C"

// Batched queries over a big, unchanged node set go through the k-d index (kdtree.cu): identical results,
// ~1e-3 of the pairs.  Building costs ~4.4 ns per node (44 ms for 1e7 nodes), so the index is built by the
// "ski rental" rule: every eligible search that had to run without it adds its estimated extra cost to a regret
// account, and the build happens when the account reaches the build cost (never more than twice the cost of the
// best choice in hindsight; a structure that is refilled or appended to between searches never pays for a
// build -- appends and clear() make the index stale and reset the account).  prgpu_knn_build_index() builds it
// at once for callers that know the tree is final (a benchmark's set-up, a roadmap).
// Measured on B200 (scripts/kd_var.py, 1e7 nodes): an indexed search costs 0.20 / 0.37 / 2.08 ms for 512 / 4096 /
// 32768 queries (team search with the fp32 leaf filter: 0.17 ms + 4.4 us * log2(n / 1024) per 1000 queries); the
// TF32 filter scan ~0.1 ms + n*nq / 1.05e10 ms, the exact fp64 kernel (below 128 queries)
// ~ n*nq / 2.3e8 ms.
static const uint32_t KD_MIN_NODES = 1u << 17;
static double brute_ms(uint64_t n, uint32_t nq) {
  const double pairs = (double)n * nq;
  return nq >= 128 ? 0.1 + pairs / 1.05e10 : pairs / 2.3e8;
}
static double indexed_ms(uint64_t n, uint32_t nq) { return 0.17 + nq * 4.4e-6 * std::max(1.0, std::log2((double)n / 1024.0)); }
static bool kd_pays(uint64_t n, uint32_t nq) { // the indexed search is expected to beat the brute-force one by > 10 %
  return nq >= 8 && brute_ms(n, nq) > 1.1 * indexed_ms(n, nq);
}
static int kd_build_now(prgpu_knn *h, cudaStream_t stream) {
  // appends may still be in flight on the caller's stream(s)
  PRGPU_CUDA_TRY(cudaDeviceSynchronize());
  const int rc = kd_build(h->kd, h->dim, h->pts, h->stride, (uint32_t)h->n, stream);
  if (rc == PRGPU_OK)
    h->kd_valid = true;
  return rc;
}
bool knn_index_ready(const prgpu_knn *h, uint32_t nq, int k) {
  return h->kd_valid && !h->force_exact && !h->no_index && k <= 32 &&
         (h->force_index || (h->n >= KD_MIN_NODES && kd_pays(h->n, nq)));
}
int knn_index_search(prgpu_knn *h, const double *queries_dev, uint32_t nq, int k, cudaStream_t stream,
                     const prgpu::KdPeerOut *peers) {
  ++h->kd_calls;
  return kd_search(h->kd, h->index_base, queries_dev, nq, k, nullptr, nullptr, stream, &h->timer, peers);
}
int knn_dispatch(prgpu_knn *h, const double *queries_dev, uint32_t nq, int k, const uint32_t *lo_dev, uint32_t *idx_dev,
                 double *dist_dev, cudaStream_t stream, bool *pending) {
  if (pending)
    *pending = false;
  const bool possible = !h->force_exact && !h->no_index && !lo_dev && k <= 32 && h->n >= 128;
  const bool eligible = possible && (h->force_index || (h->n >= KD_MIN_NODES && kd_pays(h->n, nq)));
  if (eligible) {
    if (!h->kd_valid) {
      h->kd_regret_ms += std::max(0.0, brute_ms(h->n, nq) - indexed_ms(h->n, nq));
      if (h->force_index || h->kd_regret_ms >= 4.4e-6 * (double)h->n) {
        const int rc = kd_build_now(h, stream);
        if (rc != PRGPU_OK && rc != PRGPU_ENOMEM)
          return rc; // out of memory: keep serving from the brute-force path
      }
    }
    if (h->kd_valid) {
      ++h->kd_calls;
      return kd_search(h->kd, h->index_base, queries_dev, nq, k, idx_dev, dist_dev, stream, &h->timer);
    }
  }
  return knn_search_deferred(h->dim, h->view(), queries_dev, nq, k, lo_dev, idx_dev, dist_dev, h->scratch, h->sm_count,
                             stream, &h->timer, h->force_exact, pending);
}

extern "C" {

int prgpu_knn_build_index(prgpu_knn_t *h) {
  if (!h)
    return fail(PRGPU_EINVAL, "knn_build_index: NULL handle");
  if (h->kd_valid)
    return PRGPU_OK;
  if (h->n < 128 || h->dim > 8)
    return fail(PRGPU_EINVAL, "knn_build_index: needs at least 128 nodes");
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  return kd_build_now(h, h->stream);
}
} // extern "C"

extern "C" {

int prgpu_knn_index_stats(prgpu_knn_t *h, uint64_t *indexed_nodes, float *build_ms, uint32_t *builds,
                          uint64_t *index_calls, uint64_t *leaves_visited) {
  if (!h)
    return fail(PRGPU_EINVAL, "knn_index_stats: NULL handle");
  if (indexed_nodes)
    *indexed_nodes = h->kd_valid ? h->kd.n : 0;
  if (build_ms)
    *build_ms = h->kd.build_ms;
  if (builds)
    *builds = h->kd.builds;
  if (index_calls)
    *index_calls = h->kd_calls;
  if (leaves_visited) {
    *leaves_visited = 0;
    if (h->kd.visited) {
      PRGPU_CUDA_TRY(cudaSetDevice(h->device));
      PRGPU_CUDA_TRY(cudaDeviceSynchronize());
      unsigned long long v = 0;
      PRGPU_CUDA_TRY(cudaMemcpy(&v, h->kd.visited, sizeof(v), cudaMemcpyDeviceToHost));
      *leaves_visited = v;
    }
  }
  return PRGPU_OK;
}

int prgpu_knn_nearest_k_device(prgpu_knn_t *h, const double *queries_dev, uint64_t nq, int k,
                               const uint32_t *lo_dev, uint32_t *idx_dev, double *dist_dev, void *stream) {
  if (!h || (nq && (!queries_dev || !idx_dev)))
    return fail(PRGPU_EINVAL, "knn_nearest_k: NULL argument");
  if (k < 1 || k > PRGPU_MAX_K)
    return fail(PRGPU_ELIMIT, "knn_nearest_k: k must be in [1,PRGPU_MAX_K]");
  if (nq > 0x7FFFFFFFull)
    return fail(PRGPU_ELIMIT, "knn_nearest_k: too many queries in one call");
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  if ((cudaStream_t)stream != h->stream) {
    const int rs = knn_settle(h);
    if (rs != PRGPU_OK)
      return rs;
  }
  double *dd = dist_dev;
  if (!dd) {
    int rc = h->d_dist.reserve((size_t)nq * k * sizeof(double));
    if (rc != PRGPU_OK)
      return rc;
    dd = h->d_dist.as<double>();
  }
  return knn_dispatch(h, queries_dev, (uint32_t)nq, k, lo_dev, idx_dev, dd, (cudaStream_t)stream, nullptr);
}

int prgpu_knn_nearest_k(prgpu_knn_t *h, const double *queries, uint64_t nq, int k, const uint32_t *lo,
                        uint32_t *idx, double *dist) {
  if (!h || (nq && (!queries || !idx)))
    return fail(PRGPU_EINVAL, "knn_nearest_k: NULL argument");
  if (k < 1 || k > PRGPU_MAX_K)
    return fail(PRGPU_ELIMIT, "knn_nearest_k: k must be in [1,PRGPU_MAX_K]");
  if (nq == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if (nq <= KNN_WIN_MAX_Q && h->n <= KNN_WIN_MAX_NODES && host_window_on() && h->win.ensure(KNN_WIN_BYTES)) {
    // a planner's nearest(): query and result travel through the mapped window (see HostWindow)
    std::memcpy(h->win.host + KNN_WIN_Q, queries, (size_t)nq * h->dim * sizeof(double));
    if (lo)
      std::memcpy(h->win.host + KNN_WIN_LO, lo, (size_t)nq * sizeof(uint32_t));
    rc = knn_dispatch(h, reinterpret_cast<const double *>(h->win.dev + KNN_WIN_Q), (uint32_t)nq, k,
                      lo ? reinterpret_cast<const uint32_t *>(h->win.dev + KNN_WIN_LO) : nullptr,
                      reinterpret_cast<uint32_t *>(h->win.dev + KNN_WIN_IDX),
                      reinterpret_cast<double *>(h->win.dev + KNN_WIN_DIST), h->stream, nullptr);
    if (rc != PRGPU_OK)
      return rc;
    PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->ring_used = 0;
    std::memcpy(idx, h->win.host + KNN_WIN_IDX, (size_t)nq * k * sizeof(uint32_t));
    if (dist)
      std::memcpy(dist, h->win.host + KNN_WIN_DIST, (size_t)nq * k * sizeof(double));
    return PRGPU_OK;
  }
  if ((rc = h->d_q.reserve((size_t)nq * h->dim * sizeof(double))) != PRGPU_OK)
    return rc;
  if ((rc = h->d_idx.reserve((size_t)nq * k * sizeof(uint32_t))) != PRGPU_OK)
    return rc;
  if ((rc = h->d_dist.reserve((size_t)nq * k * sizeof(double))) != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(h->d_q.p, queries, (size_t)nq * h->dim * sizeof(double),
                                 cudaMemcpyHostToDevice, h->stream));
  if (!lo && host_window_on() && knn_index_ready(h, (uint32_t)nq, k)) {
    // An indexed search writes a query's row once, when its last worker is done.  If the caller's result buffers
    // are pinned host memory the kernel writes the rows there itself (zero copy): the transfer rides under the
    // search instead of following it as two copies.
    cudaPointerAttributes ai, ad;
    std::memset(&ai, 0, sizeof(ai));
    std::memset(&ad, 0, sizeof(ad));
    const bool pin_i = cudaPointerGetAttributes(&ai, idx) == cudaSuccess && ai.type == cudaMemoryTypeHost && ai.devicePointer;
    const bool pin_d = !dist || (cudaPointerGetAttributes(&ad, dist) == cudaSuccess && ad.type == cudaMemoryTypeHost && ad.devicePointer);
    cudaGetLastError(); // (older drivers report unregistered memory as an error)
    if (pin_i && pin_d) {
      rc = knn_dispatch(h, h->d_q.as<double>(), (uint32_t)nq, k, nullptr, static_cast<uint32_t *>(ai.devicePointer),
                        dist ? static_cast<double *>(ad.devicePointer) : h->d_dist.as<double>(), h->stream, nullptr);
      if (rc != PRGPU_OK)
        return rc;
      PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
      h->ring_used = 0;
      return PRGPU_OK;
    }
  }
  const uint32_t *lo_dev = nullptr;
  if (lo) {
    if ((rc = h->d_lo.reserve((size_t)nq * sizeof(uint32_t))) != PRGPU_OK)
      return rc;
    PRGPU_CUDA_TRY(
        cudaMemcpyAsync(h->d_lo.p, lo, (size_t)nq * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    lo_dev = h->d_lo.as<uint32_t>();
  }
  rc = knn_dispatch(h, h->d_q.as<double>(), (uint32_t)nq, k, lo_dev, h->d_idx.as<uint32_t>(), h->d_dist.as<double>(),
                    h->stream, nullptr);
  if (rc != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(idx, h->d_idx.p, (size_t)nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                 h->stream));
  if (dist)
    PRGPU_CUDA_TRY(cudaMemcpyAsync(dist, h->d_dist.p, (size_t)nq * k * sizeof(double),
                                   cudaMemcpyDeviceToHost, h->stream));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->ring_used = 0;
  return PRGPU_OK;
}

int prgpu_knn_set_profiling(prgpu_knn_t *h, int on) {
  if (!h)
    return fail(PRGPU_EINVAL, "knn_set_profiling: NULL handle");
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  return h->timer.enable(on != 0);
}
int prgpu_knn_last_kernel_ms(prgpu_knn_t *h, float *ms) {
  if (!h || !ms)
    return fail(PRGPU_EINVAL, "knn_last_kernel_ms: NULL argument");
  return h->timer.read(ms);
}

int prgpu_topk_merge_device(int device, const uint32_t *idx_parts_dev, const double *dist_parts_dev,
                            int nparts, uint64_t nq, int k, uint32_t *idx_dev, double *dist_dev,
                            void *stream) {
  if (!idx_parts_dev || !dist_parts_dev || !idx_dev || !dist_dev)
    return fail(PRGPU_EINVAL, "topk_merge: NULL argument");
  int rc = select_device(device, nullptr);
  if (rc != PRGPU_OK)
    return rc;
  return topk_merge(idx_parts_dev, dist_parts_dev, nparts, (uint32_t)nq, k, idx_dev, dist_dev,
                    (cudaStream_t)stream);
}

/* ---------------------------------------------- K2 ---------------------------------------------- */

int prgpu_world_create(int device, int dof, const double *dh, int ns, const int32_t *sphere_link,
                       const double *sphere_xyzr, int nos, const double *obs_spheres, int nob,
                       const double *obs_boxes, prgpu_world_t **out) {
  if (!out)
    return fail(PRGPU_EINVAL, "world_create: out is NULL");
  *out = nullptr;
  if (dof < 1 || dof > WORLD_MAX_DOF)
    return fail(PRGPU_ELIMIT, "world_create: dof must be in [1,PRGPU_MAX_DIM]");
  if (ns < 0 || ns > WORLD_MAX_S)
    return fail(PRGPU_ELIMIT, "world_create: at most PRGPU_MAX_LINK_SPHERES link spheres");
  if (nos < 0 || nob < 0 || nos + nob > PRGPU_MAX_OBSTACLES)
    return fail(PRGPU_ELIMIT, "world_create: at most PRGPU_MAX_OBSTACLES obstacles");
  if (!dh || (ns && (!sphere_link || !sphere_xyzr)) || (nos && !obs_spheres) || (nob && !obs_boxes))
    return fail(PRGPU_EINVAL, "world_create: NULL array");
  for (int s = 0; s < ns; ++s)
    if (sphere_link[s] < 0 || sphere_link[s] > dof)
      return fail(PRGPU_EINVAL, "world_create: sphere_link out of [0,dof]");
  int sm = 0;
  int rc = select_device(device, &sm);
  if (rc != PRGPU_OK)
    return rc;

  prgpu_world *w = new prgpu_world();
  w->device = device;
  w->sm_count = sm;
  DeviceWorld &H = w->host;
  std::memset(&H, 0, sizeof(H));
  H.dof = dof;
  H.S = ns;
  H.n_os = nos;
  H.n_ob = nob;
  double reach = 0.0, rmax = 0.0, maxabs = 0.0;
  for (int i = 0; i < dof; ++i) {
    for (int c = 0; c < 5; ++c)
      H.dh[i][c] = dh[5 * i + c];
    reach += std::fabs(dh[5 * i + 0]) + std::fabs(dh[5 * i + 1]);
  }
  double maxlocal = 0.0;
  for (int s = 0; s < ns; ++s) {
    H.sphere_link[s] = sphere_link[s];
    for (int c = 0; c < 4; ++c)
      H.sphere_local[s][c] = sphere_xyzr[4 * s + c];
    maxlocal = std::max(maxlocal, std::fabs(sphere_xyzr[4 * s]) + std::fabs(sphere_xyzr[4 * s + 1]) +
                                      std::fabs(sphere_xyzr[4 * s + 2]));
    rmax = std::max(rmax, sphere_xyzr[4 * s + 3]);
  }
  for (int o = 0; o < nos; ++o)
    for (int c = 0; c < 3; ++c)
      maxabs = std::max(maxabs, std::fabs(obs_spheres[4 * o + c]));
  for (int o = 0; o < nob; ++o)
    for (int c = 0; c < 3; ++c)
      maxabs = std::max(maxabs, std::fabs(obs_boxes[6 * o + c]) + std::fabs(obs_boxes[6 * o + 3 + c]));
  // fp32 filter margin.  (a) rounding of obstacle and sphere centres to fp32 plus the fp32 distance
  // arithmetic: <= 4 u scale, taken with 16x headroom; (b) the fp32 forward-kinematics chain of the
  // fast path: per joint ~4 u relative error on every frame entry and an angle error of ~4 u pi from
  // sincosf and the rounded joint value, i.e. <= 8 u (dof+1) (reach + |local|) on a centre, taken with
  // 8x headroom.  u = 2^-24.
  const double u32 = 5.9604644775390625e-08;
  const double scale = std::max(1.0, maxabs + reach + maxlocal);
  const double delta = 16.0 * u32 * scale * 4.0 + 64.0 * u32 * (dof + 1) * (reach + maxlocal);
  H.delta = (float)delta;
  for (int s = 0; s < WORLD_MAX_S; ++s) {
    if (s < ns) {
      H.sphere_r[s] = (float)(H.sphere_local[s][3] + delta);
    } else { // padding sphere: attached to the base, far away, never hits
      H.sphere_link[s] = 0;
      H.sphere_local[s][0] = 1.0e18;
      H.sphere_local[s][1] = 0.0;
      H.sphere_local[s][2] = 0.0;
      H.sphere_local[s][3] = 0.0;
      H.sphere_r[s] = 0.0f;
    }
  }
  const double bt = rmax + delta;
  H.box_thr2 = (float)(bt * bt * (1.0 + 1e-6));
  { // the fp32 arm table of the filter kernels: spheres grouped by the frame they ride on (counting sort, stable)
    ArmTableF32 &arm = H.arm_f32;
    std::memset(&arm, 0, sizeof(arm));
    for (int i = 0; i < WORLD_MAX_DOF; ++i)
      for (int c = 0; c < 5; ++c)
        arm.dh[i][c] = (float)H.dh[i][c];
    int n = 0;
    for (int i = 0; i <= H.dof; ++i) {
      arm.link_first[i] = (uint8_t)n;
      for (int sp = 0; sp < H.S; ++sp)
        if (H.sphere_link[sp] == i) {
          arm.sphere[n] = make_float4((float)H.sphere_local[sp][0], (float)H.sphere_local[sp][1],
                                      (float)H.sphere_local[sp][2], H.sphere_r[sp]);
          arm.sphere_id[n] = (uint8_t)sp;
          ++n;
        }
    }
    arm.link_first[H.dof + 1] = (uint8_t)n;
  }

  // ---- broad phase: uniform grid over the inflated obstacle boxes --------------------------------
  std::vector<uint32_t> cell_start;
  std::vector<uint16_t> cell_item;
  H.grid_n[0] = H.grid_n[1] = H.grid_n[2] = 0;
  H.grid_items = 0;
  if (nos + nob > 0 && nos < 0x8000 && nob < 0x8000 && !getenv("PRGPU_NO_GRID")) {
    const double infl = rmax + delta + 1e-4; // + slack for the fp32 cell index of a centre
    double lo3[3] = {1e300, 1e300, 1e300}, hi3[3] = {-1e300, -1e300, -1e300};
    auto box_of = [&](int id, bool isbox, double *bl, double *bh) {
      for (int c = 0; c < 3; ++c) {
        const double ctr = isbox ? obs_boxes[6 * id + c] : obs_spheres[4 * id + c];
        const double ext = (isbox ? obs_boxes[6 * id + 3 + c] : obs_spheres[4 * id + 3]) + infl;
        bl[c] = ctr - ext;
        bh[c] = ctr + ext;
      }
    };
    for (int pass = 0; pass < 2; ++pass)
      for (int id = 0; id < (pass ? nob : nos); ++id) {
        double bl[3], bh[3];
        box_of(id, pass == 1, bl, bh);
        for (int c = 0; c < 3; ++c) {
          lo3[c] = std::min(lo3[c], bl[c]);
          hi3[c] = std::max(hi3[c], bh[c]);
        }
      }
    double cell = std::max(0.15, 2.5 * infl);
    for (int attempt = 0; attempt < 12; ++attempt, cell *= 1.3) {
      int nn[3];
      uint64_t ncell = 1;
      for (int c = 0; c < 3; ++c) {
        nn[c] = std::max(1, (int)std::ceil((hi3[c] - lo3[c]) / cell));
        ncell *= (uint64_t)nn[c];
      }
      if (ncell > 4096)
        continue;
      std::vector<std::vector<uint16_t>> lists(ncell);
      uint64_t items = 0;
      for (int pass = 0; pass < 2; ++pass)
        for (int id = 0; id < (pass ? nob : nos); ++id) {
          double bl[3], bh[3];
          box_of(id, pass == 1, bl, bh);
          int c0[3], c1[3];
          for (int c = 0; c < 3; ++c) {
            c0[c] = std::min(nn[c] - 1, std::max(0, (int)std::floor((bl[c] - lo3[c]) / cell)));
            c1[c] = std::min(nn[c] - 1, std::max(0, (int)std::floor((bh[c] - lo3[c]) / cell)));
          }
          for (int z = c0[2]; z <= c1[2]; ++z)
            for (int y = c0[1]; y <= c1[1]; ++y)
              for (int x = c0[0]; x <= c1[0]; ++x) {
                lists[((size_t)z * nn[1] + y) * nn[0] + x].push_back((uint16_t)(id | (pass ? 0x8000 : 0)));
                ++items;
              }
        }
      if (items > 16384)
        continue;
      cell_start.assign(ncell + 1, 0);
      cell_item.clear();
      for (uint64_t i = 0; i < ncell; ++i) {
        cell_start[i] = (uint32_t)cell_item.size();
        cell_item.insert(cell_item.end(), lists[i].begin(), lists[i].end());
      }
      cell_start[ncell] = (uint32_t)cell_item.size();
      for (int c = 0; c < 3; ++c) {
        H.grid_n[c] = nn[c];
        H.grid_org[c] = (float)lo3[c];
      }
      H.grid_inv = (float)(1.0 / cell);
      H.grid_items = (uint32_t)cell_item.size();
      break;
    }
  }
  const bool use_grid = H.grid_n[0] > 0;

  std::vector<float4> osf(nos), obf(2 * (size_t)nob);
  for (int o = 0; o < nos; ++o) {
    const double t = obs_spheres[4 * o + 3] + rmax + delta;
    // .w: the grid path wants the obstacle radius, the brute-force filter its level-1 threshold
    osf[o] = make_float4((float)obs_spheres[4 * o], (float)obs_spheres[4 * o + 1], (float)obs_spheres[4 * o + 2],
                         use_grid ? (float)(obs_spheres[4 * o + 3] * (1.0 + 1e-6)) : (float)(t * t * (1.0 + 1e-6)));
  }
  for (int o = 0; o < nob; ++o) {
    obf[2 * o] = make_float4((float)obs_boxes[6 * o], (float)obs_boxes[6 * o + 1],
                             (float)obs_boxes[6 * o + 2], 0.f);
    obf[2 * o + 1] = make_float4((float)obs_boxes[6 * o + 3], (float)obs_boxes[6 * o + 4],
                                 (float)obs_boxes[6 * o + 5], 0.f);
  }
  auto upload = [&](DevBuf &b, const void *src, size_t bytes) -> int {
    int r = b.reserve(std::max<size_t>(bytes, 16));
    if (r != PRGPU_OK)
      return r;
    if (bytes)
      PRGPU_CUDA_TRY(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    return PRGPU_OK;
  };
  cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete w;
    return fail(PRGPU_ECUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  if ((rc = upload(w->os_f, osf.data(), osf.size() * sizeof(float4))) != PRGPU_OK ||
      (rc = upload(w->ob_f, obf.data(), obf.size() * sizeof(float4))) != PRGPU_OK ||
      (rc = upload(w->os_d, obs_spheres, (size_t)nos * 4 * sizeof(double))) != PRGPU_OK ||
      (rc = upload(w->ob_d, obs_boxes, (size_t)nob * 6 * sizeof(double))) != PRGPU_OK ||
      (rc = upload(w->cell_start, cell_start.data(), cell_start.size() * sizeof(uint32_t))) != PRGPU_OK ||
      (rc = upload(w->cell_item, cell_item.data(), cell_item.size() * sizeof(uint16_t))) != PRGPU_OK) {
    prgpu_world_destroy(w);
    return rc;
  }
  H.os_f = w->os_f.as<float4>();
  H.ob_f = w->ob_f.as<float4>();
  H.os_d = w->os_d.as<double>();
  H.ob_d = w->ob_d.as<double>();
  H.cell_start = w->cell_start.as<uint32_t>();
  H.cell_item = w->cell_item.as<uint16_t>();

  // ---- clearance field + joint lever arms (conservative advancement along edges) ---------------
  H.df = nullptr;
  H.df_n[0] = H.df_n[1] = H.df_n[2] = 0;
  H.df_inv = H.df_out = 0.f;
  for (int i = 0; i < WORLD_MAX_DOF; ++i) {
    // joint i turns about the z axis of frame i; a sphere on frame l > i is at most the chain length
    // from that frame's origin plus its local offset away from it
    double best = 0.0;
    for (int sidx = 0; sidx < ns; ++sidx) {
      const int l = sphere_link[sidx];
      if (i >= dof || l < i + 1)
        continue;
      double len = std::sqrt(sphere_xyzr[4 * sidx] * sphere_xyzr[4 * sidx] + sphere_xyzr[4 * sidx + 1] * sphere_xyzr[4 * sidx + 1] +
                             sphere_xyzr[4 * sidx + 2] * sphere_xyzr[4 * sidx + 2]);
      for (int m2 = i; m2 < l; ++m2)
        len += std::sqrt(dh[5 * m2] * dh[5 * m2] + dh[5 * m2 + 1] * dh[5 * m2 + 1]);
      best = std::max(best, len);
    }
    H.rho[i] = (float)(best * (1.0 + 1e-6));
  }
  if (use_grid && !getenv("PRGPU_NO_DFIELD")) {
    const double margin = 0.25; // the field's box: the obstacles' bounding box grown by this much
    double lo3[3] = {1e300, 1e300, 1e300}, hi3[3] = {-1e300, -1e300, -1e300};
    for (int o = 0; o < nos; ++o)
      for (int c = 0; c < 3; ++c) {
        lo3[c] = std::min(lo3[c], obs_spheres[4 * o + c] - obs_spheres[4 * o + 3]);
        hi3[c] = std::max(hi3[c], obs_spheres[4 * o + c] + obs_spheres[4 * o + 3]);
      }
    for (int o = 0; o < nob; ++o)
      for (int c = 0; c < 3; ++c) {
        lo3[c] = std::min(lo3[c], obs_boxes[6 * o + c] - obs_boxes[6 * o + 3 + c]);
        hi3[c] = std::max(hi3[c], obs_boxes[6 * o + c] + obs_boxes[6 * o + 3 + c]);
      }
    double vol = 1.0;
    for (int c = 0; c < 3; ++c) {
      lo3[c] -= margin;
      hi3[c] += margin;
      vol *= hi3[c] - lo3[c];
    }
    // about eight million cells (32 MB, L2 resident; 1e6 cells: 2.67 ms per 1e6 edges, 8e6: 2.39 ms), never
    // finer than 1 cm
    const float inv_f = (float)(1.0 / std::max(0.01, std::cbrt(vol / 8.0e6)));
    const double cell = 1.0 / (double)inv_f; // the lookup multiplies by inv_f: build with the same number
    int n3[3];
    double org[3];
    uint64_t ncell = 1;
    for (int c = 0; c < 3; ++c) {
      org[c] = (double)(float)lo3[c];
      n3[c] = std::max(1, (int)std::ceil((hi3[c] - org[c]) / cell));
      ncell *= (uint64_t)n3[c];
    }
    // slack: fp32 position error of a centre (delta), its effect on the cell index, rounding of the lookup
    const double slack = delta + 2e-4;
    // a centre outside the box is >= margin - (org rounding) away from every obstacle
    const double out = margin - 1e-3 - slack;
    if (ncell <= (1ull << 24) && out > 0.0) {
      rc = w->df.reserve(ncell * sizeof(float));
      if (rc == PRGPU_OK)
        rc = dfield_build_launch(H.os_d, nos, H.ob_d, nob, n3, org, cell, slack, 1.0e3, w->df.as<uint32_t>(), w->stream);
      if (rc == PRGPU_OK && cudaStreamSynchronize(w->stream) != cudaSuccess)
        rc = fail(PRGPU_ECUDA, "clearance field build failed");
      if (rc != PRGPU_OK) {
        prgpu_world_destroy(w);
        return rc;
      }
      H.df = w->df.as<uint32_t>();
      for (int c = 0; c < 3; ++c) {
        H.df_n[c] = n3[c];
        H.df_org[c] = (float)org[c];
      }
      H.df_inv = inv_f;
      H.df_out = (float)out;
    }
  }
  e = cudaMalloc(&w->dev, sizeof(DeviceWorld));
  if (e == cudaSuccess)
    e = cudaMemcpy(w->dev, &H, sizeof(DeviceWorld), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    prgpu_world_destroy(w);
    return fail(PRGPU_ECUDA, std::string("world upload: ") + cudaGetErrorString(e));
  }
  *out = w;
  return PRGPU_OK;
}

int prgpu_world_destroy(prgpu_world_t *w) {
  if (!w)
    return PRGPU_OK;
  cudaSetDevice(w->device);
  if (w->stream)
    cudaStreamSynchronize(w->stream);
  w->os_f.release();
  w->ob_f.release();
  w->os_d.release();
  w->ob_d.release();
  w->cell_start.release();
  w->cell_item.release();
  w->df.release();
  w->d_a.release();
  w->d_b.release();
  w->d_valid.release();
  w->d_pose.release();
  w->win.release();
  for (void *p : {(void *)w->cm_scratch.owed, (void *)w->cm_scratch.off, w->cm_scratch.tmp, (void *)w->cm_scratch.item_edge,
                  (void *)w->cm_scratch.gaps, (void *)w->cm_scratch.queue, (void *)w->cm_scratch.queue_n,
                  (void *)w->cm_scratch.pair_ref, w->cm_scratch.pair_centre, (void *)w->cm_scratch.pair_n})
    if (p)
      cudaFree(p);
  w->timer.destroy();
  for (cudaEvent_t ev : w->copy_events)
    if (ev)
      cudaEventDestroy(ev);
  if (w->copy_stream)
    cudaStreamDestroy(w->copy_stream);
  if (w->dev)
    cudaFree(w->dev);
  if (w->stream)
    cudaStreamDestroy(w->stream);
  delete w;
  return PRGPU_OK;
}

int prgpu_world_set_profiling(prgpu_world_t *w, int on) {
  if (!w)
    return fail(PRGPU_EINVAL, "world_set_profiling: NULL handle");
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  return w->timer.enable(on != 0);
}
int prgpu_world_last_kernel_ms(prgpu_world_t *w, float *ms) {
  if (!w || !ms)
    return fail(PRGPU_EINVAL, "world_last_kernel_ms: NULL argument");
  return w->timer.read(ms);
}

int prgpu_world_dof(const prgpu_world_t *w) { return w ? w->host.dof : PRGPU_EINVAL; }

int prgpu_check_motion_batch_device(prgpu_world_t *w, const double *from_dev, const double *to_dev,
                                    uint64_t n, double resolution, int mode, int check_first,
                                    uint8_t *valid_dev, void *stream) {
  if (!w || (n && (!from_dev || !to_dev || !valid_dev)))
    return fail(PRGPU_EINVAL, "check_motion_batch: NULL argument");
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  return check_motion_launch(w->host, w->dev, from_dev, to_dev, n, resolution, mode, check_first, valid_dev,
                             w->sm_count, (cudaStream_t)stream, &w->timer, &w->cm_scratch);
}

int prgpu_check_motion_batch(prgpu_world_t *w, const double *from, const double *to, uint64_t n,
                             double resolution, int mode, int check_first, uint8_t *valid) {
  if (!w || (n && (!from || !to || !valid)))
    return fail(PRGPU_EINVAL, "check_motion_batch: NULL argument");
  if (n == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  const size_t bytes = (size_t)n * w->host.dof * sizeof(double);
  int rc;
  if ((rc = w->d_a.reserve(bytes)) != PRGPU_OK || (rc = w->d_b.reserve(bytes)) != PRGPU_OK ||
      (rc = w->d_valid.reserve(n)) != PRGPU_OK)
    return rc;
  const uint64_t CH = 1ull << 18; // edges per pipeline stage
  if (n <= 32 && host_window_on() && w->win.ensure(8192)) {
    // a planner's checkMotion(): the edge and the verdict travel through the mapped window (see HostWindow)
    std::memcpy(w->win.host, from, bytes);
    std::memcpy(w->win.host + 2048, to, bytes);
    rc = check_motion_launch(w->host, w->dev, reinterpret_cast<const double *>(w->win.dev),
                             reinterpret_cast<const double *>(w->win.dev + 2048), n, resolution, mode, check_first,
                             w->win.dev + 4096, w->sm_count, w->stream, &w->timer, &w->cm_scratch);
    if (rc != PRGPU_OK)
      return rc;
    PRGPU_CUDA_TRY(cudaStreamSynchronize(w->stream));
    std::memcpy(valid, w->win.host + 4096, n);
    return PRGPU_OK;
  }
  if (n < 2 * CH) {
    PRGPU_CUDA_TRY(cudaMemcpyAsync(w->d_a.p, from, bytes, cudaMemcpyHostToDevice, w->stream));
    PRGPU_CUDA_TRY(cudaMemcpyAsync(w->d_b.p, to, bytes, cudaMemcpyHostToDevice, w->stream));
    rc = check_motion_launch(w->host, w->dev, w->d_a.as<double>(), w->d_b.as<double>(), n, resolution, mode,
                             check_first, w->d_valid.as<uint8_t>(), w->sm_count, w->stream, &w->timer, &w->cm_scratch);
    if (rc != PRGPU_OK)
      return rc;
    PRGPU_CUDA_TRY(cudaMemcpyAsync(valid, w->d_valid.p, n, cudaMemcpyDeviceToHost, w->stream));
    PRGPU_CUDA_TRY(cudaStreamSynchronize(w->stream));
    return PRGPU_OK;
  }
  // Large batches: the upload of stage c+1 (copy stream) runs under the kernels of stage c.  Each stage's
  // launch reads one counter back on the compute stream; the copies are already queued by then.
  if (!w->copy_stream)
    PRGPU_CUDA_TRY(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
  const uint64_t stages = (n + CH - 1) / CH;
  if (w->copy_events.size() < stages) {
    const size_t old_n = w->copy_events.size();
    w->copy_events.resize(stages, nullptr);
    for (size_t i = old_n; i < stages; ++i)
      PRGPU_CUDA_TRY(cudaEventCreateWithFlags(&w->copy_events[i], cudaEventDisableTiming));
  }
  const int dof = w->host.dof;
  auto upload_stage = [&](uint64_t c) -> int {
    const uint64_t e0 = c * CH, m = std::min<uint64_t>(CH, n - e0);
    PRGPU_CUDA_TRY(cudaMemcpyAsync(w->d_a.as<double>() + e0 * dof, from + e0 * dof, m * dof * sizeof(double),
                                   cudaMemcpyHostToDevice, w->copy_stream));
    PRGPU_CUDA_TRY(cudaMemcpyAsync(w->d_b.as<double>() + e0 * dof, to + e0 * dof, m * dof * sizeof(double),
                                   cudaMemcpyHostToDevice, w->copy_stream));
    PRGPU_CUDA_TRY(cudaEventRecord(w->copy_events[c], w->copy_stream));
    return PRGPU_OK;
  };
  // the previous call's kernels may still read d_a / d_b on the compute stream only if the caller used the
  // device entry points concurrently; order the first copy after them anyway
  PRGPU_CUDA_TRY(cudaStreamSynchronize(w->stream));
  if ((rc = upload_stage(0)) != PRGPU_OK)
    return rc;
  for (uint64_t c = 0; c < stages; ++c) {
    if (c + 1 < stages && (rc = upload_stage(c + 1)) != PRGPU_OK)
      return rc;
    const uint64_t e0 = c * CH, m = std::min<uint64_t>(CH, n - e0);
    PRGPU_CUDA_TRY(cudaStreamWaitEvent(w->stream, w->copy_events[c], 0));
    rc = check_motion_launch(w->host, w->dev, w->d_a.as<double>() + e0 * dof, w->d_b.as<double>() + e0 * dof, m,
                             resolution, mode, check_first, w->d_valid.as<uint8_t>() + e0, w->sm_count, w->stream,
                             c == 0 ? &w->timer : nullptr, &w->cm_scratch);
    if (rc != PRGPU_OK)
      return rc;
    PRGPU_CUDA_TRY(cudaMemcpyAsync(valid + e0, w->d_valid.as<uint8_t>() + e0, m, cudaMemcpyDeviceToHost, w->stream));
  }
  PRGPU_CUDA_TRY(cudaStreamSynchronize(w->stream));
  return PRGPU_OK;
}

int prgpu_is_valid_batch_device(prgpu_world_t *w, const double *q_dev, uint64_t n, uint8_t *valid_dev,
                                void *stream) {
  if (!w || (n && (!q_dev || !valid_dev)))
    return fail(PRGPU_EINVAL, "is_valid_batch: NULL argument");
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  return is_valid_launch(w->host, w->dev, q_dev, n, valid_dev, (cudaStream_t)stream);
}

int prgpu_is_valid_batch(prgpu_world_t *w, const double *q, uint64_t n, uint8_t *valid) {
  if (!w || (n && (!q || !valid)))
    return fail(PRGPU_EINVAL, "is_valid_batch: NULL argument");
  if (n == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  const size_t bytes = (size_t)n * w->host.dof * sizeof(double);
  int rc;
  if ((rc = w->d_a.reserve(bytes)) != PRGPU_OK || (rc = w->d_valid.reserve(n)) != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(w->d_a.p, q, bytes, cudaMemcpyHostToDevice, w->stream));
  rc = is_valid_launch(w->host, w->dev, w->d_a.as<double>(), n, w->d_valid.as<uint8_t>(), w->stream);
  if (rc != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(valid, w->d_valid.p, n, cudaMemcpyDeviceToHost, w->stream));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(w->stream));
  return PRGPU_OK;
}

int prgpu_fk_batch_device(prgpu_world_t *w, const double *q_dev, uint64_t n, double *pose12_dev,
                          void *stream) {
  if (!w || (n && (!q_dev || !pose12_dev)))
    return fail(PRGPU_EINVAL, "fk_batch: NULL argument");
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  return fk_launch(w->host, w->dev, q_dev, n, pose12_dev, (cudaStream_t)stream);
}

int prgpu_fk_batch(prgpu_world_t *w, const double *q, uint64_t n, double *pose12) {
  if (!w || (n && (!q || !pose12)))
    return fail(PRGPU_EINVAL, "fk_batch: NULL argument");
  if (n == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  const size_t bytes = (size_t)n * w->host.dof * sizeof(double);
  int rc;
  if ((rc = w->d_a.reserve(bytes)) != PRGPU_OK || (rc = w->d_pose.reserve((size_t)n * 12 * sizeof(double))) != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(w->d_a.p, q, bytes, cudaMemcpyHostToDevice, w->stream));
  rc = fk_launch(w->host, w->dev, w->d_a.as<double>(), n, w->d_pose.as<double>(), w->stream);
  if (rc != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(pose12, w->d_pose.p, (size_t)n * 12 * sizeof(double),
                                 cudaMemcpyDeviceToHost, w->stream));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(w->stream));
  return PRGPU_OK;
}

/* ---------------------------------------------- T1 ---------------------------------------------- */
} // extern "C"

struct prgpu_rrtc_batch {
  prgpu_world *world = nullptr;
  prgpu_rrtc_params params;
  RrtcDeviceParams dp;
  uint32_t n = 0;
  DevBuf starts, goals, ss, gs, sp, gp, status, iters, ns, ng, cs, cg, ovf, resume;
  bool ran = false;
};

extern "C" {

int prgpu_rrtc_batch_create(prgpu_world_t *w, const prgpu_rrtc_params *params, uint32_t n_queries,
                            prgpu_rrtc_batch_t **out) {
  if (!out)
    return fail(PRGPU_EINVAL, "rrtc_batch_create: out is NULL");
  *out = nullptr;
  if (!w || !params)
    return fail(PRGPU_EINVAL, "rrtc_batch_create: NULL argument");
  if (params->dim != w->host.dof)
    return fail(PRGPU_EINVAL, "rrtc_batch_create: dim must equal the world's dof");
  if (params->max_nodes < 2 || params->max_iters < 1)
    return fail(PRGPU_EINVAL, "rrtc_batch_create: max_nodes >= 2 and max_iters >= 1 required");
  if (!(params->resolution_fraction > 0.0 && params->resolution_fraction < 1.0))
    return fail(PRGPU_EINVAL, "rrtc_batch_create: resolution_fraction must be in (0,1)");
  for (int e : {params->ext_first, params->ext_second})
    if (e != PRGPU_EXT_EXTEND && e != PRGPU_EXT_CONNECT)
      return fail(PRGPU_EINVAL, "rrtc_batch_create: Invalid extension type");
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  prgpu_rrtc_batch *h = new prgpu_rrtc_batch();
  h->world = w;
  h->params = *params;
  h->n = n_queries;
  RrtcDeviceParams &d = h->dp;
  std::memset(&d, 0, sizeof(d));
  d.dim = params->dim;
  double ext = 0.0; // RealVectorStateSpace::getMaximumExtent
  for (int i = 0; i < d.dim; ++i) {
    if (!(params->hi[i] >= params->lo[i])) {
      delete h;
      return fail(PRGPU_EINVAL, "rrtc_batch_create: bounds must satisfy lo <= hi");
    }
    d.lo[i] = params->lo[i];
    d.hi[i] = params->hi[i];
    const double e = params->hi[i] - params->lo[i];
    ext += e * e;
  }
  ext = std::sqrt(ext);
  d.range = params->range;
  if (d.range < 2.220446049250313e-16) // SelfConfig::configurePlannerRange
    d.range = ext * 0.2;
  d.lvs = ext * params->resolution_fraction;
  d.ext_first = params->ext_first;
  d.ext_second = params->ext_second;
  d.seed = params->seed;
  d.query_id_base = params->query_id_base;
  d.max_iters = params->max_iters;
  d.max_nodes = params->max_nodes;
  const size_t B = n_queries ? n_queries : 1, M = params->max_nodes, D = d.dim;
  int rc;
  if ((rc = h->starts.reserve(B * D * 8)) || (rc = h->goals.reserve(B * D * 8)) ||
      (rc = h->ss.reserve(B * M * D * 8)) || (rc = h->gs.reserve(B * M * D * 8)) ||
      (rc = h->sp.reserve(B * M * 4)) || (rc = h->gp.reserve(B * M * 4)) || (rc = h->status.reserve(B * 4)) ||
      (rc = h->iters.reserve(B * 4)) || (rc = h->ns.reserve(B * 4)) || (rc = h->ng.reserve(B * 4)) ||
      (rc = h->cs.reserve(B * 4)) || (rc = h->cg.reserve(B * 4)) || (rc = h->ovf.reserve((2 * B + 4) * 4)) ||
      (rc = h->resume.reserve(B * sizeof(RrtcLoopState)))) {
    prgpu_rrtc_batch_destroy(h);
    return rc;
  }
  *out = h;
  return PRGPU_OK;
}

int prgpu_rrtc_batch_destroy(prgpu_rrtc_batch_t *h) {
  if (!h)
    return PRGPU_OK;
  cudaSetDevice(h->world->device);
  cudaStreamSynchronize(h->world->stream);
  for (DevBuf *b : {&h->starts, &h->goals, &h->ss, &h->gs, &h->sp, &h->gp, &h->status, &h->iters, &h->ns,
                    &h->ng, &h->cs, &h->cg, &h->ovf, &h->resume})
    b->release();
  delete h;
  return PRGPU_OK;
}

int prgpu_rrtc_batch_run_device(prgpu_rrtc_batch_t *h, const double *starts_dev, const double *goals_dev,
                                void *stream) {
  if (!h || (h->n && (!starts_dev || !goals_dev)))
    return fail(PRGPU_EINVAL, "rrtc_batch_run: NULL argument");
  PRGPU_CUDA_TRY(cudaSetDevice(h->world->device));
  RrtcDeviceBuffers b;
  b.starts = starts_dev;
  b.goals = goals_dev;
  b.start_states = h->ss.as<double>();
  b.goal_states = h->gs.as<double>();
  b.start_parent = h->sp.as<int32_t>();
  b.goal_parent = h->gp.as<int32_t>();
  b.status = h->status.as<int32_t>();
  b.iters = h->iters.as<uint32_t>();
  b.n_start = h->ns.as<uint32_t>();
  b.n_goal = h->ng.as<uint32_t>();
  b.conn_start = h->cs.as<int32_t>();
  b.conn_goal = h->cg.as<int32_t>();
  b.overflow = h->ovf.as<uint32_t>();
  {
    const size_t B = h->n ? h->n : 1;
    b.work_counter = h->ovf.as<uint32_t>() + B;
    b.susp_count = b.work_counter + 1;
    b.team_counter = b.work_counter + 2;
    b.susp_list = b.work_counter + 4;
    b.resume = h->resume.as<RrtcLoopState>();
  }
  int rc = rrtc_batch_launch(h->world->host, h->world->dev, h->dp, b, h->n, h->world->sm_count, (cudaStream_t)stream);
  if (rc == PRGPU_OK)
    h->ran = true;
  return rc;
}

int prgpu_rrtc_batch_run(prgpu_rrtc_batch_t *h, const double *starts, const double *goals) {
  if (!h || (h->n && (!starts || !goals)))
    return fail(PRGPU_EINVAL, "rrtc_batch_run: NULL argument");
  if (h->n == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(h->world->device));
  cudaStream_t s = h->world->stream;
  const size_t bytes = (size_t)h->n * h->dp.dim * sizeof(double);
  PRGPU_CUDA_TRY(cudaMemcpyAsync(h->starts.p, starts, bytes, cudaMemcpyHostToDevice, s));
  PRGPU_CUDA_TRY(cudaMemcpyAsync(h->goals.p, goals, bytes, cudaMemcpyHostToDevice, s));
  int rc = prgpu_rrtc_batch_run_device(h, h->starts.as<double>(), h->goals.as<double>(), s);
  if (rc != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
  return PRGPU_OK;
}

int prgpu_rrtc_batch_summary(prgpu_rrtc_batch_t *h, int32_t *status, uint32_t *iterations,
                             uint32_t *n_start, uint32_t *n_goal) {
  if (!h)
    return fail(PRGPU_EINVAL, "rrtc_batch_summary: NULL handle");
  if (!h->ran)
    return fail(PRGPU_EINVAL, "rrtc_batch_summary: run() has not been called");
  PRGPU_CUDA_TRY(cudaSetDevice(h->world->device));
  const size_t b4 = (size_t)h->n * 4;
  if (status)
    PRGPU_CUDA_TRY(cudaMemcpy(status, h->status.p, b4, cudaMemcpyDeviceToHost));
  if (iterations)
    PRGPU_CUDA_TRY(cudaMemcpy(iterations, h->iters.p, b4, cudaMemcpyDeviceToHost));
  if (n_start)
    PRGPU_CUDA_TRY(cudaMemcpy(n_start, h->ns.p, b4, cudaMemcpyDeviceToHost));
  if (n_goal)
    PRGPU_CUDA_TRY(cudaMemcpy(n_goal, h->ng.p, b4, cudaMemcpyDeviceToHost));
  return PRGPU_OK;
}

int prgpu_rrtc_batch_fetch(prgpu_rrtc_batch_t *h, uint32_t query, double *start_states,
                           int32_t *start_parent, double *goal_states, int32_t *goal_parent,
                           uint32_t *path_len, double *path, uint32_t max_path) {
  if (!h)
    return fail(PRGPU_EINVAL, "rrtc_batch_fetch: NULL handle");
  if (!h->ran || query >= h->n)
    return fail(PRGPU_EINVAL, "rrtc_batch_fetch: no such query (or run() not called)");
  PRGPU_CUDA_TRY(cudaSetDevice(h->world->device));
  const size_t M = h->dp.max_nodes, D = h->dp.dim;
  uint32_t ns = 0, ng = 0;
  int32_t cs = -1, cg = -1;
  PRGPU_CUDA_TRY(cudaMemcpy(&ns, h->ns.as<uint32_t>() + query, 4, cudaMemcpyDeviceToHost));
  PRGPU_CUDA_TRY(cudaMemcpy(&ng, h->ng.as<uint32_t>() + query, 4, cudaMemcpyDeviceToHost));
  PRGPU_CUDA_TRY(cudaMemcpy(&cs, h->cs.as<int32_t>() + query, 4, cudaMemcpyDeviceToHost));
  PRGPU_CUDA_TRY(cudaMemcpy(&cg, h->cg.as<int32_t>() + query, 4, cudaMemcpyDeviceToHost));
  std::vector<double> S(ns * D), G(ng * D);
  std::vector<int32_t> PS(ns), PG(ng);
  if (ns) {
    PRGPU_CUDA_TRY(cudaMemcpy(S.data(), h->ss.as<double>() + query * M * D, ns * D * 8, cudaMemcpyDeviceToHost));
    PRGPU_CUDA_TRY(cudaMemcpy(PS.data(), h->sp.as<int32_t>() + query * M, ns * 4, cudaMemcpyDeviceToHost));
  }
  if (ng) {
    PRGPU_CUDA_TRY(cudaMemcpy(G.data(), h->gs.as<double>() + query * M * D, ng * D * 8, cudaMemcpyDeviceToHost));
    PRGPU_CUDA_TRY(cudaMemcpy(PG.data(), h->gp.as<int32_t>() + query * M, ng * 4, cudaMemcpyDeviceToHost));
  }
  if (start_states && ns)
    std::memcpy(start_states, S.data(), ns * D * 8);
  if (start_parent && ns)
    std::memcpy(start_parent, PS.data(), ns * 4);
  if (goal_states && ng)
    std::memcpy(goal_states, G.data(), ng * D * 8);
  if (goal_parent && ng)
    std::memcpy(goal_parent, PG.data(), ng * 4);
  // solution path: start chain reversed, then goal chain (pr_ompl/src/RRTConnect.cpp:338-363)
  uint32_t len = 0;
  if (cs >= 0 && cg >= 0) {
    std::vector<int32_t> c1, c2;
    for (int32_t s = cs; s >= 0; s = PS[s])
      c1.push_back(s);
    for (int32_t s = cg; s >= 0; s = PG[s])
      c2.push_back(s);
    len = (uint32_t)(c1.size() + c2.size());
    if (path) {
      if (len > max_path)
        return fail(PRGPU_ELIMIT, "rrtc_batch_fetch: path buffer too small");
      uint32_t o = 0;
      for (int i = (int)c1.size() - 1; i >= 0; --i)
        std::memcpy(path + (size_t)(o++) * D, S.data() + (size_t)c1[i] * D, D * 8);
      for (size_t i = 0; i < c2.size(); ++i)
        std::memcpy(path + (size_t)(o++) * D, G.data() + (size_t)c2[i] * D, D * 8);
    }
  }
  if (path_len)
    *path_len = len;
  return PRGPU_OK;
}

} // extern "C"

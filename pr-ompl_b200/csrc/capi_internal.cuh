// Handle structs shared by the C-ABI translation units (capi.cu, capi_nnf.cu).
#pragma once
#include <vector>
#include "check_motion.cuh"
#include "knn.cuh"
#include "kdtree.cuh"

namespace prgpu {
// grow-only device buffer
// Device memory of the handles' work buffers comes from the device's stream-ordered pool with the release threshold
// raised to "never" (capi.cu): a planner object is created and destroyed per planning query, and cudaMalloc / cudaFree
// of its ~30 buffers (up to GBs) cost 0.1 s per set-up at random; from the pool a re-created handle gets the previous
// one's memory back.  The calls keep cudaMalloc / cudaFree semantics (the memory is usable on every stream when
// dev_alloc returns; dev_free waits for the device first).  PRGPU_NO_POOL=1: plain cudaMalloc / cudaFree.
cudaError_t dev_alloc(void **p, size_t bytes);
void dev_free(void *p);
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  bool plain = false; // plain cudaMalloc memory (buffers handed to NCCL / other libraries)
  int reserve(size_t need) {
    if (need <= bytes)
      return PRGPU_OK;
    if (p)
      plain ? (void)cudaFree(p) : dev_free(p);
    p = nullptr;
    bytes = 0;
    PRGPU_CUDA_TRY(plain ? cudaMalloc(&p, need) : dev_alloc(&p, need));
    bytes = need;
    return PRGPU_OK;
  }
  // for tables that are rebuilt a little larger many times (the NNF band): grow geometrically, so that the number of
  // cudaFree + cudaMalloc pairs of several hundred MB -- each a device-wide sync, at times hundreds of ms -- stays
  // logarithmic
  int reserve_grow(size_t need) {
    if (need <= bytes)
      return PRGPU_OK;
    return reserve(need + need / 2 > 2 * bytes ? need + need / 2 : 2 * bytes);
  }
  void release() {
    if (p)
      plain ? (void)cudaFree(p) : dev_free(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T> T *as() { return static_cast<T *>(p); }
};
// A small window of mapped pinned host memory (zero copy).  A planner that drives the structures one call at a
// time -- the reference's own RRTConnect over GpuNearestNeighbors / GpuMotionValidator: one nearest(), one
// checkMotion(), one add() per growTree -- pays launch + copy latency, not bandwidth: three pageable cudaMemcpyAsync
// calls cost more than the kernel between them.  Tiny calls therefore put their arguments into this window, the
// kernels read them and write their results through the device alias of the same pages, and the one stream
// synchronisation the call needs anyway makes the results visible.
struct HostWindow {
  unsigned char *host = nullptr, *dev = nullptr;
  size_t bytes = 0;
  bool failed = false; // mapped allocations refused once: stay on the copy path
  bool ensure(size_t need) {
    if (host && bytes >= need)
      return true;
    if (failed)
      return false;
    release();
    void *p = nullptr, *d = nullptr;
    if (cudaHostAlloc(&p, need, cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) {
      cudaGetLastError();
      if (p)
        cudaFreeHost(p);
      failed = true;
      return false;
    }
    host = static_cast<unsigned char *>(p);
    dev = static_cast<unsigned char *>(d);
    bytes = need;
    return true;
  }
  void release() {
    if (host)
      cudaFreeHost(host);
    host = dev = nullptr;
    bytes = 0;
  }
};
} // namespace prgpu

struct prgpu_knn {
  int device = 0, dim = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  double *pts = nullptr; // SoA, dim x stride
  float *ptsf = nullptr; // fp32 mirror, stride x 8 (dim <= 7 only)
  double *xmax2 = nullptr; // device scalar: max |x|^2
  int force_exact = 0;             // search mode 1: exact brute-force kernel only
  int no_index = 0;                // search mode 2: brute force (filter / exact) without the k-d index
  int force_index = 0;             // search mode 3: the k-d index whenever it can answer (tests)
  prgpu::KdIndex kd;               // sub-linear index over nodes [0, kd.n); stale when kd_valid is false
  bool kd_valid = false;
  double kd_regret_ms = 0.0;       // what the eligible searches on the current node set have cost beyond an indexed
                                   // search so far (estimated); the index is built when this reaches its build cost
  unsigned long long kd_calls = 0;
  prgpu::KnnTreeView view() const {
    prgpu::KnnTreeView v;
    v.pts = pts;
    v.ptsf = ptsf;
    v.xmax2 = xmax2;
    v.stride = stride;
    v.n = (uint32_t)n;
    v.index_base = index_base;
    return v;
  }
  uint64_t stride = 0, n = 0;
  uint32_t index_base = 0;
  prgpu::KnnScratch scratch;
  prgpu::KernelTimer timer;
  prgpu::DevBuf stage, d_q, d_lo, d_idx, d_dist;
  cudaStream_t copy_stream = nullptr;       // uploads of bulk prgpu_knn_add calls
  std::vector<cudaEvent_t> copy_events;
  prgpu::HostWindow win;                    // tiny host-buffer calls (see HostWindow)
  uint32_t ring_used = 0;                   // points of small prgpu_knn_add calls whose append has not been synchronised
};
// layout of prgpu_knn::win
constexpr size_t KNN_WIN_Q = 0, KNN_WIN_LO = 4096, KNN_WIN_IDX = 8192, KNN_WIN_DIST = 16384, KNN_WIN_RING = 32768,
                 KNN_WIN_BYTES = 65536;
constexpr uint32_t KNN_WIN_MAX_Q = 64, KNN_WIN_RING_SLOTS = 512, KNN_WIN_MAX_NODES = 1u << 16;
// appends of small host adds may still be in flight on h->stream: wait for them before another stream uses the tree
int knn_settle(prgpu_knn *h);

struct prgpu_world {
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  prgpu::DeviceWorld host; // copy of the device struct (device pointers inside)
  prgpu::DeviceWorld *dev = nullptr;
  prgpu::DevBuf os_f, ob_f, os_d, ob_d, cell_start, cell_item, df;
  prgpu::DevBuf d_a, d_b, d_valid, d_pose;
  prgpu::CmScratch cm_scratch;
  prgpu::KernelTimer timer;
  cudaStream_t copy_stream = nullptr;       // uploads of the staged host-buffer checkMotion
  std::vector<cudaEvent_t> copy_events;
  prgpu::HostWindow win;                    // tiny host-buffer calls (see HostWindow)
};


// the k-d index as seen by the fused multi-GPU step (comm.cu): is it built and the right tool for this batch;
// search with the result rows written into the peers' windows
bool knn_index_ready(const prgpu_knn *h, uint32_t nq, int k);
int knn_index_search(prgpu_knn *h, const double *queries_dev, uint32_t nq, int k, cudaStream_t stream,
                     const prgpu::KdPeerOut *peers);
// search dispatch shared by the single-GPU and the sharded entry points (capi.cu)
int knn_dispatch(prgpu_knn *h, const double *queries_dev, uint32_t nq, int k, const uint32_t *lo_dev, uint32_t *idx_dev,
                 double *dist_dev, cudaStream_t stream, bool *pending);

// Multi-GPU plumbing behind the C ABI (include/prgpu.h, prgpu_comm_* and the *_sharded entry points):
// SURVEY.md §8(e).  One process per GPU; a communicator is bootstrapped from an NCCL unique id the host
// application broadcasts (MPI, a file, torch.distributed's store ...).  NCCL is bound at run time
// (dlopen of libnccl.so.2: whichever copy the process already holds -- e.g. the one bundled with torch --
// is reused), so libprgpu.so still loads on a box without NCCL when only one GPU is used.
//   * kNN over a node-sharded tree: every rank searches its shard for all queries into ONE record
//     [dist: nq*k f64 | idx: nq*k u32 | unproven-query count], a single all-gather moves the records,
//     and topk_merge_kernel merges them under the (distance, index) order -- identical to one GPU.
//     The search does not synchronise in the middle: the (almost always zero) unproven counts of ALL ranks
//     ride in the gathered records and are read once, after the merge is enqueued.
//   * edge batches: each rank validates its slice; one all-gather of 1 byte per edge.
//   * independent planning queries need no collective (each rank runs prgpu_rrtc_batch_* on its slice).
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <vector>
#include "capi_internal.cuh"
#include "check_motion.cuh"
#include "knn.cuh"

using namespace prgpu;

namespace {
// the few NCCL declarations used (nccl.h is not required at build time)
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.lib)
    return PRGPU_OK;
  void *lib = nullptr;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (lib)
      break;
  }
  if (!lib)
    return fail(PRGPU_ENODEV, std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror());
  NcclApi a;
  a.lib = lib;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
  a.AllGather = (decltype(a.AllGather))dlsym(lib, "ncclAllGather");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
  a.GetVersion = (decltype(a.GetVersion))dlsym(lib, "ncclGetVersion");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString)
    return fail(PRGPU_ENODEV, "NCCL library lacks a required symbol");
  g_nccl = a;
  return PRGPU_OK;
}
int nccl_fail(const char *what, ncclResult_t r) {
  return fail(PRGPU_ECUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}
constexpr int NCCL_INT8 = 0; // ncclInt8 / ncclChar
} // namespace

struct prgpu_comm {
  int device = -1, rank = 0, nranks = 1;
  ncclComm_t nccl = nullptr;
  prgpu_allgather_fn custom = nullptr;
  void *ctx = nullptr;
  DevBuf send, recv; // plain cudaMalloc memory (set at creation): these are handed to NCCL
  uint32_t *flags_host = nullptr; // pinned, nranks words
  uint64_t redo_steps = 0;
  // symmetric windows (prgpu_comm_enable_peer_windows): every rank's window mapped into this process over
  // CUDA IPC; two halves used alternately + one page of arrival flags
  char *win = nullptr;                  // this rank's window
  char *peer_win[KD_MAX_PEERS] = {};    // all ranks' windows as seen from here (own entry = win)
  uint64_t win_half = 0;                // bytes per half
  uint32_t win_epoch = 0;
  uint64_t fused_steps = 0;
};

namespace {
constexpr uint64_t WIN_FLAGS_BYTES = 4096;
// arrival flags: rank r's word `me` of the half's flag row is set to `epoch` by rank `me` when its rows are in
// r's window; then every rank waits for all words of its own row
__global__ void comm_signal_wait_kernel(KdPeerOut po, unsigned long long flags_off, uint32_t epoch) {
  const int r = threadIdx.x;
  __threadfence_system();
  if (r < po.nranks) {
    volatile uint32_t *theirs = reinterpret_cast<volatile uint32_t *>(po.base[r] + flags_off);
    theirs[po.rank] = epoch;
    __threadfence_system();
    volatile uint32_t *mine = reinterpret_cast<volatile uint32_t *>(po.base[po.rank] + flags_off);
    while (mine[r] != epoch)
      __nanosleep(100);
  }
  __threadfence_system();
}
} // namespace

static int comm_allgather(prgpu_comm *c, const void *send, void *recv, uint64_t bytes, cudaStream_t stream) {
  if (c->custom) {
    const int rc = c->custom(c->ctx, send, recv, bytes, (void *)stream);
    return rc == 0 ? PRGPU_OK : fail(PRGPU_ECUDA, "custom all-gather callback failed");
  }
  const ncclResult_t r = g_nccl.AllGather(send, recv, (size_t)bytes, NCCL_INT8, c->nccl, stream);
  return r == 0 ? PRGPU_OK : nccl_fail("ncclAllGather", r);
}

extern "C" {

int prgpu_comm_unique_id(void *id128) {
  if (!id128)
    return fail(PRGPU_EINVAL, "comm_unique_id: NULL argument");
  int rc = nccl_load();
  if (rc != PRGPU_OK)
    return rc;
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0)
    return nccl_fail("ncclGetUniqueId", r);
  std::memcpy(id128, &id, sizeof(id));
  return PRGPU_OK;
}

int prgpu_comm_init(int device, int rank, int nranks, const void *id128, prgpu_comm_t **out) {
  if (!out)
    return fail(PRGPU_EINVAL, "comm_init: out is NULL");
  *out = nullptr;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks)
    return fail(PRGPU_EINVAL, "comm_init: bad rank / nranks / id");
  int rc = select_device(device, nullptr);
  if (rc != PRGPU_OK)
    return rc;
  if ((rc = nccl_load()) != PRGPU_OK)
    return rc;
  prgpu_comm *c = new prgpu_comm();
  c->send.plain = c->recv.plain = true;
  c->device = device;
  c->rank = rank;
  c->nranks = nranks;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  const ncclResult_t r = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
  if (r != 0) {
    delete c;
    return nccl_fail("ncclCommInitRank", r);
  }
  if (cudaMallocHost(&c->flags_host, sizeof(uint32_t) * nranks) != cudaSuccess) {
    g_nccl.CommDestroy(c->nccl);
    delete c;
    return fail(PRGPU_ENOMEM, "comm_init: pinned allocation failed");
  }
  *out = c;
  return PRGPU_OK;
}

int prgpu_comm_init_custom(int device, int rank, int nranks, prgpu_allgather_fn fn, void *ctx, prgpu_comm_t **out) {
  if (!out)
    return fail(PRGPU_EINVAL, "comm_init_custom: out is NULL");
  *out = nullptr;
  if (!fn || nranks < 1 || rank < 0 || rank >= nranks)
    return fail(PRGPU_EINVAL, "comm_init_custom: bad rank / nranks / callback");
  if (device >= 0) {
    int rc = select_device(device, nullptr);
    if (rc != PRGPU_OK)
      return rc;
  }
  prgpu_comm *c = new prgpu_comm();
  c->send.plain = c->recv.plain = true;
  c->device = device; // -1: host-only communicator (plumbing tests without a GPU); compute entry points refuse it
  c->rank = rank;
  c->nranks = nranks;
  c->custom = fn;
  c->ctx = ctx;
  if (device >= 0 && cudaMallocHost(&c->flags_host, sizeof(uint32_t) * nranks) != cudaSuccess) {
    delete c;
    return fail(PRGPU_ENOMEM, "comm_init_custom: pinned allocation failed");
  }
  *out = c;
  return PRGPU_OK;
}

int prgpu_comm_destroy(prgpu_comm_t *c) {
  if (!c)
    return PRGPU_OK;
  if (c->device >= 0) {
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->send.release();
    c->recv.release();
    if (c->flags_host)
      cudaFreeHost(c->flags_host);
    for (int r = 0; r < c->nranks && r < KD_MAX_PEERS; ++r)
      if (c->peer_win[r] && r != c->rank)
        cudaIpcCloseMemHandle(c->peer_win[r]);
    if (c->win)
      cudaFree(c->win);
  }
  if (c->nccl)
    g_nccl.CommDestroy(c->nccl);
  delete c;
  return PRGPU_OK;
}

int prgpu_comm_rank(const prgpu_comm_t *c) { return c ? c->rank : -1; }
int prgpu_comm_size(const prgpu_comm_t *c) { return c ? c->nranks : 0; }

int prgpu_comm_allgather(prgpu_comm_t *c, const void *send, void *recv, uint64_t bytes_per_rank, void *stream) {
  if (!c || !send || !recv)
    return fail(PRGPU_EINVAL, "comm_allgather: NULL argument");
  if (c->device >= 0)
    PRGPU_CUDA_TRY(cudaSetDevice(c->device));
  return comm_allgather(c, send, recv, bytes_per_rank, (cudaStream_t)stream);
}

int prgpu_comm_stats(const prgpu_comm_t *c, uint64_t *redo_steps) {
  if (!c)
    return fail(PRGPU_EINVAL, "comm_stats: NULL handle");
  if (redo_steps)
    *redo_steps = c->redo_steps;
  return PRGPU_OK;
}

int prgpu_knn_nearest_k_sharded_device(prgpu_knn_t *h, prgpu_comm_t *c, const double *queries_dev, uint64_t nq, int k,
                                       uint32_t *idx_dev, double *dist_dev, void *stream_) {
  if (!h || !c || (nq && (!queries_dev || !idx_dev || !dist_dev)))
    return fail(PRGPU_EINVAL, "knn_nearest_k_sharded: NULL argument");
  if (c->device < 0 || c->device != h->device)
    return fail(PRGPU_EINVAL, "knn_nearest_k_sharded: communicator and tree live on different devices");
  if (k < 1 || k > PRGPU_MAX_K)
    return fail(PRGPU_ELIMIT, "knn_nearest_k_sharded: k must be in [1,PRGPU_MAX_K]");
  if (nq == 0)
    return PRGPU_OK;
  if (nq > 0x7FFFFFFFull)
    return fail(PRGPU_ELIMIT, "knn_nearest_k_sharded: too many queries in one call");
  if ((cudaStream_t)stream_ != h->stream) { // small host adds may still be appending on the handle's stream
    const int rs = knn_settle(h);
    if (rs != PRGPU_OK)
      return rs;
  }
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t ent = (size_t)nq * k;
  // record: [dist f64 x ent | idx u32 x ent | flag u32 | pad] ; 16-byte multiple so that every rank's
  // dist block stays 8-byte aligned inside the gathered buffer
  const size_t flag_off = ent * 12;
  const size_t chunk = (flag_off + 4 + 15) / 16 * 16;
  int rc;
  if ((rc = c->send.reserve(chunk)) != PRGPU_OK || (rc = c->recv.reserve(chunk * c->nranks)) != PRGPU_OK)
    return rc;
  char *send = c->send.as<char>(), *recv = c->recv.as<char>();
  double *ld = reinterpret_cast<double *>(send);
  uint32_t *li = reinterpret_cast<uint32_t *>(send + ent * 8);
  auto step = [&](bool redo, uint32_t my_flags) -> int {
    bool pending = false;
    int r2;
    if (!redo) {
      r2 = knn_dispatch(h, queries_dev, (uint32_t)nq, k, nullptr, li, ld, stream, &pending);
    } else {
      r2 = knn_search_finish(h->dim, h->view(), queries_dev, (uint32_t)nq, k, nullptr, li, ld, h->scratch, h->sm_count,
                             stream, my_flags);
    }
    if (r2 != PRGPU_OK)
      return r2;
    if (pending)
      PRGPU_CUDA_TRY(cudaMemcpyAsync(send + flag_off, h->scratch.flags, 4, cudaMemcpyDeviceToDevice, stream));
    else
      PRGPU_CUDA_TRY(cudaMemsetAsync(send + flag_off, 0, 4, stream));
    if (c->nranks == 1) {
      PRGPU_CUDA_TRY(cudaMemcpyAsync(recv, send, chunk, cudaMemcpyDeviceToDevice, stream));
    } else if ((r2 = comm_allgather(c, send, recv, chunk, stream)) != PRGPU_OK)
      return r2;
    return topk_merge_strided(reinterpret_cast<const uint32_t *>(recv + ent * 8), reinterpret_cast<const double *>(recv),
                              chunk / 4, chunk / 8, c->nranks, (uint32_t)nq, k, idx_dev, dist_dev, stream);
  };
  if ((rc = step(false, 0)) != PRGPU_OK)
    return rc;
  // every rank's unproven-query count, read once behind the merge
  PRGPU_CUDA_TRY(cudaMemcpy2DAsync(c->flags_host, 4, recv + flag_off, chunk, 4, c->nranks, cudaMemcpyDeviceToHost,
                                   stream));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(stream));
  bool any = false;
  for (int r = 0; r < c->nranks; ++r)
    any = any || c->flags_host[r] != 0;
  if (!any)
    return PRGPU_OK;
  // rare: some shard could not prove some queries.  Every rank saw the same counts, so every rank takes
  // this branch: fix the local rows with the exact kernel, gather and merge again.
  ++c->redo_steps;
  return step(true, c->flags_host[c->rank]);
}

int prgpu_comm_enable_peer_windows(prgpu_comm_t *c, uint64_t bytes) {
  if (!c || bytes == 0)
    return fail(PRGPU_EINVAL, "comm_enable_peer_windows: NULL / empty");
  if (c->device < 0 || c->custom || c->nranks > KD_MAX_PEERS)
    return fail(PRGPU_EINVAL, "comm_enable_peer_windows: needs an NCCL communicator on a device (<= 16 ranks)");
  if (c->win)
    return bytes <= c->win_half ? PRGPU_OK : fail(PRGPU_EINVAL, "comm_enable_peer_windows: already enabled, smaller");
  PRGPU_CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t half = (bytes + 4095) / 4096 * 4096;
  char *win = nullptr;
  PRGPU_CUDA_TRY(cudaMalloc(&win, 2 * half + 2 * WIN_FLAGS_BYTES));
  PRGPU_CUDA_TRY(cudaMemset(win, 0, 2 * half + 2 * WIN_FLAGS_BYTES));
  // exchange the IPC handles through the communicator's own all-gather
  cudaIpcMemHandle_t mine;
  cudaError_t e = cudaIpcGetMemHandle(&mine, win);
  int rc = e == cudaSuccess ? PRGPU_OK : fail(PRGPU_ECUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  const size_t hb = 64;
  static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle size");
  std::vector<char> all(hb * c->nranks, 0);
  if (rc == PRGPU_OK && ((rc = c->send.reserve(hb)) != PRGPU_OK || (rc = c->recv.reserve(hb * c->nranks)) != PRGPU_OK)) {
  }
  if (rc == PRGPU_OK) {
    char tmp[64] = {0};
    std::memcpy(tmp, &mine, sizeof(mine));
    cudaMemcpy(c->send.p, tmp, hb, cudaMemcpyHostToDevice);
    rc = comm_allgather(c, c->send.p, c->recv.p, hb, nullptr);
    if (rc == PRGPU_OK && cudaDeviceSynchronize() != cudaSuccess)
      rc = fail(PRGPU_ECUDA, "comm_enable_peer_windows: handle exchange failed");
    if (rc == PRGPU_OK)
      cudaMemcpy(all.data(), c->recv.p, hb * c->nranks, cudaMemcpyDeviceToHost);
  }
  // every rank must learn whether EVERY rank could map every window (else nobody uses them): count successes
  int ok = rc == PRGPU_OK ? 1 : 0;
  for (int r = 0; ok && r < c->nranks; ++r) {
    if (r == c->rank) {
      c->peer_win[r] = win;
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, all.data() + hb * r, sizeof(hd));
    void *p = nullptr;
    e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      fail(PRGPU_ECUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
      break;
    }
    c->peer_win[r] = static_cast<char *>(p);
  }
  char okb[64] = {0};
  okb[0] = (char)ok;
  std::vector<char> oks(hb * c->nranks, 0);
  cudaMemcpy(c->send.p, okb, hb, cudaMemcpyHostToDevice);
  if (comm_allgather(c, c->send.p, c->recv.p, hb, nullptr) == PRGPU_OK && cudaDeviceSynchronize() == cudaSuccess)
    cudaMemcpy(oks.data(), c->recv.p, hb * c->nranks, cudaMemcpyDeviceToHost);
  else
    ok = 0;
  for (int r = 0; r < c->nranks; ++r)
    ok = ok && oks[hb * r] != 0;
  if (!ok) {
    for (int r = 0; r < c->nranks; ++r) {
      if (c->peer_win[r] && r != c->rank)
        cudaIpcCloseMemHandle(c->peer_win[r]);
      c->peer_win[r] = nullptr;
    }
    cudaFree(win);
    return fail(PRGPU_ENODEV, "comm_enable_peer_windows: peer memory is not available between these ranks "
                              "(the all-gather path stays in use)");
  }
  c->win = win;
  c->win_half = half;
  return PRGPU_OK;
}

int prgpu_comm_stats2(const prgpu_comm_t *c, uint64_t *fused_steps) {
  if (!c)
    return fail(PRGPU_EINVAL, "comm_stats2: NULL handle");
  if (fused_steps)
    *fused_steps = c->fused_steps;
  return PRGPU_OK;
}

int prgpu_knn_nearest_k_query_sharded_device(prgpu_knn_t *h, prgpu_comm_t *c, const double *queries_dev, uint64_t nq_local,
                                             int k, uint32_t *idx_all_dev, double *dist_all_dev, void *stream_) {
  if (!h || !c || (nq_local && (!queries_dev || !idx_all_dev || !dist_all_dev)))
    return fail(PRGPU_EINVAL, "knn_nearest_k_query_sharded: NULL argument");
  if (c->device < 0 || c->device != h->device)
    return fail(PRGPU_EINVAL, "knn_nearest_k_query_sharded: communicator and tree live on different devices");
  if (k < 1 || k > PRGPU_MAX_K)
    return fail(PRGPU_ELIMIT, "knn_nearest_k_query_sharded: k must be in [1,PRGPU_MAX_K]");
  if (nq_local == 0)
    return PRGPU_OK;
  if (nq_local > 0x7FFFFFFFull)
    return fail(PRGPU_ELIMIT, "knn_nearest_k_query_sharded: too many queries in one call");
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t ent = (size_t)nq_local * k;
  int rc;
  if (stream != h->stream && (rc = knn_settle(h)) != PRGPU_OK) // small host adds may still be appending
    return rc;
  // ---- fused form: the index search writes every result row straight into every rank's window (NVLink peer
  //      stores from the search kernel's epilogue); then one tiny kernel exchanges arrival flags; no collective
  //      call and no staging copy on the path.  Taken when the windows are mapped and the index answers.
  if (c->win && c->nranks > 1 && knn_index_ready(h, (uint32_t)nq_local, k) &&
      (uint64_t)ent * 12 * c->nranks <= c->win_half) {
    const uint32_t epoch = ++c->win_epoch;
    const uint64_t half_off = (epoch & 1u) * c->win_half;
    KdPeerOut po;
    std::memset(&po, 0, sizeof(po));
    for (int r = 0; r < c->nranks; ++r)
      po.base[r] = c->peer_win[r] + half_off;
    po.idx_off = (unsigned long long)ent * 8 * c->nranks;
    po.nranks = c->nranks;
    po.rank = c->rank;
    if ((rc = knn_index_search(h, queries_dev, (uint32_t)nq_local, k, stream, &po)) != PRGPU_OK)
      return rc;
    KdPeerOut pf = po; // the flag rows sit behind the two halves
    for (int r = 0; r < c->nranks; ++r)
      pf.base[r] = c->peer_win[r];
    comm_signal_wait_kernel<<<1, 32, 0, stream>>>(pf, 2 * c->win_half + (epoch & 1u) * WIN_FLAGS_BYTES, epoch); PRGPU_COUNT_LAUNCH(1);
    PRGPU_CUDA_TRY(cudaGetLastError());
    PRGPU_CUDA_TRY(cudaMemcpyAsync(dist_all_dev, c->win + half_off, ent * 8 * c->nranks, cudaMemcpyDeviceToDevice, stream));
    PRGPU_CUDA_TRY(cudaMemcpyAsync(idx_all_dev, c->win + half_off + po.idx_off, ent * 4 * c->nranks,
                                   cudaMemcpyDeviceToDevice, stream));
    ++c->fused_steps;
    return PRGPU_OK;
  }
  const size_t chunk = (ent * 12 + 15) / 16 * 16; // [dist f64 x ent | idx u32 x ent], 16-byte multiple
  if ((rc = c->send.reserve(chunk)) != PRGPU_OK || (rc = c->recv.reserve(chunk * c->nranks)) != PRGPU_OK)
    return rc;
  char *send = c->send.as<char>(), *recv = c->recv.as<char>();
  // this rank's queries against the WHOLE tree it holds (index, filter or exact path: knn_dispatch decides)
  if ((rc = knn_dispatch(h, queries_dev, (uint32_t)nq_local, k, nullptr, reinterpret_cast<uint32_t *>(send + ent * 8),
                         reinterpret_cast<double *>(send), stream, nullptr)) != PRGPU_OK)
    return rc;
  if (c->nranks == 1) {
    PRGPU_CUDA_TRY(cudaMemcpyAsync(recv, send, chunk, cudaMemcpyDeviceToDevice, stream));
  } else if ((rc = comm_allgather(c, send, recv, chunk, stream)) != PRGPU_OK)
    return rc;
  // records -> the two result tables, rank-major
  PRGPU_CUDA_TRY(cudaMemcpy2DAsync(dist_all_dev, ent * 8, recv, chunk, ent * 8, c->nranks, cudaMemcpyDeviceToDevice, stream));
  PRGPU_CUDA_TRY(cudaMemcpy2DAsync(idx_all_dev, ent * 4, recv + ent * 8, chunk, ent * 4, c->nranks,
                                   cudaMemcpyDeviceToDevice, stream));
  return PRGPU_OK;
}

int prgpu_knn_nearest_k_query_sharded(prgpu_knn_t *h, prgpu_comm_t *c, const double *queries, uint64_t nq_local, int k,
                                      uint32_t *idx_all, double *dist_all) {
  if (!h || !c || (nq_local && (!queries || !idx_all)))
    return fail(PRGPU_EINVAL, "knn_nearest_k_query_sharded: NULL argument");
  if (nq_local == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  const size_t all = (size_t)nq_local * k * (size_t)c->nranks;
  int rc;
  if ((rc = h->d_q.reserve((size_t)nq_local * h->dim * sizeof(double))) != PRGPU_OK ||
      (rc = h->d_idx.reserve(all * sizeof(uint32_t))) != PRGPU_OK || (rc = h->d_dist.reserve(all * sizeof(double))) != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(h->d_q.p, queries, (size_t)nq_local * h->dim * sizeof(double), cudaMemcpyHostToDevice,
                                 h->stream));
  rc = prgpu_knn_nearest_k_query_sharded_device(h, c, h->d_q.as<double>(), nq_local, k, h->d_idx.as<uint32_t>(),
                                                h->d_dist.as<double>(), h->stream);
  if (rc != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(idx_all, h->d_idx.p, all * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  if (dist_all)
    PRGPU_CUDA_TRY(cudaMemcpyAsync(dist_all, h->d_dist.p, all * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PRGPU_OK;
}

int prgpu_knn_nearest_k_sharded(prgpu_knn_t *h, prgpu_comm_t *c, const double *queries, uint64_t nq, int k,
                                uint32_t *idx, double *dist) {
  if (!h || !c || (nq && (!queries || !idx)))
    return fail(PRGPU_EINVAL, "knn_nearest_k_sharded: NULL argument");
  if (nq == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = h->d_q.reserve((size_t)nq * h->dim * sizeof(double))) != PRGPU_OK ||
      (rc = h->d_idx.reserve((size_t)nq * k * sizeof(uint32_t))) != PRGPU_OK ||
      (rc = h->d_dist.reserve((size_t)nq * k * sizeof(double))) != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(h->d_q.p, queries, (size_t)nq * h->dim * sizeof(double), cudaMemcpyHostToDevice,
                                 h->stream));
  rc = prgpu_knn_nearest_k_sharded_device(h, c, h->d_q.as<double>(), nq, k, h->d_idx.as<uint32_t>(),
                                          h->d_dist.as<double>(), h->stream);
  if (rc != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(idx, h->d_idx.p, (size_t)nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  if (dist)
    PRGPU_CUDA_TRY(cudaMemcpyAsync(dist, h->d_dist.p, (size_t)nq * k * sizeof(double), cudaMemcpyDeviceToHost,
                                   h->stream));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return PRGPU_OK;
}

int prgpu_check_motion_batch_sharded_device(prgpu_world_t *w, prgpu_comm_t *c, const double *from_dev,
                                            const double *to_dev, uint64_t n, double resolution, int mode,
                                            int check_first, uint8_t *valid_dev, void *stream_) {
  if (!w || !c || (n && (!from_dev || !to_dev || !valid_dev)))
    return fail(PRGPU_EINVAL, "check_motion_batch_sharded: NULL argument");
  if (c->device < 0 || c->device != w->device)
    return fail(PRGPU_EINVAL, "check_motion_batch_sharded: communicator and world live on different devices");
  if (n == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  const uint64_t slice = ((n + c->nranks - 1) / c->nranks + 15) / 16 * 16; // edges per rank, padded
  const uint64_t e0 = std::min<uint64_t>(n, slice * c->rank), e1 = std::min<uint64_t>(n, e0 + slice);
  int rc;
  if ((rc = c->send.reserve(slice)) != PRGPU_OK || (rc = c->recv.reserve(slice * c->nranks)) != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemsetAsync(c->send.p, 0, slice, stream));
  const int dof = w->host.dof;
  if (e1 > e0) {
    rc = check_motion_launch(w->host, w->dev, from_dev + e0 * dof, to_dev + e0 * dof, e1 - e0, resolution, mode,
                             check_first, c->send.as<uint8_t>(), w->sm_count, stream, &w->timer, &w->cm_scratch);
    if (rc != PRGPU_OK)
      return rc;
  }
  if (c->nranks == 1) {
    PRGPU_CUDA_TRY(cudaMemcpyAsync(valid_dev, c->send.p, n, cudaMemcpyDeviceToDevice, stream));
    return PRGPU_OK;
  }
  if ((rc = comm_allgather(c, c->send.p, c->recv.p, slice, stream)) != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(valid_dev, c->recv.p, n, cudaMemcpyDeviceToDevice, stream));
  return PRGPU_OK;
}
}

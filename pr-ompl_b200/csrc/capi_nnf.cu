// C ABI of the NNF path (include/prgpu.h, prgpu_nnf_*): host-side graph bookkeeping in flat
// arrays around the K1 / K2 / K3 kernels.  What the reference keeps in three Boost graphs with
// string-keyed hash maps (frechet/src/NNFrechet.cpp:467-599) is here: vertex ids in creation
// order, an edge list in creation order, a CSR of predecessors and a level-sorted vertex order.
#include <algorithm>
#include <cstdlib>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstring>
#include <numeric>
#include <vector>
#include "capi_internal.cuh"
#include "nnf.cuh"

using namespace prgpu;

struct prgpu_nnf {
  prgpu_world *world = nullptr; // NULL in the host-callback form (prgpu_nnf_create_graph)
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // host-callback form: node weights handed over by the caller, edge verdicts from a callback
  bool host_form = false, weights_set = false;
  std::vector<double> ftab_host; // V x R by vertex id
  DevBuf d_ftab;                 // V x R by sorted position
  prgpu_nnf_edge_eval_fn eval_fn = nullptr;
  void *eval_user = nullptr;
  prgpu_nnf_params params;
  int R = 0, dof = 0;
  uint32_t V = 0, E = 0, nIk = 0, Lv = 0;
  // host graph
  std::vector<double> ref_pose;                      // R x 12
  std::vector<uint32_t> knn;                         // nIk x k (vertex ids)
  std::vector<std::pair<uint32_t, uint32_t>> edges;  // creation order
  std::vector<uint32_t> pred_off_v;                  // V+1, by vertex id
  std::vector<std::pair<uint32_t, uint32_t>> pred_v; // (pred vertex, edge id), creation order
  std::vector<uint32_t> pos_of_vertex, vertex_of_pos;
  std::vector<double> states, pose;                  // V x dof, V x 12 (downloaded)
  std::vector<uint8_t> evaluated, blocked;           // per edge (reference semantics)
  std::vector<int8_t> known;                         // per edge: -1 unknown, 0 collides, 1 free
  // device
  DevBuf d_states, d_pose, d_lvl_start, d_level, d_pred_off, d_pred_pos, d_pred_eid, d_blocked, d_pnn, d_pref, d_B,
      d_path, d_pn, d_cost;
  NnfGraphDev g;
  KernelTimer timer;
  // banded form
  int mode = 0;                 // 0 banded dataflow DP (default), 1 full-table wavefront, 2 banded + cluster barriers
  std::vector<uint32_t> lvl_start_host;
  DevBuf d_done;
  uint32_t epoch = 0;
  std::vector<uint32_t> level_of_vertex, slot_of_eid;
  DevBuf d_band_lo, d_band_w, d_col_off, d_meta, d_edge, d_Bc, d_scan_tmp, d_rowmin;
  NnfBandDev bg;
  double tau_ub = 0.0;
  bool band_valid = false, band_full = false;
  uint32_t dirty_level = 0;
  uint64_t band_cells = 0;
  uint32_t band_rebuilds = 0;
  uint32_t total_iters = 0, n_evaluated = 0, n_collisions = 0;
  uint64_t knn_filter_calls = 0, knn_fallback_queries = 0;
  // device-resident LazySP loop (mode 0 with a device world): the per-edge LazySP state lives on the device
  // while that loop runs; the host mirrors (known / evaluated / blocked) are refreshed on demand
  DevBuf d_succ_off, d_succ_pos, d_succ_pb, d_vop, d_stamp, d_worklist, d_known, d_evaluated, d_hop, d_path_pos, d_path_slot,
      d_path_cell, d_out_path, d_trace, d_lazy_out;
  bool host_state_current = true, dev_state_current = false;
  uint32_t lazy_epoch = 0;
  uint64_t repaired_vertices = 0, lazy_launches = 0;
  uint64_t lazy_cycles[5] = {0, 0, 0, 0, 0};
  uint64_t lazy_bailouts = 0;
};

namespace {
template <class T> int upload(DevBuf &b, const std::vector<T> &v, cudaStream_t s) {
  int rc = b.reserve(std::max<size_t>(v.size() * sizeof(T), 16));
  if (rc != PRGPU_OK)
    return rc;
  if (!v.empty())
    PRGPU_CUDA_TRY(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return PRGPU_OK;
}
double host_task_distance(const double *a12, const double *b12) {
  volatile double dx = a12[3] - b12[3], dy = a12[7] - b12[7], dz = a12[11] - b12[11];
  volatile double xx = dx * dx, yy = dy * dy, zz = dz * dz;
  volatile double s = xx + yy;
  s = s + zz;
  return std::sqrt(s);
}
__global__ void nnf_block_edge_kernel(NnfEdgeRec *edge, uint32_t slot, uint8_t *blocked, uint32_t eid) {
  edge[slot].band_w_blocked |= 0x80000000u;
  blocked[eid] = 1;
}

// (re)build the band for the current tau_ub: per-vertex row interval, column offsets, packed records
// host copies of the vertex states (and, device form, FK poses), fetched on first use
int nnf_fetch_vertices(prgpu_nnf *h) {
  if (!h->states.empty() || h->V == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  std::vector<double> st((size_t)h->V * h->dof), po;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(st.data(), h->d_states.p, st.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  if (!h->host_form) {
    po.resize((size_t)h->V * 12);
    PRGPU_CUDA_TRY(cudaMemcpyAsync(po.data(), h->d_pose.p, po.size() * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->states.swap(st);
  h->pose.swap(po);
  return PRGPU_OK;
}

int nnf_build_band(prgpu_nnf *h, cudaStream_t s) {
  const int R = h->R;
  const uint32_t V = h->V;
  int rc = nnf_band_launch(R, V, h->g.pnn, h->g.pref, h->g.ftab, h->tau_ub, h->d_band_lo.as<uint32_t>(),
                           h->d_band_w.as<unsigned long long>(), s);
  if (rc != PRGPU_OK)
    return rc;
  size_t tmp = 0;
  rc = nnf_exclusive_scan_u64(h->d_band_w.as<unsigned long long>(), h->d_col_off.as<unsigned long long>(), V, nullptr,
                              tmp, s);
  if (rc != PRGPU_OK)
    return rc;
  if ((rc = h->d_scan_tmp.reserve(std::max<size_t>(tmp, 16))) != PRGPU_OK)
    return rc;
  rc = nnf_exclusive_scan_u64(h->d_band_w.as<unsigned long long>(), h->d_col_off.as<unsigned long long>(), V,
                              h->d_scan_tmp.p, tmp, s);
  if (rc != PRGPU_OK)
    return rc;
  unsigned long long last_off = 0, last_w = 0;
  PRGPU_CUDA_TRY(cudaMemcpyAsync(&last_off, h->d_col_off.as<unsigned long long>() + (V - 1), 8,
                                 cudaMemcpyDeviceToHost, s));
  PRGPU_CUDA_TRY(cudaMemcpyAsync(&last_w, h->d_band_w.as<unsigned long long>() + (V - 1), 8, cudaMemcpyDeviceToHost, s));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
  h->band_cells = last_off + last_w;
  h->band_full = h->band_cells == (uint64_t)R * V;
  // The band is rebuilt a little wider several times per solve (the goal cost creeps up with every knocked-out
  // edge).  A cudaFree + cudaMalloc pair of several hundred MB is a device-wide sync that at times takes hundreds of
  // ms, so the two cell tables are sized ONCE for the whole R x V table when that is a small part of the free memory
  // (3.9 GB each at the benchmark size, of 180 GB), and grow geometrically otherwise.
  size_t want = std::max<size_t>((size_t)h->band_cells * 8, 16);
  if (want > h->d_Bc.bytes) {
    size_t free_b = 0, total_b = 0;
    const size_t full = (size_t)R * V * 8;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && full <= free_b / 8)
      want = full;
  }
  if ((rc = h->d_Bc.reserve_grow(want)) != PRGPU_OK)
    return rc;
  h->bg.Bc = h->d_Bc.as<double>();
  h->bg.hop = nullptr;
  if (h->band_cells < (1ull << 36) && h->V < NNF_HOP_MAX_V) {
    if ((rc = h->d_hop.reserve_grow(std::max<size_t>(want, h->d_Bc.bytes))) != PRGPU_OK)
      return rc;
    h->bg.hop = h->d_hop.as<unsigned long long>();
  }
  rc = nnf_band_pack_launch(V, h->E, h->d_band_lo.as<uint32_t>(), h->d_band_w.as<unsigned long long>(),
                            h->d_col_off.as<unsigned long long>(), h->g.pred_off, h->g.pred_pos, h->g.pred_eid,
                            h->g.blocked, h->d_meta.as<NnfVertexMeta>(), h->d_edge.as<NnfEdgeRec>(), s);
  if (rc != PRGPU_OK)
    return rc;
  h->band_valid = true;
  h->dirty_level = 0;
  ++h->band_rebuilds;
  return PRGPU_OK;
}

// per-edge LazySP state: host mirrors <-> device arrays of the device-resident loop
int nnf_state_to_host(prgpu_nnf *h) {
  if (h->host_state_current)
    return PRGPU_OK;
  cudaStream_t s = h->stream;
  if (h->E) {
    PRGPU_CUDA_TRY(cudaMemcpyAsync(h->known.data(), h->d_known.p, h->E, cudaMemcpyDeviceToHost, s));
    PRGPU_CUDA_TRY(cudaMemcpyAsync(h->evaluated.data(), h->d_evaluated.p, h->E, cudaMemcpyDeviceToHost, s));
    PRGPU_CUDA_TRY(cudaMemcpyAsync(h->blocked.data(), h->d_blocked.p, h->E, cudaMemcpyDeviceToHost, s));
    PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
  }
  h->host_state_current = true;
  return PRGPU_OK;
}
int nnf_state_to_device(prgpu_nnf *h) {
  if (h->dev_state_current)
    return PRGPU_OK;
  cudaStream_t s = h->stream;
  if (h->E) { // (blocked edges are mirrored on the device by both loops as they happen)
    PRGPU_CUDA_TRY(cudaMemcpyAsync(h->d_known.p, h->known.data(), h->E, cudaMemcpyHostToDevice, s));
    PRGPU_CUDA_TRY(cudaMemcpyAsync(h->d_evaluated.p, h->evaluated.data(), h->E, cudaMemcpyHostToDevice, s));
    PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
  }
  h->dev_state_current = true;
  return PRGPU_OK;
}

// lower bound on the optimum: every start-goal path visits every reference row
int nnf_initial_tau(prgpu_nnf *h, cudaStream_t s) {
  int rc = nnf_rowmin_launch(h->R, h->V, h->g.pnn, h->g.pref, h->g.ftab, h->d_rowmin.as<unsigned long long>(), s);
  if (rc != PRGPU_OK)
    return rc;
  std::vector<double> rowmin(h->R);
  PRGPU_CUDA_TRY(cudaMemcpyAsync(rowmin.data(), h->d_rowmin.p, (size_t)h->R * 8, cudaMemcpyDeviceToHost, s));
  PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
  double lb = 0.0;
  for (double x : rowmin)
    lb = std::max(lb, x);
  h->tau_ub = std::max(lb * 1.5, 1e-6);
  return PRGPU_OK;
}
} // namespace

extern "C" {

int prgpu_nnf_set_mode(prgpu_nnf_t *h, int mode) {
  if (!h || mode < 0 || mode > 2)
    return fail(PRGPU_EINVAL, "nnf_set_mode: mode must be 0 (banded dataflow), 1 (full table) or 2 (banded, cluster barriers)");
  h->mode = mode;
  h->dirty_level = 0; // the other mode's table may be stale
  return PRGPU_OK;
}

int prgpu_nnf_band_info(prgpu_nnf_t *h, double *tau, uint64_t *cells, uint32_t *rebuilds) {
  if (!h)
    return fail(PRGPU_EINVAL, "nnf_band_info: NULL handle");
  if (tau)
    *tau = h->tau_ub;
  if (cells)
    *cells = h->band_cells;
  if (rebuilds)
    *rebuilds = h->band_rebuilds;
  return PRGPU_OK;
}

int prgpu_nnf_destroy(prgpu_nnf_t *h) {
  if (!h)
    return PRGPU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (DevBuf *b : {&h->d_succ_off, &h->d_succ_pos, &h->d_succ_pb, &h->d_vop, &h->d_stamp, &h->d_worklist, &h->d_known, &h->d_evaluated,
                    &h->d_hop, &h->d_path_pos, &h->d_path_slot, &h->d_path_cell, &h->d_out_path, &h->d_trace,
                    &h->d_lazy_out})
    b->release();
  for (DevBuf *b : {&h->d_ftab, &h->d_states, &h->d_pose, &h->d_lvl_start, &h->d_level, &h->d_pred_off, &h->d_pred_pos,
                    &h->d_pred_eid, &h->d_blocked, &h->d_pnn, &h->d_pref, &h->d_B, &h->d_path, &h->d_pn, &h->d_cost,
                    &h->d_band_lo, &h->d_band_w, &h->d_col_off, &h->d_meta, &h->d_edge, &h->d_Bc, &h->d_scan_tmp,
                    &h->d_rowmin, &h->d_done})
    b->release();
  h->timer.destroy();
  if (h->own_stream && h->stream)
    cudaStreamDestroy(h->stream);
  delete h;
  return PRGPU_OK;
}
} // extern "C"

namespace {
// buildNNGraph + (device form) buildTensorProductGraph.  w == NULL: host-callback form -- poses and
// node weights are the caller's (prgpu_nnf_set_weights), ref_pose12 is not used.
int nnf_create_common(prgpu_world_t *w, int device, int dof_in, const prgpu_nnf_params *params, int R,
                      const double *ref_pose12, const uint32_t *layer_counts, const double *ik_states,
                      prgpu_nnf_t **out) {
  if (!out)
    return fail(PRGPU_EINVAL, "nnf_create: out is NULL");
  *out = nullptr;
  if (!params || (w && !ref_pose12) || !layer_counts)
    return fail(PRGPU_EINVAL, "nnf_create: NULL argument");
  if (!w && (dof_in < 1 || dof_in > PRGPU_MAX_DIM))
    return fail(PRGPU_EINVAL, "nnf_create_graph: dof out of range");
  if (R < 2)
    return fail(PRGPU_EINVAL, "nnf_create: referencePath is empty!"); // setRefPath, NNFrechet.cpp:24-25
  if (params->num_nn <= 0 || params->num_nn > PRGPU_MAX_K)
    return fail(PRGPU_EINVAL, "nnf_create: numNN must be positive! (and <= PRGPU_MAX_K)");
  if (params->discretization <= 0)
    return fail(PRGPU_EINVAL, "nnf_create: discretization must be positive!");
  if (!(params->check_resolution > 0.0))
    return fail(PRGPU_EINVAL, "nnf_create: check_resolution must be positive");
  if (w)
    device = w->device;
  PRGPU_CUDA_TRY(cudaSetDevice(device));
  const int dof = w ? w->host.dof : dof_in, k = params->num_nn, D = params->discretization;
  // PRGPU_NNF_TRACE=1: the set-up's phases on stderr (host clock)
  const bool trace = getenv("PRGPU_NNF_TRACE") != nullptr;
  double t_prev = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  auto lapse = [&](const char *what) {
    if (!trace)
      return;
    const double t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    fprintf(stderr, "[nnf] setup %-12s %9.3f ms\n", what, t - t_prev);
    t_prev = t;
  };

  prgpu_nnf *h = new prgpu_nnf();
  h->world = w;
  h->device = device;
  h->host_form = w == nullptr;
  if (w) {
    h->stream = w->stream;
    h->sm_count = w->sm_count;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
      e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
      delete h;
      return fail(PRGPU_ECUDA, std::string("nnf_create_graph: ") + cudaGetErrorString(e));
    }
    h->own_stream = true;
  }
  cudaStream_t s = h->stream;
  h->params = *params;
  h->R = R;
  h->dof = dof;
  if (ref_pose12)
    h->ref_pose.assign(ref_pose12, ref_pose12 + (size_t)R * 12);
  uint64_t N = 0;
  for (int r = 0; r < R; ++r)
    N += layer_counts[r];
  if (N && !ik_states) {
    delete h;
    return fail(PRGPU_EINVAL, "nnf_create: ik_states is NULL");
  }
  h->nIk = (uint32_t)N;
  const uint32_t M0 = layer_counts[0], ML = layer_counts[R - 1];
  const uint32_t nInterior = (uint32_t)N - M0 - ML;
#define NNF_TRY(expr)                                                                              \
  do {                                                                                             \
    int _rc = (expr);                                                                              \
    if (_rc != PRGPU_OK) {                                                                         \
      prgpu_nnf_destroy(h);                                                                        \
      return _rc;                                                                                  \
    }                                                                                              \
  } while (0)
#define NNF_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      prgpu_nnf_destroy(h);                                                                        \
      return fail(PRGPU_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    }                                                                                              \
  } while (0)

  // ---- findKnnNodes for every node of layers 0..R-2 (NNFrechet.cpp:452-459) --------------------
  // candidates = strictly deeper layers = [suffix of the interior nodes] U [the last layer];
  // two device indices whose reported index is the NN-graph vertex id, merged per query under
  // (distance, vertex id) -- the order of util.cpp:49.
  const uint32_t nq = (uint32_t)N - ML; // queries: layer 0 then interior layers
  h->knn.assign((size_t)N * k, PRGPU_INVALID_INDEX);
  if (nq > 0) {
    DevBuf d_ik, d_q, d_lo, d_pi, d_pd, d_oi, d_od;
    auto cleanup = [&]() {
      for (DevBuf *b : {&d_ik, &d_q, &d_lo, &d_pi, &d_pd, &d_oi, &d_od})
        b->release();
    };
    prgpu_knn_t *A = nullptr, *B = nullptr;
    int rc = PRGPU_OK;
    do {
      if ((rc = d_ik.reserve(std::max<size_t>(N * dof * 8, 16))) != PRGPU_OK)
        break;
      cudaMemcpyAsync(d_ik.p, ik_states, N * dof * 8, cudaMemcpyHostToDevice, s);
      if ((rc = prgpu_knn_create(device, dof, std::max<uint32_t>(ML, 1), &A)) != PRGPU_OK)
        break;
      if ((rc = prgpu_knn_create(device, dof, std::max<uint32_t>(nInterior, 1), &B)) != PRGPU_OK)
        break;
      prgpu_knn_set_index_base(A, 2 + M0);
      prgpu_knn_set_index_base(B, 2 + M0 + ML);
      if ((rc = prgpu_knn_add_device(A, d_ik.as<double>() + (size_t)M0 * dof, ML, s)) != PRGPU_OK)
        break;
      if ((rc = prgpu_knn_add_device(B, d_ik.as<double>() + (size_t)(M0 + ML) * dof, nInterior, s)) != PRGPU_OK)
        break;
      // queries and their suffix bounds into the interior index
      std::vector<double> q((size_t)nq * dof);
      std::memcpy(q.data(), ik_states, (size_t)M0 * dof * 8);
      std::memcpy(q.data() + (size_t)M0 * dof, ik_states + (size_t)(M0 + ML) * dof, (size_t)nInterior * dof * 8);
      std::vector<uint32_t> lo(nq, 0);
      uint32_t cum = 0, qi = M0;
      for (int r = 1; r <= R - 2; ++r) {
        cum += layer_counts[r]; // first interior index of layer r+1
        for (uint32_t i = 0; i < layer_counts[r]; ++i)
          lo[qi++] = cum;
      }
      if ((rc = upload(d_q, q, s)) != PRGPU_OK || (rc = upload(d_lo, lo, s)) != PRGPU_OK)
        break;
      const size_t part = (size_t)nq * k;
      if ((rc = d_pi.reserve(2 * part * 4)) != PRGPU_OK || (rc = d_pd.reserve(2 * part * 8)) != PRGPU_OK ||
          (rc = d_oi.reserve(part * 4)) != PRGPU_OK || (rc = d_od.reserve(part * 8)) != PRGPU_OK)
        break;
      if ((rc = prgpu_knn_nearest_k_device(A, d_q.as<double>(), nq, k, nullptr, d_pi.as<uint32_t>(),
                                           d_pd.as<double>(), s)) != PRGPU_OK)
        break;
      if ((rc = prgpu_knn_nearest_k_device(B, d_q.as<double>(), nq, k, d_lo.as<uint32_t>(),
                                           d_pi.as<uint32_t>() + part, d_pd.as<double>() + part, s)) != PRGPU_OK)
        break;
      if ((rc = topk_merge(d_pi.as<uint32_t>(), d_pd.as<double>(), 2, nq, k, d_oi.as<uint32_t>(), d_od.as<double>(),
                           s)) != PRGPU_OK)
        break;
      std::vector<uint32_t> res(part);
      if (cudaMemcpyAsync(res.data(), d_oi.p, part * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
          cudaStreamSynchronize(s) != cudaSuccess) {
        rc = fail(PRGPU_ECUDA, "nnf_create: kNN download failed");
        break;
      }
      // rows back into IK order (layer 0 | last layer (none) | interior)
      std::memcpy(h->knn.data(), res.data(), (size_t)M0 * k * 4);
      std::memcpy(h->knn.data() + (size_t)(M0 + ML) * k, res.data() + (size_t)M0 * k, (size_t)nInterior * k * 4);
    } while (0);
    cudaStreamSynchronize(s);
    for (prgpu_knn_t *idx : {A, B}) {
      uint64_t fc = 0, fq = 0;
      if (idx && prgpu_knn_search_stats(idx, &fc, &fq) == PRGPU_OK) {
        h->knn_filter_calls += fc;
        h->knn_fallback_queries += fq;
      }
    }
    prgpu_knn_destroy(A);
    prgpu_knn_destroy(B);
    cleanup();
    NNF_TRY(rc);
  }

  lapse("knn");
  // ---- NN graph in creation order (:407-464) -----------------------------------------------------
  const uint32_t c0 = 0, c1 = 1;
  auto ikVertex = [](uint32_t i) { return 2 + i; };
  for (uint32_t i = 0; i < M0; ++i)
    h->edges.emplace_back(c0, ikVertex(i)); // :423
  for (uint32_t i = 0; i < ML; ++i)
    h->edges.emplace_back(ikVertex(M0 + i), c1); // :435
  std::vector<uint32_t> sub_src, sub_dst, sub_step, sub_layer;
  {
    const size_t n_sub = (size_t)nq * k * D;
    h->edges.reserve((size_t)M0 + ML + n_sub + (size_t)nq * k);
    sub_src.reserve(n_sub);
    sub_dst.reserve(n_sub);
    sub_step.reserve(n_sub);
    sub_layer.reserve(n_sub);
  }
  std::vector<uint32_t> layer_of_ik(N);
  {
    uint32_t i = 0;
    for (; i < M0; ++i)
      layer_of_ik[i] = 0;
    for (uint32_t j = 0; j < ML; ++j)
      layer_of_ik[i++] = (uint32_t)(R - 1);
    for (int r = 1; r <= R - 2; ++r)
      for (uint32_t j = 0; j < layer_counts[r]; ++j)
        layer_of_ik[i++] = (uint32_t)r;
  }
  uint32_t nextVertex = 2 + (uint32_t)N;
  auto expand = [&](uint32_t ik) { // addSubsampledEdge for every neighbour of one node (:461-462)
    for (int j = 0; j < k; ++j) {
      const uint32_t nbr = h->knn[(size_t)ik * k + j];
      if (nbr == PRGPU_INVALID_INDEX)
        break;
      uint32_t prev = ikVertex(ik);
      for (int i = 0; i < D; ++i) {
        const uint32_t a = nextVertex++;
        sub_src.push_back(ikVertex(ik));
        sub_dst.push_back(nbr);
        sub_step.push_back((uint32_t)i);
        sub_layer.push_back(layer_of_ik[ik]);
        h->edges.emplace_back(prev, a); // :388
        if (i == D - 1)
          h->edges.emplace_back(a, nbr); // :391
        prev = a;
      }
    }
  };
  for (uint32_t i = 0; i < M0; ++i) // refIndex 0
    expand(i);
  for (uint32_t i = M0 + ML; i < N; ++i) // refIndex 1..R-2 (vertex order == layer order)
    expand(i);
  const uint32_t V = nextVertex, E = (uint32_t)h->edges.size();
  h->V = V;
  h->E = E;

  lapse("expand");
  // ---- vertex data on the device: states, FK poses ----------------------------------------------
  NNF_TRY(h->d_states.reserve((size_t)V * dof * 8));
  NNF_TRY(h->d_pose.reserve((size_t)V * 12 * 8));
  NNF_CUDA(cudaMemsetAsync(h->d_states.p, 0, 2 * (size_t)dof * 8, s));
  if (N)
    NNF_CUDA(cudaMemcpyAsync(h->d_states.as<double>() + 2 * (size_t)dof, ik_states, N * dof * 8,
                             cudaMemcpyHostToDevice, s));
  if (w) {
    NNF_CUDA(cudaMemcpyAsync(h->d_pose.p, ref_pose12, 12 * 8, cudaMemcpyHostToDevice, s)); // c0 pose (:409)
    NNF_CUDA(cudaMemcpyAsync(h->d_pose.as<double>() + 12, ref_pose12 + 12 * (size_t)(R - 1), 12 * 8,
                             cudaMemcpyHostToDevice, s)); // c1 pose (:414)
    NNF_TRY(fk_launch(w->host, w->dev, h->d_states.as<double>() + 2 * (size_t)dof, N,
                      h->d_pose.as<double>() + 2 * 12, s)); // sampleIKNodes :295
  }
  {
    DevBuf d_src, d_dst, d_step;
    int rc = upload(d_src, sub_src, s);
    if (rc == PRGPU_OK)
      rc = upload(d_dst, sub_dst, s);
    if (rc == PRGPU_OK)
      rc = upload(d_step, sub_step, s);
    if (rc == PRGPU_OK)
      rc = nnf_subsample_launch(w ? w->dev : nullptr, dof, (uint32_t)sub_src.size(), 2 + (uint32_t)N, d_src.as<uint32_t>(),
                                d_dst.as<uint32_t>(), d_step.as<uint32_t>(), D, h->d_states.as<double>(),
                                h->d_pose.as<double>(), s);
    cudaStreamSynchronize(s);
    d_src.release();
    d_dst.release();
    d_step.release();
    NNF_TRY(rc);
  }
  // (host copies of the vertex states / poses -- 185 MB at the benchmark size -- are fetched on first use:
  //  prgpu_nnf_get_vertices and the host LazySP loop; the device-resident loop never needs them)

  lapse("subsample");
  // ---- predecessors (by vertex, creation order), topological levels, level-sorted order ---------
  h->pred_off_v.assign(V + 1, 0);
  for (auto &e : h->edges)
    h->pred_off_v[e.second + 1]++;
  for (uint32_t v = 0; v < V; ++v)
    h->pred_off_v[v + 1] += h->pred_off_v[v];
  h->pred_v.resize(E);
  {
    std::vector<uint32_t> cur(h->pred_off_v.begin(), h->pred_off_v.end() - 1);
    for (uint32_t e = 0; e < E; ++e)
      h->pred_v[cur[h->edges[e].second]++] = std::make_pair(h->edges[e].first, e);
  }
  // level: c0 = 0; IK node of layer r = 1 + r(D+1); i-th sub-sampled vertex of an edge leaving
  // layer r = 1 + r(D+1) + 1 + i; c1 = 2 + (R-1)(D+1)
  const uint32_t Lv = 3 + (uint32_t)(R - 1) * (uint32_t)(D + 1);
  h->Lv = Lv;
  std::vector<uint32_t> level(V);
  level[c0] = 0;
  level[c1] = Lv - 1;
  for (uint32_t i = 0; i < N; ++i)
    level[ikVertex(i)] = 1 + layer_of_ik[i] * (uint32_t)(D + 1);
  for (size_t i = 0; i < sub_src.size(); ++i)
    level[2 + N + i] = 1 + sub_layer[i] * (uint32_t)(D + 1) + 1 + sub_step[i];
  std::vector<uint32_t> lvl_start(Lv + 1, 0);
  for (uint32_t v = 0; v < V; ++v)
    lvl_start[level[v] + 1]++;
  for (uint32_t l = 0; l < Lv; ++l)
    lvl_start[l + 1] += lvl_start[l];
  h->pos_of_vertex.resize(V);
  h->vertex_of_pos.resize(V);
  {
    std::vector<uint32_t> cur(lvl_start.begin(), lvl_start.end() - 1);
    for (uint32_t v = 0; v < V; ++v) { // stable: vertex id order inside a level
      const uint32_t p = cur[level[v]]++;
      h->pos_of_vertex[v] = p;
      h->vertex_of_pos[p] = v;
    }
  }
  std::vector<uint32_t> level_s(V), pred_off_s(V + 1, 0), pred_pos(E), pred_eid(E);
  for (uint32_t p = 0; p < V; ++p) {
    const uint32_t v = h->vertex_of_pos[p];
    level_s[p] = level[v];
    pred_off_s[p + 1] = pred_off_s[p] + (h->pred_off_v[v + 1] - h->pred_off_v[v]);
  }
  for (uint32_t p = 0; p < V; ++p) {
    const uint32_t v = h->vertex_of_pos[p];
    uint32_t o = pred_off_s[p];
    for (uint32_t e = h->pred_off_v[v]; e < h->pred_off_v[v + 1]; ++e, ++o) {
      pred_pos[o] = h->pos_of_vertex[h->pred_v[e].first];
      pred_eid[o] = h->pred_v[e].second;
    }
  }
  h->level_of_vertex = level;
  h->lvl_start_host = lvl_start;
  h->slot_of_eid.assign(E, 0);
  for (uint32_t o = 0; o < E; ++o)
    h->slot_of_eid[pred_eid[o]] = o;
  lapse("csr");
  NNF_TRY(upload(h->d_lvl_start, lvl_start, s));
  NNF_TRY(upload(h->d_level, level_s, s));
  NNF_TRY(upload(h->d_pred_off, pred_off_s, s));
  NNF_TRY(upload(h->d_pred_pos, pred_pos, s));
  NNF_TRY(upload(h->d_pred_eid, pred_eid, s));
  h->evaluated.assign(E, 0);
  h->blocked.assign(E, 0);
  h->known.assign(E, -1);
  NNF_TRY(upload(h->d_blocked, h->blocked, s));
  NNF_TRY(upload(h->d_vop, h->vertex_of_pos, s));
  if (w) {
    NNF_TRY(h->d_pnn.reserve((size_t)V * 3 * 8));
    NNF_TRY(nnf_translations_launch(h->d_pose.as<double>(), h->d_vop.as<uint32_t>(), V, h->d_pnn.as<double>(), s));
    // successors by sorted position + the work arrays of the device-resident LazySP loop
    std::vector<uint32_t> succ_off(V + 1, 0), succ_pos(E), succ_pb(E);
    for (uint32_t o = 0; o < E; ++o)
      succ_off[pred_pos[o] + 1]++;
    for (uint32_t p = 0; p < V; ++p)
      succ_off[p + 1] += succ_off[p];
    {
      std::vector<uint32_t> cur(succ_off.begin(), succ_off.end() - 1);
      for (uint32_t p = 0; p < V; ++p)
        for (uint32_t o = pred_off_s[p]; o < pred_off_s[p + 1]; ++o)
          succ_pos[cur[pred_pos[o]]++] = p;
      for (uint32_t e = 0; e < E; ++e)
        succ_pb[e] = pred_off_s[succ_pos[e]];
    }
    NNF_TRY(upload(h->d_succ_off, succ_off, s));
    NNF_TRY(upload(h->d_succ_pos, succ_pos, s));
    NNF_TRY(upload(h->d_succ_pb, succ_pb, s));
    NNF_TRY(h->d_stamp.reserve((size_t)V * 4));
    NNF_CUDA(cudaMemsetAsync(h->d_stamp.p, 0, (size_t)V * 4, s));
    NNF_TRY(h->d_worklist.reserve((size_t)V * 4));
    NNF_TRY(h->d_known.reserve(std::max<size_t>(E, 16)));
    NNF_TRY(h->d_evaluated.reserve(std::max<size_t>(E, 16)));
    const uint32_t cap_hops = Lv + 8;
    NNF_TRY(h->d_path_pos.reserve((size_t)cap_hops * 4));
    NNF_TRY(h->d_path_slot.reserve((size_t)cap_hops * 4));
    NNF_TRY(h->d_path_cell.reserve((size_t)cap_hops * 8));
    NNF_TRY(h->d_out_path.reserve((size_t)cap_hops * 4));
    NNF_TRY(h->d_lazy_out.reserve(sizeof(NnfLazyOut)));
    NNF_CUDA(cudaStreamSynchronize(s));
  }
  if (w) {
    std::vector<double> pref((size_t)R * 3);
    for (int r = 0; r < R; ++r) {
      pref[3 * r + 0] = ref_pose12[12 * (size_t)r + 3];
      pref[3 * r + 1] = ref_pose12[12 * (size_t)r + 7];
      pref[3 * r + 2] = ref_pose12[12 * (size_t)r + 11];
    }
    NNF_TRY(upload(h->d_pref, pref, s));
  }
  NNF_TRY(h->d_band_lo.reserve((size_t)V * 4));
  NNF_TRY(h->d_band_w.reserve((size_t)V * 8));
  NNF_TRY(h->d_col_off.reserve((size_t)V * 8));
  NNF_TRY(h->d_meta.reserve((size_t)V * sizeof(NnfVertexMeta)));
  NNF_TRY(h->d_edge.reserve(std::max<size_t>((size_t)E * sizeof(NnfEdgeRec), 16)));
  NNF_TRY(h->d_rowmin.reserve((size_t)R * 8));
  NNF_TRY(h->d_done.reserve((size_t)V * 4));
  NNF_CUDA(cudaMemsetAsync(h->d_done.p, 0, (size_t)V * 4, s));
  NNF_TRY(h->d_path.reserve(((size_t)R + Lv + 8) * 2 * 4));
  NNF_TRY(h->d_pn.reserve(16));
  NNF_TRY(h->d_cost.reserve(16));
  NNF_TRY(h->timer.enable(true));
  NNF_CUDA(cudaStreamSynchronize(s));
  lapse("uploads");

  NnfGraphDev &g = h->g;
  g.R = R;
  g.V = V;
  g.Lv = Lv;
  g.lvl_start = h->d_lvl_start.as<uint32_t>();
  g.level = h->d_level.as<uint32_t>();
  g.pred_off = h->d_pred_off.as<uint32_t>();
  g.pred_pos = h->d_pred_pos.as<uint32_t>();
  g.pred_eid = h->d_pred_eid.as<uint32_t>();
  g.blocked = h->d_blocked.as<uint8_t>();
  g.pnn = w ? h->d_pnn.as<double>() : nullptr;
  g.pref = w ? h->d_pref.as<double>() : nullptr;
  g.ftab = nullptr; // host-callback form: set by prgpu_nnf_set_weights
  g.B = nullptr; // allocated on first use of the full-table mode
  g.start_pos = h->pos_of_vertex[c0];
  g.goal_pos = h->pos_of_vertex[c1];
  NnfBandDev &bg = h->bg;
  bg.R = R;
  bg.V = V;
  bg.Lv = Lv;
  bg.lvl_start = g.lvl_start;
  bg.meta = h->d_meta.as<NnfVertexMeta>();
  bg.edge = h->d_edge.as<NnfEdgeRec>();
  bg.pred_pos = g.pred_pos;
  bg.pnn = g.pnn;
  bg.pref = g.pref;
  bg.ftab = nullptr;
  bg.Bc = nullptr;
  bg.hop = nullptr;
  bg.start_pos = g.start_pos;
  bg.goal_pos = g.goal_pos;
  *out = h;
  return PRGPU_OK;
#undef NNF_TRY
#undef NNF_CUDA
}
} // namespace

extern "C" {
int prgpu_nnf_create(prgpu_world_t *w, const prgpu_nnf_params *params, int R, const double *ref_pose12,
                     const uint32_t *layer_counts, const double *ik_states, prgpu_nnf_t **out) {
  if (!w) {
    if (out)
      *out = nullptr;
    return fail(PRGPU_EINVAL, "nnf_create: NULL argument");
  }
  return nnf_create_common(w, 0, 0, params, R, ref_pose12, layer_counts, ik_states, out);
}

int prgpu_nnf_create_graph(int device, int dof, const prgpu_nnf_params *params, int R, const uint32_t *layer_counts,
                           const double *ik_states, prgpu_nnf_t **out) {
  return nnf_create_common(nullptr, device, dof, params, R, nullptr, layer_counts, ik_states, out);
}

int prgpu_nnf_set_weights(prgpu_nnf_t *h, const double *weights) {
  if (!h || !weights)
    return fail(PRGPU_EINVAL, "nnf_set_weights: NULL argument");
  if (!h->host_form)
    return fail(PRGPU_EINVAL, "nnf_set_weights: the handle was created with a device world (prgpu_nnf_create)");
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  const size_t R = (size_t)h->R, V = h->V;
  for (size_t i = 0; i < R * V; ++i)
    if (!(weights[i] >= 0.0))
      return fail(PRGPU_EINVAL, "nnf_set_weights: node weights must be non-negative numbers");
  h->ftab_host.assign(weights, weights + R * V);
  std::vector<double> sorted(R * V);
  for (size_t p = 0; p < V; ++p)
    std::memcpy(sorted.data() + p * R, weights + (size_t)h->vertex_of_pos[p] * R, R * 8);
  int rc = upload(h->d_ftab, sorted, h->stream);
  if (rc != PRGPU_OK)
    return rc;
  PRGPU_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->g.ftab = h->bg.ftab = h->d_ftab.as<double>();
  h->weights_set = true;
  h->band_valid = false; // a new table: the band and every cost are stale
  h->tau_ub = 0.0;
  h->dirty_level = 0;
  return PRGPU_OK;
}

int prgpu_nnf_set_edge_evaluator(prgpu_nnf_t *h, prgpu_nnf_edge_eval_fn fn, void *user) {
  if (!h)
    return fail(PRGPU_EINVAL, "nnf_set_edge_evaluator: NULL handle");
  h->eval_fn = fn;
  h->eval_user = user;
  return PRGPU_OK;
}

int prgpu_nnf_graph_info(prgpu_nnf_t *h, uint32_t *n_vertices, uint32_t *n_edges, uint32_t *n_ik) {
  if (!h)
    return fail(PRGPU_EINVAL, "nnf_graph_info: NULL handle");
  if (n_vertices)
    *n_vertices = h->V;
  if (n_edges)
    *n_edges = h->E;
  if (n_ik)
    *n_ik = h->nIk;
  return PRGPU_OK;
}

int prgpu_nnf_get_knn(prgpu_nnf_t *h, uint32_t *knn) {
  if (!h || !knn)
    return fail(PRGPU_EINVAL, "nnf_get_knn: NULL argument");
  std::memcpy(knn, h->knn.data(), h->knn.size() * 4);
  return PRGPU_OK;
}

int prgpu_nnf_get_vertices(prgpu_nnf_t *h, double *states, double *pose12) {
  if (!h)
    return fail(PRGPU_EINVAL, "nnf_get_vertices: NULL handle");
  const int rcf = nnf_fetch_vertices(h);
  if (rcf != PRGPU_OK)
    return rcf;
  if (states)
    std::memcpy(states, h->states.data(), h->states.size() * 8);
  if (pose12) {
    if (h->host_form)
      return fail(PRGPU_EINVAL, "nnf_get_vertices: poses belong to the caller's FK in the host-callback form");
    std::memcpy(pose12, h->pose.data(), h->pose.size() * 8);
  }
  return PRGPU_OK;
}

int prgpu_nnf_knn_stats(prgpu_nnf_t *h, uint64_t *filter_calls, uint64_t *fallback_queries) {
  if (!h)
    return fail(PRGPU_EINVAL, "nnf_knn_stats: NULL handle");
  if (filter_calls)
    *filter_calls = h->knn_filter_calls;
  if (fallback_queries)
    *fallback_queries = h->knn_fallback_queries;
  return PRGPU_OK;
}

int prgpu_nnf_get_edges(prgpu_nnf_t *h, uint32_t *edges, uint8_t *flags) {
  if (!h)
    return fail(PRGPU_EINVAL, "nnf_get_edges: NULL handle");
  if (flags) {
    PRGPU_CUDA_TRY(cudaSetDevice(h->device));
    int rc = nnf_state_to_host(h);
    if (rc != PRGPU_OK)
      return rc;
  }
  for (uint32_t e = 0; e < h->E; ++e) {
    if (edges) {
      edges[2 * (size_t)e] = h->edges[e].first;
      edges[2 * (size_t)e + 1] = h->edges[e].second;
    }
    if (flags)
      flags[e] = (uint8_t)((h->evaluated[e] ? 1 : 0) | (h->blocked[e] ? 2 : 0));
  }
  return PRGPU_OK;
}

int prgpu_nnf_loop_stats(prgpu_nnf_t *h, uint64_t *launches, uint64_t *repaired_vertices, uint64_t *phase_cycles) {
  if (!h)
    return fail(PRGPU_EINVAL, "nnf_loop_stats: NULL handle");
  if (phase_cycles) {
    for (int i = 0; i < 5; ++i)
      phase_cycles[i] = h->lazy_cycles[i];
    phase_cycles[5] = h->lazy_bailouts;
  }
  if (launches)
    *launches = h->lazy_launches;
  if (repaired_vertices)
    *repaired_vertices = h->repaired_vertices;
  return PRGPU_OK;
}

int prgpu_nnf_last_dp_ms(prgpu_nnf_t *h, float *ms) {
  if (!h || !ms)
    return fail(PRGPU_EINVAL, "nnf_last_dp_ms: NULL argument");
  return h->timer.read(ms);
}

int prgpu_nnf_solve(prgpu_nnf_t *h, uint32_t max_lazy_iters, prgpu_nnf_result *res, uint32_t *path_vertices,
                    uint32_t cap_path, double *goal_cost_trace) {
  if (!h || !res)
    return fail(PRGPU_EINVAL, "nnf_solve: NULL argument");
  if (h->host_form && !h->weights_set)
    return fail(PRGPU_EINVAL, "nnf_solve: node weights not set (prgpu_nnf_set_weights)");
  if (h->host_form && !h->eval_fn)
    return fail(PRGPU_EINVAL, "nnf_solve: edge evaluator not set (prgpu_nnf_set_edge_evaluator)");
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int R = h->R, dof = h->dof;
  const uint32_t V = h->V, c0 = 0, c1 = 1;
  std::memset(res, 0, sizeof(*res));
  res->n_vertices = V;
  res->n_edges = h->E;
  res->n_ik = h->nIk;
  const uint32_t cap_pairs = (uint32_t)R + h->Lv + 8;
  std::vector<uint32_t> tp(2 * (size_t)cap_pairs);
  std::vector<uint32_t> finalPath;
  bool solutionFound = false;
  uint32_t iters = 0;
  static const bool host_loop_forced = getenv("PRGPU_NNF_HOST_LOOP") != nullptr;
  const bool device_loop = h->mode == 0 && !h->eval_fn && h->world && !host_loop_forced && h->V < NNF_HOP_MAX_V &&
                           h->d_stamp.p != nullptr;
  if (device_loop) {
    // ---- the whole loop on the device: full search only after a band (re)build, local repair otherwise ----
    int rc = nnf_state_to_device(h);
    if (rc != PRGPU_OK)
      return rc;
    h->host_state_current = false;
    if (goal_cost_trace && (rc = h->d_trace.reserve(std::max<size_t>((size_t)max_lazy_iters * 8, 16))) != PRGPU_OK)
      return rc;
    // PRGPU_NNF_TRACE=1: one line per launch on stderr (host clock; every step below ends in a stream sync)
    static const bool trace = getenv("PRGPU_NNF_TRACE") != nullptr;
    auto now_ms = []() {
      return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    double t_prev = now_ms();
    auto lapse = [&](const char *what, uint32_t a, uint32_t b) {
      if (!trace)
        return;
      cudaStreamSynchronize(s);
      const double t = now_ms();
      fprintf(stderr, "[nnf] %-10s %9.3f ms  %u %u\n", what, t - t_prev, a, b);
      t_prev = t;
    };
    while (true) {
      if (!h->band_valid) {
        if (h->tau_ub == 0.0 && (rc = nnf_initial_tau(h, s)) != PRGPU_OK)
          return rc;
        if ((rc = nnf_build_band(h, s)) != PRGPU_OK)
          return rc;
        lapse("band", 0, 0);
      }
      if (!h->bg.hop)
        return fail(PRGPU_ELIMIT, "nnf_solve: band too large for the hop words");
      if (h->dirty_level < h->Lv) {
        rc = nnf_flow_dp_launch(h->bg, h->lvl_start_host[h->dirty_level], ++h->epoch, h->d_done.as<uint32_t>(),
                                h->sm_count, s, &h->timer);
        if (rc != PRGPU_OK)
          return rc;
        lapse("sweep", h->dirty_level, h->Lv);
        h->dirty_level = h->Lv;
      }
      NnfLazyDev z;
      z.succ_off = h->d_succ_off.as<uint32_t>();
      z.succ_pos = h->d_succ_pos.as<uint32_t>();
      z.succ_pb = h->d_succ_pb.as<uint32_t>();
      z.level = h->g.level;
      z.vertex_of_pos = h->d_vop.as<uint32_t>();
      z.pred_eid = h->g.pred_eid;
      z.edge_rw = h->d_edge.as<NnfEdgeRec>();
      z.stamp = h->d_stamp.as<uint32_t>();
      z.worklist = h->d_worklist.as<uint32_t>();
      z.states = h->d_states.as<double>();
      z.known = h->d_known.as<int8_t>();
      z.evaluated = h->d_evaluated.as<uint8_t>();
      z.blocked = h->d_blocked.as<uint8_t>();
      z.path_pos = h->d_path_pos.as<uint32_t>();
      z.path_slot = h->d_path_slot.as<uint32_t>();
      z.path_cell = h->d_path_cell.as<unsigned long long>();
      z.cap_hops = h->Lv + 8;
      z.out_path = h->d_out_path.as<uint32_t>();
      z.trace = goal_cost_trace ? h->d_trace.as<double>() + iters : nullptr;
      z.out = h->d_lazy_out.as<NnfLazyOut>();
      z.max_iters = max_lazy_iters - iters;
      z.epoch = h->lazy_epoch;
      static const uint32_t wave_budget = getenv("PRGPU_NNF_WAVE_BUDGET") ? (uint32_t)atoi(getenv("PRGPU_NNF_WAVE_BUDGET")) : 2048u;
      z.wave_budget = std::min<uint32_t>(wave_budget, 2048u); // (the kernel's level list holds 2048 columns)
      z.tau_ub = h->tau_ub;
      z.check_resolution = h->params.check_resolution;
      z.band_full = h->band_full ? 1 : 0;
      z.dof = dof;
      if ((rc = nnf_lazysp_launch(h->world->host, h->world->dev, h->bg, z, s)) != PRGPU_OK)
        return rc;
      ++h->lazy_launches;
      NnfLazyOut o;
      PRGPU_CUDA_TRY(cudaMemcpyAsync(&o, h->d_lazy_out.p, sizeof(o), cudaMemcpyDeviceToHost, s));
      PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
      lapse("lazy", o.iters, o.status);
      iters += o.iters;
      h->total_iters += o.iters;
      h->n_evaluated += o.n_evaluated;
      h->n_collisions += o.n_collisions;
      h->lazy_epoch = o.epoch;
      h->repaired_vertices += o.repaired_vertices;
      for (int i = 0; i < 5; ++i)
        h->lazy_cycles[i] += o.cycles[i];
      if (o.iters)
        res->goal_cost = o.goal_cost;
      if (o.final_error >= 0.0) // a path was extracted in this launch
        res->final_error = o.final_error;
      if (o.status == NNF_LAZY_NEED_FLOW) { // a wide repair wave: whole-GPU sweep from its level, then back to the loop
        h->dirty_level = std::min(h->dirty_level, o.pad);
        ++h->lazy_bailouts;
        continue;
      }
      if (o.status == NNF_LAZY_NEED_REBUILD) {
        static const double headroom = getenv("PRGPU_NNF_TAU_HEADROOM") ? atof(getenv("PRGPU_NNF_TAU_HEADROOM")) : 1.15;
        h->tau_ub = o.goal_cost < 1e300 ? o.goal_cost * headroom : h->tau_ub * 2.0;
        if ((rc = nnf_build_band(h, s)) != PRGPU_OK)
          return rc;
        lapse("reband", 0, 0);
        continue;
      }
      if (o.status == NNF_LAZY_ERROR)
        return fail(PRGPU_ECUDA, "nnf_solve: inconsistent bottleneck table while following the hops");
      if (o.status == NNF_LAZY_SOLVED) {
        solutionFound = true;
        finalPath.resize(o.path_len);
        if (o.path_len)
          PRGPU_CUDA_TRY(cudaMemcpyAsync(finalPath.data(), h->d_out_path.p, (size_t)o.path_len * 4, cudaMemcpyDeviceToHost, s));
      }
      break; // solved, no path, or the iteration budget of this call is spent
    }
    if (goal_cost_trace && iters)
      PRGPU_CUDA_TRY(cudaMemcpyAsync(goal_cost_trace, h->d_trace.p, (size_t)iters * 8, cudaMemcpyDeviceToHost, s));
    PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
  } else {
    int rc0 = nnf_state_to_host(h);
    if (rc0 == PRGPU_OK)
      rc0 = nnf_fetch_vertices(h); // evaluateEdge batches and mFinalError of the host loop read the host copies
    if (rc0 != PRGPU_OK)
      return rc0;
    h->dev_state_current = false;
  }
  while (!device_loop && iters < max_lazy_iters) { // NNFrechet.cpp:106
    int rc;
    uint32_t np = 0;
    double goalCost = 0.0;
    if (h->mode == 1) { // full-table wavefront
      if (!h->g.B) {
        if ((rc = h->d_B.reserve((size_t)R * V * 8)) != PRGPU_OK)
          return rc;
        h->g.B = h->d_B.as<double>();
      }
      rc = nnf_dp_launch(h->g, h->sm_count, s, &h->timer); // computeShortestPath :109-110
      if (rc != PRGPU_OK)
        return rc;
      rc = nnf_backtrack_launch(h->g, h->d_path.as<uint32_t>(), cap_pairs, h->d_pn.as<uint32_t>(),
                                h->d_cost.as<double>(), s);
      if (rc != PRGPU_OK)
        return rc;
      PRGPU_CUDA_TRY(cudaMemcpyAsync(&np, h->d_pn.p, 4, cudaMemcpyDeviceToHost, s));
      PRGPU_CUDA_TRY(cudaMemcpyAsync(&goalCost, h->d_cost.p, 8, cudaMemcpyDeviceToHost, s));
      PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
    } else { // banded column-wise DP; widen the band until it provably contains the optimum
      if (!h->band_valid) {
        if (h->tau_ub == 0.0 && (rc = nnf_initial_tau(h, s)) != PRGPU_OK)
          return rc;
        if ((rc = nnf_build_band(h, s)) != PRGPU_OK)
          return rc;
      }
      while (true) {
        if (h->mode == 2)
          rc = nnf_band_dp_launch(h->bg, h->dirty_level, s, &h->timer);
        else
          rc = nnf_flow_dp_launch(h->bg, h->dirty_level >= h->Lv ? h->V : h->lvl_start_host[h->dirty_level],
                                  ++h->epoch, h->d_done.as<uint32_t>(), h->sm_count, s, &h->timer);
        if (rc != PRGPU_OK)
          return rc;
        h->dirty_level = h->Lv; // table consistent
        rc = nnf_band_backtrack_launch(h->bg, h->d_path.as<uint32_t>(), cap_pairs, h->d_pn.as<uint32_t>(),
                                       h->d_cost.as<double>(), s);
        if (rc != PRGPU_OK)
          return rc;
        PRGPU_CUDA_TRY(cudaMemcpyAsync(&np, h->d_pn.p, 4, cudaMemcpyDeviceToHost, s));
        PRGPU_CUDA_TRY(cudaMemcpyAsync(&goalCost, h->d_cost.p, 8, cudaMemcpyDeviceToHost, s));
        PRGPU_CUDA_TRY(cudaStreamSynchronize(s));
        if (goalCost <= h->tau_ub || h->band_full)
          break; // exact: the optimum lies inside the band (or the band is the whole table)
        // a finite cost found inside the band is an upper bound of the optimum: use it directly
        // (with some headroom: the goal cost creeps up with every knocked-out edge, and a rebuild costs
        // about ten searches)
        static const double headroom = getenv("PRGPU_NNF_TAU_HEADROOM") ? atof(getenv("PRGPU_NNF_TAU_HEADROOM")) : 1.15;
        h->tau_ub = goalCost < 1e300 ? goalCost * headroom : h->tau_ub * 2.0;
        if ((rc = nnf_build_band(h, s)) != PRGPU_OK)
          return rc;
      }
    }
    if (goal_cost_trace)
      goal_cost_trace[iters] = goalCost;
    ++iters;
    ++h->total_iters;
    res->goal_cost = goalCost;
    if (np == 0) // search found no path (:112-113)
      break;
    if (np == 0xFFFFFFFFu || np > cap_pairs)
      return fail(PRGPU_ECUDA, "nnf_solve: inconsistent bottleneck table while backtracking");
    PRGPU_CUDA_TRY(cudaMemcpyAsync(tp.data(), h->d_path.p, (size_t)np * 8, cudaMemcpyDeviceToHost, s));
    PRGPU_CUDA_TRY(cudaStreamSynchronize(s));

    // extractNNPath (:615-645); the device path runs goal -> start
    std::vector<uint32_t> nnPath;
    double bottleneck = 0;
    for (uint32_t i = np; i-- > 0;) {
      const uint32_t r = tp[2 * i], v = h->vertex_of_pos[tp[2 * i + 1]];
      if (v == c0 || v == c1)
        continue;
      if (nnPath.empty() || nnPath.back() != v)
        nnPath.push_back(v);
      const double c = h->host_form
                           ? h->ftab_host[(size_t)v * R + r]
                           : host_task_distance(h->ref_pose.data() + 12 * (size_t)r, h->pose.data() + 12 * (size_t)v);
      if (c > bottleneck)
        bottleneck = c;
    }
    res->final_error = bottleneck;

    // edges of the path; evaluate the not-yet-known ones in one batch (evaluateEdge is a pure
    // function of the edge, so knowing a verdict early is invisible to the LazySP bookkeeping)
    std::vector<uint32_t> eids;
    for (size_t i = 0; i + 1 < nnPath.size(); ++i) {
      const uint32_t u = nnPath[i], v = nnPath[i + 1];
      uint32_t eid = PRGPU_INVALID_INDEX;
      for (uint32_t e = h->pred_off_v[v]; e < h->pred_off_v[v + 1]; ++e)
        if (h->pred_v[e].first == u)
          eid = h->pred_v[e].second;
      if (eid == PRGPU_INVALID_INDEX)
        return fail(PRGPU_ECUDA, "nnf_solve: path uses a non-existent NN edge");
      eids.push_back(eid);
    }
    std::vector<uint32_t> todo;
    for (uint32_t eid : eids)
      if (h->known[eid] < 0 && std::find(todo.begin(), todo.end(), eid) == todo.end())
        todo.push_back(eid);
    if (!todo.empty()) {
      std::vector<double> from(todo.size() * dof), to(todo.size() * dof);
      for (size_t i = 0; i < todo.size(); ++i) {
        std::memcpy(from.data() + i * dof, h->states.data() + (size_t)h->edges[todo[i]].first * dof, dof * 8);
        std::memcpy(to.data() + i * dof, h->states.data() + (size_t)h->edges[todo[i]].second * dof, dof * 8);
      }
      std::vector<uint8_t> ok(todo.size());
      if (h->eval_fn) { // the caller's validity checker (evaluateEdge over si_->isValid, NNFrechet.cpp:655-683)
        if (h->eval_fn(h->eval_user, from.data(), to.data(), (uint32_t)todo.size(), ok.data()) != 0)
          return fail(PRGPU_EINVAL, "nnf_solve: the edge evaluator reported an error");
      } else {
        rc = prgpu_check_motion_batch(h->world, from.data(), to.data(), todo.size(), h->params.check_resolution,
                                      PRGPU_MODE_NNF_ALPHA, 0, ok.data());
        if (rc != PRGPU_OK)
          return rc;
      }
      for (size_t i = 0; i < todo.size(); ++i)
        h->known[todo[i]] = ok[i] ? 1 : 0;
    }
    bool collisionFree = true;
    finalPath.clear();
    for (size_t i = 0; i < eids.size(); ++i) { // :120-145
      const uint32_t eid = eids[i];
      if (!h->evaluated[eid]) {
        h->evaluated[eid] = 1;
        ++h->n_evaluated;
        if (h->known[eid] == 0) { // in collision: markEdgeInCollision (:131-136, :685-702)
          h->blocked[eid] = 1;
          ++h->n_collisions;
          nnf_block_edge_kernel<<<1, 1, 0, s>>>(h->d_edge.as<NnfEdgeRec>(), h->slot_of_eid[eid],
                                                h->d_blocked.as<uint8_t>(), eid); PRGPU_COUNT_LAUNCH(1);
          PRGPU_CUDA_TRY(cudaGetLastError());
          h->dirty_level = std::min(h->dirty_level, h->level_of_vertex[h->edges[eid].second]);
          collisionFree = false;
          break;
        }
      }
      if (finalPath.empty()) {
        finalPath.push_back(nnPath[i]);
        finalPath.push_back(nnPath[i + 1]);
      } else
        finalPath.push_back(nnPath[i + 1]);
    }
    if (collisionFree) { // :148-151
      solutionFound = true;
      if (finalPath.empty() && nnPath.size() == 1)
        finalPath = nnPath;
      break;
    }
  }
  res->lazy_iters = iters;
  res->solved = solutionFound ? 1 : 0;
  res->n_evaluated = h->n_evaluated;
  res->n_collisions = h->n_collisions;
  if (solutionFound) {
    res->path_len = (uint32_t)finalPath.size();
    for (size_t i = 0; i < finalPath.size(); ++i) {
      if (i >= cap_path || !path_vertices) {
        res->overflow = 1;
        break;
      }
      path_vertices[i] = finalPath[i];
    }
  }
  return PRGPU_OK;
}

} // extern "C"

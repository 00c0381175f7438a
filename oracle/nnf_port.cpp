// ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle.h.
// Flat-array CPU restatement ("port") of the NNF planner, function by function:
//   linDoubleSpace / linIntSpace     <- frechet/src/util.cpp:3-21
//   sampleNNGraphNodes (IK budget)   <- frechet/src/NNFrechet.cpp:310-343
//   buildNNGraph                     <- frechet/src/NNFrechet.cpp:398-465
//   findKnnNodes                     <- frechet/src/util.cpp:34-58   (full sort of (distance, vertex))
//   addSubsampledEdge                <- frechet/src/NNFrechet.cpp:345-396
//   addTensorProductNodes/Edge       <- frechet/src/NNFrechet.cpp:467-547  (f = task distance; w = max(f[u], f[v]))
//   connectTensorProductNodes        <- frechet/src/NNFrechet.cpp:549-599  (steps A: ref, B: NN, C: both)
//   LPAStar::computeShortestPath     <- frechet/src/LPAStar.cpp:125-159    (min-bottleneck; DAG only, LPAStar.h:14-15)
//   extractNNPath / mFinalError      <- frechet/src/NNFrechet.cpp:615-645
//   evaluateEdge                     <- frechet/src/NNFrechet.cpp:655-683
//   markEdgeInCollision + solve loop <- frechet/src/NNFrechet.cpp:685-702, :95-178
// The bottleneck search is restated as a topological dynamic program over the tensor-product
// DAG: B[v] = min over predecessors u of max(B[u], w(u,v)), B[start] = 0, which is the fixed
// point LPA* converges to.  Among equally good predecessors the literal
// LPA* keeps whichever its heap and hash containers reach first (implementation-defined);
// here the rule is fixed: first minimum in the order [(r-1,n)], then for each NN predecessor p
// in edge-creation order [(r,p), (r-1,p)].  The bottleneck cost is unaffected by that choice.
// Pinned against the reference itself: frechet/src/*.cpp compiled unmodified into
// oracle/_ref/libref_frechet.so (oracle/Makefile, oracle/ref_nnf_driver.cpp) and the golden vectors
// generated from it (tests/golden/make_golden_nnf.py, tests/test_oracle_nnf_ref.py), plus the
// independent checks of tests/test_oracle_nnf.py.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <random>
#include <thread>
#include <vector>
#include "oracle.h"
#include "synth_world.h"

namespace {
const double INF = DBL_MAX; // LPAStar.h:57 mInfVal

double taskDistance(const double *a12, const double *b12) {
  const double dx = a12[3] - b12[3], dy = a12[7] - b12[7], dz = a12[11] - b12[11];
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}
} // namespace

extern "C" void orc_lin_int_space(double mn, double mx, uint64_t n, int32_t *out) {
  double delta = (mx - mn) / double(n - 1); // util.cpp:5
  for (uint64_t i = 0; i < n; i++) {
    double cur = mn + i * delta; // util.cpp:7
    out[i] = (int)cur;           // util.cpp:17
  }
}

extern "C" void orc_nnf_layer_counts(int seed, int R, int W, int M, uint32_t *counts) {
  std::default_random_engine gen; // NNFrechet.hpp mRandomGenerator
  gen.seed(seed);                 // NNFrechet.cpp:11,20
  for (int i = 0; i < R; ++i)
    counts[i] = 0;
  counts[0] = M;                  // :418-424
  counts[R - 1] = M;              // :428-436
  int budget = W * M - 2 * M;     // :438-440
  const int begin = 1, end = R - 2; // :318-319
  for (int i = 0; i < budget; i++) {
    std::uniform_int_distribution<int> uniform(begin, end); // :324 (fresh object per draw)
    counts[uniform(gen)]++;
  }
}

extern "C" int orc_nnf_run(const void *world, int R, const double *ref_pose12, int k, int D, double check_res,
                           const uint32_t *layer_counts, const double *ik_states, uint32_t max_lazy_iters,
                           int nthreads, orc_nnf_result *res, uint32_t *knn_out, double *vertex_states,
                           double *vertex_pose12, uint32_t cap_vertices, uint32_t *path_vertices,
                           uint32_t cap_path, double *goal_cost_trace) {
  const oracle::SynthWorld *w = static_cast<const oracle::SynthWorld *>(world);
  const int dof = w->dof;
  std::memset(res, 0, sizeof(*res));
  if (R < 2 || k < 1 || D < 0)
    return -1;

  // ---- NN graph vertices: c0, c1, IK nodes (layer 0, layer R-1, interior layers) -------------
  std::vector<std::vector<double>> state; // per vertex (empty for dummies)
  std::vector<std::vector<double>> pose;  // per vertex, 12
  std::vector<std::vector<uint32_t>> layers(R);
  auto addVertex = [&](const double *q, const double *p12) {
    state.emplace_back(q ? std::vector<double>(q, q + dof) : std::vector<double>());
    pose.emplace_back(p12, p12 + 12);
    return (uint32_t)state.size() - 1;
  };
  const uint32_t c0 = addVertex(nullptr, ref_pose12);                         // :407-409
  const uint32_t c1 = addVertex(nullptr, ref_pose12 + 12 * (size_t)(R - 1)); // :412-414
  std::vector<std::pair<uint32_t, uint32_t>> edges;                          // creation order
  size_t ikCursor = 0;
  auto sampleLayer = [&](int r) { // sampleIKNodes :282-308 for the pre-drawn IK solutions of layer r
    for (uint32_t i = 0; i < layer_counts[r]; ++i) {
      const double *q = ik_states + (ikCursor++) * dof;
      double p12[12];
      w->fk(q, p12); // :295 every pose is recomputed with FK
      layers[r].push_back(addVertex(q, p12));
    }
  };
  sampleLayer(0);
  for (uint32_t v : layers[0])
    edges.emplace_back(c0, v); // :423
  if (R > 1) {
    sampleLayer(R - 1);
    for (uint32_t v : layers[R - 1])
      edges.emplace_back(v, c1); // :435
  }
  for (int r = 1; r <= R - 2; ++r) // :330-340
    sampleLayer(r);
  const uint32_t nIk = (uint32_t)ikCursor;
  res->n_ik = nIk;

  // ---- kNN to strictly deeper layers + sub-sampled edges (:452-464) ---------------------------
  // findKnnNodes for all nodes first (parallel over nodes; it reads only IK states), then the
  // graph mutations in the reference's order.
  std::vector<std::vector<uint32_t>> knn(state.size());
  {
    std::vector<uint32_t> order; // (r, node) in reference order
    std::vector<int> layerOf(state.size(), -1);
    for (int r = 0; r < R; ++r)
      for (uint32_t v : layers[r])
        layerOf[v] = r;
    auto work = [&](size_t b, size_t e) {
      std::vector<std::pair<double, uint32_t>> d;
      for (size_t v = b; v < e; ++v) {
        if (layerOf[v] < 0)
          continue;
        d.clear();
        for (int rr = layerOf[v] + 1; rr < R; ++rr) // flattenVertexList(refIndex+1 .. end), util.cpp:23-32
          for (uint32_t u : layers[rr])
            d.emplace_back(orc_distance(state[v].data(), state[u].data(), dof), u); // util.cpp:44
        std::sort(d.begin(), d.end()); // util.cpp:49: (distance, vertex)
        for (int i = 0; i < k && i < (int)d.size(); ++i) // util.cpp:52-55
          knn[v].push_back(d[i].second);
      }
    };
    const size_t nV = state.size();
    if (nthreads <= 1)
      work(0, nV);
    else {
      std::vector<std::thread> th;
      size_t chunk = (nV + nthreads - 1) / nthreads;
      for (int t = 0; t < nthreads; ++t) {
        size_t b = std::min(nV, t * chunk), e = std::min(nV, b + chunk);
        if (b < e)
          th.emplace_back(work, b, e);
      }
      for (auto &x : th)
        x.join();
    }
  }
  if (knn_out)
    for (uint32_t i = 0; i < nIk; ++i)
      for (int j = 0; j < k; ++j)
        knn_out[(size_t)i * k + j] = j < (int)knn[2 + i].size() ? knn[2 + i][j] : 0xFFFFFFFFu;
  for (int r = 0; r < R; ++r)
    for (uint32_t node : layers[r])
      for (uint32_t nbr : knn[node]) { // addSubsampledEdge :345-396
        uint32_t prev = node;
        std::vector<double> q(dof);
        for (int i = 0; i < D; ++i) {
          orc_interpolate(state[node].data(), state[nbr].data(), (1.0 + i) / (D + 1), dof, q.data()); // :366-367
          double p12[12];
          w->fk(q.data(), p12); // :372
          uint32_t a = addVertex(q.data(), p12);
          edges.emplace_back(prev, a); // :388
          if (i == D - 1)
            edges.emplace_back(a, nbr); // :391
          prev = a;
        }
      }
  const uint32_t V = (uint32_t)state.size();
  const uint32_t E = (uint32_t)edges.size();
  res->n_vertices = V;
  res->n_edges = E;
  if (V > cap_vertices)
    res->overflow = 1;
  for (uint32_t v = 0; v < V && v < cap_vertices; ++v) {
    if (vertex_states)
      for (int i = 0; i < dof; ++i)
        vertex_states[(size_t)v * dof + i] = state[v].empty() ? 0.0 : state[v][i];
    if (vertex_pose12)
      std::memcpy(vertex_pose12 + (size_t)v * 12, pose[v].data(), 12 * sizeof(double));
  }

  // predecessor lists in edge-creation order + a topological order of the NN graph
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> preds(V); // (pred vertex, edge id)
  for (uint32_t e = 0; e < E; ++e)
    preds[edges[e].second].emplace_back(edges[e].first, e);
  std::vector<uint32_t> topo;
  {
    std::vector<uint32_t> indeg(V, 0);
    std::vector<std::vector<uint32_t>> succ(V);
    for (auto &e : edges) {
      indeg[e.second]++;
      succ[e.first].push_back(e.second);
    }
    std::vector<uint32_t> stack;
    for (uint32_t v = 0; v < V; ++v)
      if (indeg[v] == 0)
        stack.push_back(v);
    while (!stack.empty()) {
      uint32_t v = stack.back();
      stack.pop_back();
      topo.push_back(v);
      for (uint32_t s : succ[v])
        if (--indeg[s] == 0)
          stack.push_back(s);
    }
    if (topo.size() != V)
      return -2; // not a DAG: cannot happen (edges only go to deeper layers)
  }

  // ---- tensor-product graph: node weights f[r][v] (:495) ---------------------------------------
  std::vector<double> f((size_t)R * V);
  for (int r = 0; r < R; ++r)
    for (uint32_t v = 0; v < V; ++v)
      f[(size_t)r * V + v] = taskDistance(ref_pose12 + 12 * (size_t)r, pose[v].data());

  // ---- LazySP (:106-152) ---------------------------------------------------------------------------
  std::vector<uint8_t> evaluated(E, 0), blocked(E, 0);
  std::vector<double> B((size_t)R * V);
  std::vector<uint32_t> finalPath;
  bool solutionFound = false;
  uint32_t iters = 0;
  while (iters < max_lazy_iters) {
    // computeShortestPath as a topological DP
    for (int r = 0; r < R; ++r)
      for (uint32_t v : topo) {
        double best = INF;
        if (r == 0 && v == c0)
          best = 0.0; // rhs(start) = 0 (LPAStar.cpp:13)
        else {
          const double fv = f[(size_t)r * V + v];
          auto relax = [&](int ru, uint32_t u) { // lookahead = max(g(u), w(u,v)) (LPAStar.cpp:31-33)
            const double bu = B[(size_t)ru * V + u];
            if (bu == INF)
              return;
            const double wuv = std::max(f[(size_t)ru * V + u], fv); // NNFrechet.cpp:530
            const double c = std::max(bu, wuv);
            if (c < best)
              best = c;
          };
          if (r > 0)
            relax(r - 1, v); // step A
          for (auto &pe : preds[v]) {
            if (blocked[pe.second])
              continue; // edge length = mInfVal (NNFrechet.cpp:692-693)
            relax(r, pe.first); // step B
            if (r > 0)
              relax(r - 1, pe.first); // step C
          }
        }
        B[(size_t)r * V + v] = best;
      }
    const double goalCost = B[(size_t)(R - 1) * V + c1];
    if (goal_cost_trace)
      goal_cost_trace[iters] = goalCost;
    ++iters;
    res->goal_cost = goalCost;
    if (goalCost == INF) // followBackpointers returns an empty path (LPAStar.cpp:174-175) -> :112-113
      break;

    // followBackpointers with the fixed tie rule; TPG path goal -> start
    std::vector<std::pair<int, uint32_t>> tpath;
    {
      int r = R - 1;
      uint32_t v = c1;
      tpath.emplace_back(r, v);
      while (!(r == 0 && v == c0)) {
        const double fv = f[(size_t)r * V + v];
        const double target = B[(size_t)r * V + v];
        int br = -1;
        uint32_t bv = 0;
        auto test = [&](int ru, uint32_t u) {
          if (br >= 0)
            return;
          const double bu = B[(size_t)ru * V + u];
          if (bu == INF)
            return;
          if (std::max(bu, std::max(f[(size_t)ru * V + u], fv)) == target) {
            br = ru;
            bv = u;
          }
        };
        if (r > 0)
          test(r - 1, v);
        for (auto &pe : preds[v]) {
          if (blocked[pe.second])
            continue;
          test(r, pe.first);
          if (r > 0)
            test(r - 1, pe.first);
        }
        if (br < 0)
          return -3;
        r = br;
        v = bv;
        tpath.emplace_back(r, v);
      }
      std::reverse(tpath.begin(), tpath.end());
    }
    // extractNNPath (:615-645)
    std::vector<uint32_t> nnPath;
    double bottleneck = 0;
    for (auto &tv : tpath) {
      if (tv.second == c0 || tv.second == c1) // :629-630
        continue;
      if (nnPath.empty() || nnPath.back() != tv.second) // :633-635
        nnPath.push_back(tv.second);
      const double c = f[(size_t)tv.first * V + tv.second];
      if (c > bottleneck) // :638-640
        bottleneck = c;
    }
    res->final_error = bottleneck;

    bool collisionFree = true;
    finalPath.clear();
    for (size_t i = 0; i + 1 < nnPath.size(); ++i) { // :120
      const uint32_t u = nnPath[i], v = nnPath[i + 1];
      uint32_t eid = 0xFFFFFFFFu;
      for (auto &pe : preds[v])
        if (pe.first == u)
          eid = pe.second;
      if (eid == 0xFFFFFFFFu)
        return -4;
      if (!evaluated[eid]) { // checkEdgeEvaluation :647-653
        evaluated[eid] = 1;  // evaluateEdge :658-660
        res->n_evaluated++;
        uint8_t ok = 0;
        orc_check_motion_batch(world, state[u].data(), state[v].data(), 1, check_res, ORC_MODE_NNF_ALPHA, 0, &ok,
                               nullptr, nullptr, 1);
        if (!ok) { // in collision
          blocked[eid] = 1; // markEdgeInCollision :685-702
          res->n_collisions++;
          collisionFree = false;
          break;
        }
      }
      if (finalPath.empty()) { // :139-144
        finalPath.push_back(u);
        finalPath.push_back(v);
      } else
        finalPath.push_back(v);
    }
    if (collisionFree) { // :148-151
      solutionFound = true;
      if (finalPath.empty() && nnPath.size() == 1)
        finalPath = nnPath; // single-vertex path (the reference would return an empty PathGeometric)
      break;
    }
  }
  res->lazy_iters = iters;
  res->solved = solutionFound ? 1 : 0;
  if (solutionFound) {
    res->path_len = (uint32_t)finalPath.size();
    for (size_t i = 0; i < finalPath.size(); ++i) {
      if (i >= cap_path) {
        res->overflow = 1;
        break;
      }
      path_vertices[i] = finalPath[i];
    }
  }
  return 0;
}

// Independent of the DP above: forward reachability over the tensor-product cells with weight <= tau,
// row by row, pushing along out-edges in a topological order of the NN graph (Kahn).
extern "C" int orc_nnf_reachable(int R, const double *ref_pose12, uint32_t V, const double *pose12, uint32_t E,
                                 const uint32_t *edges, const uint8_t *blocked, double tau, int strict) {
  std::vector<uint32_t> off(V + 1, 0), succ(E), indeg(V, 0);
  for (uint32_t e = 0; e < E; ++e) {
    if (blocked && blocked[e])
      continue;
    off[edges[2 * (size_t)e] + 1]++;
  }
  for (uint32_t v = 0; v < V; ++v)
    off[v + 1] += off[v];
  {
    std::vector<uint32_t> cur(off.begin(), off.end() - 1);
    for (uint32_t e = 0; e < E; ++e) {
      if (blocked && blocked[e])
        continue;
      succ[cur[edges[2 * (size_t)e]]++] = edges[2 * (size_t)e + 1];
      indeg[edges[2 * (size_t)e + 1]]++;
    }
  }
  std::vector<uint32_t> topo;
  topo.reserve(V);
  for (uint32_t v = 0; v < V; ++v)
    if (indeg[v] == 0)
      topo.push_back(v);
  for (size_t i = 0; i < topo.size(); ++i)
    for (uint32_t o = off[topo[i]]; o < off[topo[i] + 1]; ++o)
      if (--indeg[succ[o]] == 0)
        topo.push_back(succ[o]);
  if (topo.size() != V)
    return -2;
  std::vector<uint8_t> cur(V, 0), next(V, 0);
  cur[0] = 1; // (0, c0)
  for (int r = 0; r < R; ++r) {
    const double *pr = ref_pose12 + 12 * (size_t)r;
    std::fill(next.begin(), next.end(), 0);
    for (uint32_t v : topo) {
      if (!cur[v])
        continue;
      const double f = taskDistance(pr, pose12 + 12 * (size_t)v);
      if (strict ? !(f < tau) : !(f <= tau)) {
        cur[v] = 0;
        continue;
      }
      next[v] = 1; // step A
      for (uint32_t o = off[v]; o < off[v + 1]; ++o) {
        cur[succ[o]] = 1;  // step B
        next[succ[o]] = 1; // step C
      }
    }
    if (r == R - 1)
      return cur[1] ? 1 : 0;
    cur.swap(next);
  }
  return 0;
}

// ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle.h.
// A sub-linear CPU nearest-neighbour structure with exactly the answers of orc_knn: the baseline the
// reference would REALLY run.  pr_ompl::RRTConnect takes its trees from
// SelfConfig::getDefaultNearestNeighbors (pr_ompl/src/RRTConnect.cpp:131-134), which for a metric space
// is OMPL's NearestNeighborsGNAT (un-vendored, absent here).  A GNAT and a k-d
// tree are both exact tree searches; in R^7 with a Euclidean metric the k-d tree is the stronger CPU
// baseline, so that is what the GPU numbers are quoted against (bench.py cpu_baseline.kind "port",
// structure "kdtree").  Distances are RealVectorStateSpace::distance (orc_distance), results ordered by
// (distance, insertion index) like NearestNeighborsLinear -- tests/test_oracle.py checks equality with
// orc_knn bit for bit, ties and duplicates included.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>
#include "oracle.h"

namespace {
struct KdNode {
  uint32_t begin, end; // leaf: range of `order`
  int32_t left, right; // children (-1: leaf)
  int32_t dim;
  double split;
};
struct KdTree {
  int dim = 0;
  uint64_t n = 0;
  std::vector<double> pts;     // reordered copy, n x dim (leaf ranges contiguous)
  std::vector<uint32_t> order; // original (insertion) index of reordered point i
  std::vector<KdNode> nodes;
};
const uint32_t LEAF = 24;

int32_t buildRec(KdTree &t, const double *src, std::vector<uint32_t> &idx, uint32_t b, uint32_t e, int depthPar) {
  KdNode node;
  node.begin = b;
  node.end = e;
  node.left = node.right = -1;
  node.dim = -1;
  node.split = 0.0;
  if (e - b <= LEAF) {
    t.nodes.push_back(node);
    return (int32_t)t.nodes.size() - 1;
  }
  const int dim = t.dim;
  // widest dimension of this cell's points
  int best = 0;
  double bestW = -1.0;
  for (int d = 0; d < dim; ++d) {
    double lo = std::numeric_limits<double>::infinity(), hi = -lo;
    for (uint32_t i = b; i < e; ++i) {
      const double v = src[(size_t)idx[i] * dim + d];
      lo = std::min(lo, v);
      hi = std::max(hi, v);
    }
    if (hi - lo > bestW) {
      bestW = hi - lo;
      best = d;
    }
  }
  if (!(bestW > 0.0)) { // all points identical: one (large) leaf
    t.nodes.push_back(node);
    return (int32_t)t.nodes.size() - 1;
  }
  const uint32_t mid = b + (e - b) / 2;
  std::nth_element(idx.begin() + b, idx.begin() + mid, idx.begin() + e, [&](uint32_t x, uint32_t y) {
    return src[(size_t)x * dim + best] < src[(size_t)y * dim + best];
  });
  node.dim = best;
  node.split = src[(size_t)idx[mid] * dim + best];
  t.nodes.push_back(node);
  const int32_t me = (int32_t)t.nodes.size() - 1;
  (void)depthPar;
  const int32_t l = buildRec(t, src, idx, b, mid, 0);
  const int32_t r = buildRec(t, src, idx, mid, e, 0);
  t.nodes[me].left = l;
  t.nodes[me].right = r;
  return me;
}

struct Best { // k smallest (distance, index), kept as a max-heap
  std::vector<std::pair<double, uint32_t>> h;
  int k;
  double worst() const { return (int)h.size() < k ? std::numeric_limits<double>::infinity() : h.front().first; }
  void offer(double d, uint32_t i) {
    std::pair<double, uint32_t> c(d, i);
    if ((int)h.size() < k) {
      h.push_back(c);
      std::push_heap(h.begin(), h.end());
    } else if (c < h.front()) {
      std::pop_heap(h.begin(), h.end());
      h.back() = c;
      std::push_heap(h.begin(), h.end());
    }
  }
};

void searchRec(const KdTree &t, int32_t ni, const double *q, Best &best) {
  const KdNode &nd = t.nodes[ni];
  if (nd.left < 0) {
    for (uint32_t i = nd.begin; i < nd.end; ++i)
      best.offer(orc_distance(t.pts.data() + (size_t)i * t.dim, q, t.dim), t.order[i]);
    return;
  }
  const double diff = q[nd.dim] - nd.split;
  const int32_t nearC = diff < 0.0 ? nd.left : nd.right, farC = diff < 0.0 ? nd.right : nd.left;
  searchRec(t, nearC, q, best);
  // every point of the far side is at least |diff| away; a hair of slack covers the rounding of the
  // computed distances, and "<=" keeps equal-distance points with a lower index reachable
  if (std::fabs(diff) * (1.0 - 1e-12) <= best.worst())
    searchRec(t, farC, q, best);
}
} // namespace

extern "C" void *orc_kdtree_build(const double *pts, uint64_t n, int dim) {
  KdTree *t = new KdTree();
  t->dim = dim;
  t->n = n;
  std::vector<uint32_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0u);
  t->nodes.reserve(n / (LEAF / 2) + 16);
  if (n)
    buildRec(*t, pts, idx, 0, (uint32_t)n, 0);
  t->pts.resize(n * dim);
  for (uint64_t i = 0; i < n; ++i)
    std::memcpy(t->pts.data() + i * dim, pts + (size_t)idx[i] * dim, sizeof(double) * dim);
  t->order.swap(idx);
  return t;
}
extern "C" void orc_kdtree_destroy(void *tree) { delete static_cast<KdTree *>(tree); }

extern "C" int orc_kdtree_knn(const void *tree, const double *queries, uint64_t nq, int k, uint32_t *idx_out,
                              double *dist_out, int nthreads) {
  const KdTree *t = static_cast<const KdTree *>(tree);
  if (!t || k <= 0)
    return -1;
  auto work = [&](uint64_t b, uint64_t e) {
    Best best;
    best.k = k;
    for (uint64_t qi = b; qi < e; ++qi) {
      best.h.clear();
      if (t->n)
        searchRec(*t, 0, queries + qi * t->dim, best);
      std::sort(best.h.begin(), best.h.end());
      for (int j = 0; j < k; ++j) {
        const bool have = j < (int)best.h.size();
        idx_out[qi * k + j] = have ? best.h[j].second : 0xFFFFFFFFu;
        if (dist_out)
          dist_out[qi * k + j] = have ? best.h[j].first : std::numeric_limits<double>::infinity();
      }
    }
  };
  if (nthreads <= 1 || nq < 2)
    work(0, nq);
  else {
    std::vector<std::thread> th;
    const uint64_t chunk = (nq + nthreads - 1) / nthreads;
    for (int i = 0; i < nthreads; ++i) {
      const uint64_t b = std::min<uint64_t>(nq, i * chunk), e = std::min<uint64_t>(nq, b + chunk);
      if (b < e)
        th.emplace_back(work, b, e);
    }
    for (auto &x : th)
      x.join();
  }
  return 0;
}

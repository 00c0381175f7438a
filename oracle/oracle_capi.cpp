// ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle.h.
// kNN, world and edge-validation entry points.  The arithmetic is reached through
// the same OMPL-API objects the reference calls (RealVectorStateSpace::distance /
// interpolate, SpaceInformation::checkMotion -> DiscreteMotionValidator, a
// StateValidityChecker subclass), restated under pr-ompl_b200/ompl_min and
// oracle/defaults because OMPL itself is absent.
#include <algorithm>
#include <cmath>
#include <limits>
#include <thread>
#include <vector>
#include "oracle.h"
#include "sample_stream.h"
#include "synth_world.h"
#include "synth_checker.h"
#include "defaults/DiscreteMotionValidator.h"
#include "defaults/NearestNeighborsLinear.h"
#include "ompl/base/SpaceInformation.h"
#include "ompl/base/spaces/RealVectorStateSpace.h"

namespace ob = ompl::base;
using oracle::SynthWorld;
using oracle::SynthChecker;

namespace {
template <class F> void parallelFor(uint64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n < 2) {
    f(0, n, 0);
    return;
  }
  std::vector<std::thread> th;
  uint64_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    uint64_t b = std::min<uint64_t>(n, t * chunk), e = std::min<uint64_t>(n, b + chunk);
    if (b < e)
      th.emplace_back([=] { f(b, e, t); });
  }
  for (auto &x : th)
    x.join();
}

} // namespace

extern "C" {

double orc_distance(const double *a, const double *b, int dim) {
  // RealVectorStateSpace::distance
  double dist = 0.0;
  for (int i = 0; i < dim; ++i) {
    double diff = a[i] - b[i];
    dist += diff * diff;
  }
  return std::sqrt(dist);
}

void orc_interpolate(const double *from, const double *to, double t, int dim, double *out) {
  // RealVectorStateSpace::interpolate (Appendix B.2)
  for (int i = 0; i < dim; ++i)
    out[i] = from[i] + (to[i] - from[i]) * t;
}

// NearestNeighborsLinear::nearest / nearestK (Appendix B.4) and frechet/src/util.cpp:34-58
// (findKnnNodes: full sort of (distance, vertex) pairs, first numK) -- both reduce to
// "order by (sqrt-distance, index), take k".  k == 1 follows the first-strict-minimum scan literally.
int orc_knn(const double *pts, uint64_t n, int dim, const double *queries, uint64_t nq, int k,
            const uint32_t *lo, uint32_t *idx_out, double *dist_out, int nthreads) {
  if (k <= 0 || dim <= 0)
    return -1;
  parallelFor(nq, nthreads, [=](uint64_t qb, uint64_t qe, int) {
    std::vector<std::pair<double, uint32_t>> d;
    for (uint64_t q = qb; q < qe; ++q) {
      const double *qq = queries + q * dim;
      const uint64_t first = lo ? lo[q] : 0;
      uint32_t *io = idx_out + q * k;
      double *dd = dist_out ? dist_out + q * k : nullptr;
      for (int j = 0; j < k; ++j) {
        io[j] = 0xFFFFFFFFu;
        if (dd)
          dd[j] = std::numeric_limits<double>::infinity();
      }
      if (first >= n)
        continue;
      if (k == 1) {
        uint64_t pos = n;
        double dmin = 0.0;
        for (uint64_t i = first; i < n; ++i) {
          double distance = orc_distance(pts + i * dim, qq, dim);
          if (pos == n || dmin > distance) {
            pos = i;
            dmin = distance;
          }
        }
        io[0] = (uint32_t)pos;
        if (dd)
          dd[0] = dmin;
      } else {
        d.resize(n - first);
        for (uint64_t i = first; i < n; ++i)
          d[i - first] = std::make_pair(orc_distance(pts + i * dim, qq, dim), (uint32_t)i);
        uint64_t kk = std::min<uint64_t>(k, d.size());
        std::partial_sort(d.begin(), d.begin() + kk, d.end());
        for (uint64_t j = 0; j < kk; ++j) {
          io[j] = d[j].second;
          if (dd)
            dd[j] = d[j].first;
        }
      }
    }
  });
  return 0;
}

void *orc_world_create(int dof, const double *dh, int ns, const int32_t *sphere_link,
                       const double *sphere_xyzr, int nos, const double *obs_spheres, int nob,
                       const double *obs_boxes) {
  SynthWorld *w = new SynthWorld();
  w->dof = dof;
  for (int i = 0; i < dof; ++i) {
    w->a.push_back(dh[5 * i + 0]);
    w->d.push_back(dh[5 * i + 1]);
    w->ca.push_back(dh[5 * i + 2]);
    w->sa.push_back(dh[5 * i + 3]);
    w->theta0.push_back(dh[5 * i + 4]);
  }
  w->sphere_link.assign(sphere_link, sphere_link + ns);
  w->sphere_local.assign(sphere_xyzr, sphere_xyzr + 4 * ns);
  w->obs_sphere.assign(obs_spheres, obs_spheres + 4 * nos);
  w->obs_box.assign(obs_boxes, obs_boxes + 6 * nob);
  return w;
}
void orc_world_destroy(void *world) { delete static_cast<SynthWorld *>(world); }
double orc_clearance(const void *world, const double *q) {
  return static_cast<const SynthWorld *>(world)->clearance(q);
}
void orc_fk(const void *world, const double *q, double *pose12, double *centres) {
  const SynthWorld *w = static_cast<const SynthWorld *>(world);
  if (pose12)
    w->fk(q, pose12);
  if (centres)
    w->linkSpheres(q, centres);
}

int orc_is_valid_batch(const void *world, const double *q, uint64_t n, uint8_t *valid, double *clearance,
                       int nthreads) {
  const SynthWorld *w = static_cast<const SynthWorld *>(world);
  const int dim = w->dof;
  parallelFor(n, nthreads, [=](uint64_t b, uint64_t e, int) {
    for (uint64_t i = b; i < e; ++i) {
      double c = w->clearance(q + i * dim);
      if (valid)
        valid[i] = c > 0.0;
      if (clearance)
        clearance[i] = c;
    }
  });
  return 0;
}

int orc_check_motion_batch(const void *world, const double *from, const double *to, uint64_t n,
                           double resolution, int mode, int check_first, uint8_t *valid,
                           double *min_clearance, uint32_t *n_tested, int nthreads) {
  const SynthWorld *w = static_cast<const SynthWorld *>(world);
  const int dim = w->dof;
  if (!(resolution > 0.0))
    return -1;
  parallelFor(n, nthreads, [=](uint64_t b, uint64_t e, int) {
    // One OMPL stack per thread (DiscreteMotionValidator keeps mutable counters).
    // longestValidSegment_ is pinned to `resolution` directly (upstream derives it as
    // maxExtent * fraction; the product of two doubles need not reproduce a given value).
    struct PinnedSpace : ob::RealVectorStateSpace {
      using ob::RealVectorStateSpace::RealVectorStateSpace;
      void pin(double lvs) { longestValidSegment_ = lvs; }
      void setup() override {}
    };
    auto pspace = std::make_shared<PinnedSpace>(dim);
    pspace->setBounds(-1.0, 1.0);
    pspace->pin(resolution);
    auto psi = std::make_shared<ob::SpaceInformation>(pspace);
    auto *pchk = new SynthChecker(psi.get(), w);
    psi->setStateValidityChecker(ob::StateValidityCheckerPtr(pchk));
    psi->setMotionValidator(ob::MotionValidatorPtr(new ob::DiscreteMotionValidator(psi)));
    ob::State *s1 = psi->allocState(), *s2 = psi->allocState(), *t = psi->allocState();
    double *v1 = s1->as<ob::RealVectorStateSpace::StateType>()->values;
    double *v2 = s2->as<ob::RealVectorStateSpace::StateType>()->values;
    for (uint64_t i = b; i < e; ++i) {
      for (int k = 0; k < dim; ++k) {
        v1[k] = from[i * dim + k];
        v2[k] = to[i * dim + k];
      }
      bool ok;
      if (mode == ORC_MODE_OMPL_DISCRETE) {
        // pr_ompl/src/RRTConnect.cpp:198
        ok = check_first ? (psi->getStateValidityChecker()->isValid(s1) && psi->checkMotion(s1, s2))
                         : psi->checkMotion(s1, s2);
      } else {
        // frechet/src/NNFrechet.cpp:655-683 (evaluateEdge); returns true == in collision
        double length = pspace->distance(s1, s2);
        int numQ = (int)std::ceil(length / resolution);
        double step = 1.0 / numQ;
        bool inCollision = false;
        for (double alpha = 0.0; alpha <= 1.0; alpha += step) {
          pspace->interpolate(s1, s2, alpha, t);
          if (!psi->getStateValidityChecker()->isValid(t)) {
            inCollision = true;
            break;
          }
        }
        ok = !inCollision;
      }
      valid[i] = ok ? 1 : 0;
      if (min_clearance || n_tested) {
        // second pass without early exit over the full tested set
        pchk->track_ = true;
        pchk->minClear_ = std::numeric_limits<double>::infinity();
        pchk->count_ = 0;
        if (mode == ORC_MODE_OMPL_DISCRETE) {
          if (check_first)
            pchk->isValid(s1);
          pchk->isValid(s2);
          int nd = (int)pspace->validSegmentCount(s1, s2);
          for (int j = 1; j < nd; ++j) {
            pspace->interpolate(s1, s2, (double)j / (double)nd, t);
            pchk->isValid(t);
          }
        } else {
          double length = pspace->distance(s1, s2);
          int numQ = (int)std::ceil(length / resolution);
          double step = 1.0 / numQ;
          for (double alpha = 0.0; alpha <= 1.0; alpha += step) {
            pspace->interpolate(s1, s2, alpha, t);
            pchk->isValid(t);
          }
        }
        pchk->track_ = false;
        if (min_clearance)
          min_clearance[i] = pchk->minClear_;
        if (n_tested)
          n_tested[i] = pchk->count_;
      }
    }
    psi->freeState(s1);
    psi->freeState(s2);
    psi->freeState(t);
  });
  return 0;
}

double orc_sample_uniform(uint64_t seed, uint64_t query, uint64_t it, uint64_t dim, double lo, double hi) {
  return oracle::sampleUniform(seed, query, it, dim, lo, hi);
}
} // extern "C"

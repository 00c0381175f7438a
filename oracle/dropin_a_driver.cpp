// TEST INFRASTRUCTURE ("Drop-in A"): the reference's OWN planner source,
// /root/reference/pr_ompl/src/RRTConnect.cpp, compiled UNMODIFIED (never copied), running on top of the
// product's OMPL-facing classes:
//   ompl::NearestNeighbors<Motion*>      -> pr_gpu::GpuNearestNeighbors (prgpu_knn_*), obtained either from
//                                           SelfConfig::getDefaultNearestNeighbors (RRTConnect.cpp:131-134)
//                                           or through setNearestNeighbors<NN>() (RRTConnect.h:113-118)
//   StateValidityChecker/MotionValidator -> pr_gpu::GpuStateValidityChecker / GpuMotionValidator (prgpu_world_*)
// Exports the entry points of pr-ompl_b200/plugin/plugin_test_api.h (so tests/plugin_lib.py can drive
// either library) plus dropin_set_nn_path and dropin_rrtc_bench.  Built into oracle/_ref/libdropin_a.so
// by oracle/Makefile where /root/reference is mounted.
#include <pr_ompl/RRTConnect.h> // the reference header
#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/geometric/PathGeometric.h>
#include "plugin_test_api.h"
#include "pr_gpu/CounterSampler.h"
#include "pr_gpu/GpuNearestNeighbors.h"
#include "pr_gpu/GpuWorld.h"

namespace ob = ompl::base;
namespace og = ompl::geometric;
static thread_local std::string g_err;
static uint32_t g_first_iters = 0;
static int g_clear_between = 0;
static int g_nn_path = 0; // 0: SelfConfig default; 1: setNearestNeighbors<pr_gpu::GpuNearestNeighbors>()

namespace {
class Probe : public pr_ompl::RRTConnect {
public:
  explicit Probe(const ob::SpaceInformationPtr &si) : pr_ompl::RRTConnect(si) {}
  uint32_t dump(bool start, int dim, uint32_t cap, double *states, int32_t *parent, bool &overflow) const {
    std::vector<Motion *> m;
    (start ? tStart_ : tGoal_)->list(m);
    std::map<const Motion *, int32_t> index;
    for (std::size_t i = 0; i < m.size(); ++i)
      index[m[i]] = (int32_t)i;
    for (std::size_t i = 0; i < m.size(); ++i) {
      if (i >= cap) {
        overflow = true;
        break;
      }
      std::memcpy(states + i * dim, m[i]->state->as<ob::RealVectorStateSpace::StateType>()->values,
                  sizeof(double) * dim);
      parent[i] = m[i]->parent ? index.at(m[i]->parent) : -1;
    }
    return (uint32_t)m.size();
  }
  bool usesGpuNearestNeighbors() const {
    return dynamic_cast<pr_gpu::GpuNearestNeighbors<Motion *> *>(tStart_.get()) != nullptr &&
           dynamic_cast<pr_gpu::GpuNearestNeighbors<Motion *> *>(tGoal_.get()) != nullptr;
  }
};

ob::SpaceInformationPtr makeSpaceInformation(const pr_gpu::GpuWorldPtr &world, int dim, const double *lo,
                                             const double *hi, double resolution_fraction,
                                             pr_gpu::SampleCounter *counter) {
  auto space = std::make_shared<ob::RealVectorStateSpace>(dim);
  ob::RealVectorBounds bounds(dim);
  for (int i = 0; i < dim; ++i) {
    bounds.setLow(i, lo[i]);
    bounds.setHigh(i, hi[i]);
  }
  space->setBounds(bounds);
  space->setStateSamplerAllocator([counter](const ob::StateSpace *sp) {
    return ob::StateSamplerPtr(new pr_gpu::CounterSampler(sp, counter));
  });
  auto si = std::make_shared<ob::SpaceInformation>(space);
  si->setStateValidityChecker(ob::StateValidityCheckerPtr(new pr_gpu::GpuStateValidityChecker(si, world)));
  si->setStateValidityCheckingResolution(resolution_fraction);
  si->setMotionValidator(ob::MotionValidatorPtr(new pr_gpu::GpuMotionValidator(si, world)));
  si->setup();
  return si;
}

void configure(Probe &planner, int dim, double range, const char *extension_type) {
  planner.setRange(range);
  planner.params().setParam("extension_type", extension_type); // throws std::runtime_error on a bad string
  if (g_nn_path == 1) {
    pr_gpu::setDefaultDimension(dim);
    planner.setNearestNeighbors<pr_gpu::GpuNearestNeighbors>();
  }
}
} // namespace

extern "C" {
const char *plg_last_error(void) { return g_err.c_str(); }
void plg_set_two_phase(uint32_t first_iters, int clear_between) {
  g_first_iters = first_iters;
  g_clear_between = clear_between;
}
void dropin_set_nn_path(int path) { g_nn_path = path; }

void *plg_world_create(int device, int dof, const double *dh, int ns, const int32_t *link, const double *xyzr,
                       int nos, const double *os, int nob, const double *obx) {
  try {
    return new pr_gpu::GpuWorldPtr(new pr_gpu::GpuWorld(device, dof, dh, ns, link, xyzr, nos, os, nob, obx));
  } catch (std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void plg_world_destroy(void *world) { delete static_cast<pr_gpu::GpuWorldPtr *>(world); }

int plg_rrtc_solve(void *worldp, int dim, const double *lo, const double *hi, const double *start,
                   const double *goal, double range, const char *extension_type, double resolution_fraction,
                   uint64_t seed, uint64_t query_id, uint32_t max_iters, uint32_t max_nodes, uint32_t max_path,
                   plg_rrtc_result *res, double *start_states, int32_t *start_parent, double *goal_states,
                   int32_t *goal_parent, double *path, uint32_t *pd_counts) {
  try {
    const pr_gpu::GpuWorldPtr &world = *static_cast<pr_gpu::GpuWorldPtr *>(worldp);
    pr_gpu::SampleCounter counter;
    counter.seed = seed;
    counter.query = query_id;
    auto si = makeSpaceInformation(world, dim, lo, hi, resolution_fraction, &counter);
    ob::State *s = si->allocState(), *g = si->allocState();
    std::memcpy(s->as<ob::RealVectorStateSpace::StateType>()->values, start, sizeof(double) * dim);
    std::memcpy(g->as<ob::RealVectorStateSpace::StateType>()->values, goal, sizeof(double) * dim);
    auto pdef = std::make_shared<ob::ProblemDefinition>(si);
    pdef->setStartAndGoalStates(s, g);
    si->freeState(s);
    si->freeState(g);

    Probe planner(si);
    configure(planner, dim, range, extension_type);
    planner.setProblemDefinition(pdef);
    planner.setup();
    if (!planner.usesGpuNearestNeighbors()) {
      g_err = "the reference planner did not end up with pr_gpu::GpuNearestNeighbors";
      return -3;
    }
    ob::PlannerStatus st = ob::PlannerStatus::UNKNOWN;
    if (g_first_iters) { // solve() twice: continue the trees, or clear() and start over
      const uint32_t first = g_first_iters;
      st = planner.solve(ob::PlannerTerminationCondition([&counter, first] { return counter.count >= first; }));
      if (g_clear_between) {
        planner.clear();
        pdef->clearSolutionPaths();
        counter.count = 0;
        planner.setup();
        st = ob::PlannerStatus::UNKNOWN;
      }
    }
    if (st != ob::PlannerStatus::EXACT_SOLUTION)
      st = planner.solve(
          ob::PlannerTerminationCondition([&counter, max_iters] { return counter.count >= max_iters; }));

    std::memset(res, 0, sizeof(*res));
    res->status = (int32_t)(ob::PlannerStatus::StatusType)st;
    res->iterations = counter.count;
    bool overflow = false;
    res->n_start = planner.dump(true, dim, max_nodes, start_states, start_parent, overflow);
    res->n_goal = planner.dump(false, dim, max_nodes, goal_states, goal_parent, overflow);
    if (pdef->hasSolution()) {
      const og::PathGeometric *p = pdef->getSolutionPath()->as<og::PathGeometric>();
      res->path_len = (uint32_t)p->getStateCount();
      for (uint32_t i = 0; i < res->path_len; ++i) {
        if (i >= max_path) {
          overflow = true;
          break;
        }
        std::memcpy(path + (std::size_t)i * dim, p->getState(i)->as<ob::RealVectorStateSpace::StateType>()->values,
                    sizeof(double) * dim);
      }
    }
    res->overflow = overflow;
    if (pd_counts) {
      ob::PlannerData data(si);
      planner.getPlannerData(data);
      pd_counts[0] = data.numVertices();
      pd_counts[1] = data.numEdges();
      pd_counts[2] = data.numStartVertices();
      pd_counts[3] = data.numGoalVertices();
    }
    return 0;
  } catch (std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

/* plans n queries one after the other (one planner instance each, as an OMPL application would) and
 * returns the wall time of setup()+solve() summed over the queries, the solved count and the number of
 * growTree-level GPU round trips (nearest + validity calls = samples drawn is a proxy: iterations). */
int dropin_rrtc_bench(void *worldp, int dim, const double *lo, const double *hi, uint32_t n, const double *starts,
                      const double *goals, double range, const char *extension_type, double resolution_fraction,
                      uint64_t seed, uint32_t max_iters, double *seconds, uint32_t *n_solved,
                      uint64_t *total_iterations) {
  try {
    const pr_gpu::GpuWorldPtr &world = *static_cast<pr_gpu::GpuWorldPtr *>(worldp);
    double total = 0.0;
    uint32_t solved = 0;
    uint64_t iters = 0;
    for (uint32_t q = 0; q < n; ++q) {
      pr_gpu::SampleCounter counter;
      counter.seed = seed;
      counter.query = q;
      auto si = makeSpaceInformation(world, dim, lo, hi, resolution_fraction, &counter);
      ob::State *s = si->allocState(), *g = si->allocState();
      std::memcpy(s->as<ob::RealVectorStateSpace::StateType>()->values, starts + (size_t)q * dim, sizeof(double) * dim);
      std::memcpy(g->as<ob::RealVectorStateSpace::StateType>()->values, goals + (size_t)q * dim, sizeof(double) * dim);
      auto pdef = std::make_shared<ob::ProblemDefinition>(si);
      pdef->setStartAndGoalStates(s, g);
      si->freeState(s);
      si->freeState(g);
      Probe planner(si);
      configure(planner, dim, range, extension_type);
      planner.setProblemDefinition(pdef);
      const auto t0 = std::chrono::steady_clock::now();
      planner.setup();
      ob::PlannerStatus st =
          planner.solve(ob::PlannerTerminationCondition([&counter, max_iters] { return counter.count >= max_iters; }));
      total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      solved += st == ob::PlannerStatus::EXACT_SOLUTION ? 1 : 0;
      iters += counter.count;
    }
    *seconds = total;
    *n_solved = solved;
    if (total_iterations)
      *total_iterations = iters;
    return 0;
  } catch (std::exception &e) {
    g_err = e.what();
    return -1;
  }
}
}

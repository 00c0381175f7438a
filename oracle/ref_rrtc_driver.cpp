// ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle.h.
// Drives the reference's OWN planner, pr_ompl::RRTConnect, compiled UNMODIFIED from
// /root/reference/pr_ompl/src/RRTConnect.cpp (never copied into this repo), over the
// ompl_min API stand-in with the oracle's NearestNeighborsLinear, DiscreteMotionValidator,
// the synthetic-world StateValidityChecker and the injected counter-based sampler.
// Built into oracle/_ref/libref_rrtconnect.so by oracle/Makefile; exports ref_rrtc_solve.
#include <pr_ompl/RRTConnect.h>
#include <cstring>
#include <map>
#include "defaults/DiscreteMotionValidator.h"
#include "oracle.h"
#include "sample_stream.h"
#include "synth_checker.h"

namespace ob = ompl::base;
namespace og = ompl::geometric;

namespace {
struct SampleCounter {
  uint64_t seed, query;
  uint32_t count = 0;
};

class CounterSampler : public ob::StateSampler {
public:
  CounterSampler(const ob::StateSpace *space, SampleCounter *c) : ob::StateSampler(space), c_(c) {}
  void sampleUniform(ob::State *state) override {
    const ob::RealVectorStateSpace *sp = static_cast<const ob::RealVectorStateSpace *>(space_);
    const ob::RealVectorBounds &b = sp->getBounds();
    double *v = state->as<ob::RealVectorStateSpace::StateType>()->values;
    for (unsigned int i = 0; i < sp->getDimension(); ++i)
      v[i] = oracle::sampleUniform(c_->seed, c_->query, c_->count, i, b.low[i], b.high[i]);
    c_->count++;
  }
  void sampleUniformNear(ob::State *, const ob::State *, double) override {
    throw ompl::Exception("CounterSampler: sampleUniformNear not used by RRTConnect");
  }
  void sampleGaussian(ob::State *, const ob::State *, double) override {
    throw ompl::Exception("CounterSampler: sampleGaussian not used by RRTConnect");
  }

private:
  SampleCounter *c_;
};

// Read-only view of the planner's protected trees (no change to the reference class).
class Probe : public pr_ompl::RRTConnect {
public:
  explicit Probe(const ob::SpaceInformationPtr &si) : pr_ompl::RRTConnect(si) {}
  // returns nodes in insertion order (NearestNeighborsLinear::list) with parent indices
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
      const double *v = m[i]->state->as<ob::RealVectorStateSpace::StateType>()->values;
      std::memcpy(states + i * dim, v, sizeof(double) * dim);
      parent[i] = m[i]->parent ? index.at(m[i]->parent) : -1;
    }
    return (uint32_t)m.size();
  }
};
} // namespace

static uint32_t g_first_iters = 0;
static int g_clear_between = 0;
extern "C" void ref_set_two_phase(uint32_t first_iters, int clear_between) {
  g_first_iters = first_iters;
  g_clear_between = clear_between;
}

extern "C" int ref_rrtc_solve(const void *world, int dim, const double *lo, const double *hi,
                              const double *start, const double *goal, double range, int ext_first,
                              int ext_second, double resolution_fraction, uint64_t seed,
                              uint64_t query_id, uint32_t max_iters, uint32_t max_nodes,
                              uint32_t max_path, orc_rrtc_result *res, double *start_states,
                              int32_t *start_parent, double *goal_states, int32_t *goal_parent,
                              double *path) {
  try {
    const oracle::SynthWorld *w = static_cast<const oracle::SynthWorld *>(world);
    auto space = std::make_shared<ob::RealVectorStateSpace>(dim);
    ob::RealVectorBounds bounds(dim);
    for (int i = 0; i < dim; ++i) {
      bounds.setLow(i, lo[i]);
      bounds.setHigh(i, hi[i]);
    }
    space->setBounds(bounds);
    SampleCounter counter{seed, query_id, 0};
    space->setStateSamplerAllocator([&counter](const ob::StateSpace *sp) {
      return ob::StateSamplerPtr(new CounterSampler(sp, &counter));
    });
    auto si = std::make_shared<ob::SpaceInformation>(space);
    si->setStateValidityChecker(ob::StateValidityCheckerPtr(new oracle::SynthChecker(si.get(), w)));
    si->setStateValidityCheckingResolution(resolution_fraction);
    si->setMotionValidator(ob::MotionValidatorPtr(new ob::DiscreteMotionValidator(si)));
    si->setup();

    ob::State *s = si->allocState(), *g = si->allocState();
    std::memcpy(s->as<ob::RealVectorStateSpace::StateType>()->values, start, sizeof(double) * dim);
    std::memcpy(g->as<ob::RealVectorStateSpace::StateType>()->values, goal, sizeof(double) * dim);
    auto pdef = std::make_shared<ob::ProblemDefinition>(si);
    pdef->setStartAndGoalStates(s, g);
    si->freeState(s);
    si->freeState(g);

    Probe planner(si);
    planner.setRange(range);
    planner.setExtensionTypes(std::make_pair(
        ext_first ? pr_ompl::RRTConnect::EXTTYPE_CONNECT : pr_ompl::RRTConnect::EXTTYPE_EXTEND,
        ext_second ? pr_ompl::RRTConnect::EXTTYPE_CONNECT : pr_ompl::RRTConnect::EXTTYPE_EXTEND));
    planner.setProblemDefinition(pdef);
    planner.setup();
    ob::PlannerStatus st = ob::PlannerStatus::UNKNOWN;
    if (g_first_iters) { // solve() twice: continue the trees (RRTConnect.cpp:230-236 finds no new start), or
                         // clear() (:166-176) and start over with the sample counter back at 0
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
    if (st != ob::PlannerStatus::EXACT_SOLUTION) {
      ob::PlannerTerminationCondition ptc([&counter, max_iters] { return counter.count >= max_iters; });
      st = planner.solve(ptc);
    }

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
        std::memcpy(path + (std::size_t)i * dim,
                    p->getState(i)->as<ob::RealVectorStateSpace::StateType>()->values,
                    sizeof(double) * dim);
      }
    }
    res->overflow = overflow ? 1 : 0;
    return 0;
  } catch (std::exception &e) {
    std::fprintf(stderr, "ref_rrtc_solve: %s\n", e.what());
    return -1;
  }
}

// ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle.h.
// Flat-array CPU restatement ("port") of pr_ompl::RRTConnect, function by function:
//   growTree  <- pr_ompl/src/RRTConnect.cpp:178-217
//   solve     <- pr_ompl/src/RRTConnect.cpp:219-376
// with the OMPL primitives it calls restated from SURVEY.md Appendix B
// (RealVectorStateSpace::distance/interpolate B.1/B.2, DiscreteMotionValidator B.3,
// NearestNeighborsLinear::nearest B.4, configurePlannerRange B.5, PlannerInputStates with a
// single GoalState B.7).  It deliberately shares no code with ref_rrtc_driver.cpp: the
// tests require both to produce identical trees (tests/test_oracle_rrtconnect.py), which
// pins this port -- and the ompl_min stand-in -- against the reference's own source.
#include <cmath>
#include <cstring>
#include <vector>
#include "oracle.h"
#include "sample_stream.h"
#include "synth_world.h"

namespace {
enum GrowState { TRAPPED, ADVANCED, REACHED }; // RRTConnect.h:165-173
enum { ST_INVALID_START = 1, ST_INVALID_GOAL = 2, ST_TIMEOUT = 4, ST_EXACT = 6 };

struct Tree {
  int dim;
  std::vector<double> x;      // n x dim, insertion order
  std::vector<int32_t> parent;
  std::size_t size() const { return parent.size(); }
  const double *at(std::size_t i) const { return x.data() + i * dim; }
  int32_t add(const double *q, int32_t par) {
    x.insert(x.end(), q, q + dim);
    parent.push_back(par);
    return (int32_t)parent.size() - 1;
  }
};

struct Ctx {
  const oracle::SynthWorld *w;
  int dim;
  double maxDistance; // range
  double lvs;         // longest valid segment
};

double dist(const Ctx &c, const double *a, const double *b) { return orc_distance(a, b, c.dim); }

// Appendix B.4: first strict minimum in insertion order
int32_t nearest(const Ctx &c, const Tree &t, const double *q) {
  std::size_t pos = t.size();
  double dmin = 0.0;
  for (std::size_t i = 0; i < t.size(); ++i) {
    double d = dist(c, t.at(i), q);
    if (pos == t.size() || dmin > d) {
      pos = i;
      dmin = d;
    }
  }
  return (int32_t)pos;
}

// Appendix B.3: s1 assumed valid; s2 first; interior points j/nd, AND of all (order irrelevant
// to the result, so the plain 1..nd-1 order is used here).
bool checkMotion(const Ctx &c, const double *s1, const double *s2) {
  if (!c.w->isValid(s2))
    return false;
  int nd = (int)(unsigned int)std::ceil(dist(c, s1, s2) / c.lvs);
  std::vector<double> t(c.dim);
  for (int j = 1; j < nd; ++j) {
    orc_interpolate(s1, s2, (double)j / (double)nd, c.dim, t.data());
    if (!c.w->isValid(t.data()))
      return false;
  }
  return true;
}

struct Tgi { // RRTConnect.h:157-162
  std::vector<double> xstate;
  int32_t xmotion;
  bool start;
};

// RRTConnect.cpp:178-217
GrowState growTree(const Ctx &c, Tree &tree, Tgi &tgi, const double *rstate) {
  int32_t nmotion = nearest(c, tree, rstate);                       // :181
  bool reach = true;
  const double *dstate = rstate;
  double d = dist(c, tree.at(nmotion), rstate);                     // :188
  if (d > c.maxDistance) {                                          // :189
    orc_interpolate(tree.at(nmotion), rstate, c.maxDistance / d, c.dim, tgi.xstate.data()); // :191
    dstate = tgi.xstate.data();
    reach = false;
  }
  // :198 (goal tree validates in reverse, new state first)
  std::vector<double> near(tree.at(nmotion), tree.at(nmotion) + c.dim); // tree.x may reallocate on add
  bool validMotion = tgi.start ? checkMotion(c, near.data(), dstate)
                               : (c.w->isValid(dstate) && checkMotion(c, dstate, near.data()));
  if (validMotion) {                                                // :200-213
    tgi.xmotion = tree.add(dstate, nmotion);
    return reach ? REACHED : ADVANCED;
  }
  return TRAPPED;                                                   // :215
}
} // namespace

extern "C" int orc_rrtc_solve(const void *world, int dim, const double *lo, const double *hi,
                              const double *start, const double *goal, double range, int ext_first,
                              int ext_second, double resolution_fraction, uint64_t seed,
                              uint64_t query_id, uint32_t max_iters, uint32_t max_nodes,
                              uint32_t max_path, orc_rrtc_result *res, double *start_states,
                              int32_t *start_parent, double *goal_states, int32_t *goal_parent,
                              double *path) {
  Ctx c;
  c.w = static_cast<const oracle::SynthWorld *>(world);
  c.dim = dim;
  double ext = 0.0; // RealVectorStateSpace::getMaximumExtent
  for (int i = 0; i < dim; ++i) {
    double e = hi[i] - lo[i];
    ext += e * e;
  }
  ext = std::sqrt(ext);
  c.maxDistance = range;
  if (c.maxDistance < 2.220446049250313e-16) // Appendix B.5 (RRTConnect.cpp:128-129)
    c.maxDistance = ext * 0.2;
  c.lvs = ext * resolution_fraction;

  Tree tStart{dim, {}, {}}, tGoal{dim, {}, {}};
  std::memset(res, 0, sizeof(*res));
  auto inBounds = [&](const double *q) { // RealVectorStateSpace::satisfiesBounds
    for (int i = 0; i < dim; ++i)
      if (q[i] - 2.220446049250313e-16 > hi[i] || q[i] + 2.220446049250313e-16 < lo[i])
        return false;
    return true;
  };
  auto finish = [&](int status, uint32_t iters, const std::vector<std::vector<double>> &p) {
    res->status = status;
    res->iterations = iters;
    res->n_start = (uint32_t)tStart.size();
    res->n_goal = (uint32_t)tGoal.size();
    res->path_len = (uint32_t)p.size();
    bool overflow = false;
    auto dump = [&](const Tree &t, double *st, int32_t *par) {
      for (std::size_t i = 0; i < t.size(); ++i) {
        if (i >= max_nodes) {
          overflow = true;
          break;
        }
        std::memcpy(st + i * dim, t.at(i), sizeof(double) * dim);
        par[i] = t.parent[i];
      }
    };
    dump(tStart, start_states, start_parent);
    dump(tGoal, goal_states, goal_parent);
    for (std::size_t i = 0; i < p.size(); ++i) {
      if (i >= max_path) {
        overflow = true;
        break;
      }
      std::memcpy(path + i * dim, p[i].data(), sizeof(double) * dim);
    }
    res->overflow = overflow;
    return 0;
  };

  // :230-242  start states (PlannerInputStates::nextStart: in bounds and valid)
  if (inBounds(start) && c.w->isValid(start))
    tStart.add(start, -1);
  if (tStart.size() == 0)
    return finish(ST_INVALID_START, 0, {});
  // :244-248 couldSample() is true for a GoalState

  Tgi tgi;
  tgi.xstate.resize(dim);
  std::vector<double> rstate(dim);
  bool startTree = true;
  unsigned sampledGoals = 0; // PlannerInputStates::sampledGoalsCount_
  uint32_t it = 0;           // samples drawn

  while (it < max_iters) { // :263  (ptc: sample budget)
    Tree &tree = startTree ? tStart : tGoal;                        // :265
    tgi.start = startTree;
    startTree = !startTree;                                         // :267
    Tree &otherTree = startTree ? tStart : tGoal;

    if (tGoal.size() == 0 || sampledGoals < tGoal.size() / 2) {     // :270
      // nextGoal for a single GoalState: one sample ever (maxSampleCount()==1), Appendix B.7
      if (sampledGoals < 1) {
        sampledGoals++;
        if (inBounds(goal) && c.w->isValid(goal))
          tGoal.add(goal, -1);                                      // :275-278
      }
      if (tGoal.size() == 0)                                        // :281-285
        break;
    }

    for (int i = 0; i < dim; ++i)                                   // :289
      rstate[i] = oracle::sampleUniform(seed, query_id, it, i, lo[i], hi[i]);
    ++it;

    GrowState gs;
    tgi.xmotion = -1;                                               // :295
    do {
      gs = growTree(c, tree, tgi, rstate.data());
    } while (gs == ADVANCED && ext_first == 1);                     // :296-300
    if (tgi.xmotion < 0)                                            // :303
      continue;
    int32_t addedMotion = tgi.xmotion;                              // :307
    tgi.xmotion = -1;
    if (gs != REACHED)                                              // :313-314
      std::memcpy(rstate.data(), tree.at(addedMotion), sizeof(double) * dim);
    tgi.start = startTree;                                          // :316
    do {
      gs = growTree(c, otherTree, tgi, rstate.data());
    } while (gs == ADVANCED && ext_second == 1);                    // :317-321

    int32_t startMotion = startTree ? tgi.xmotion : addedMotion;    // :323-324
    int32_t goalMotion = startTree ? addedMotion : tgi.xmotion;
    if (gs == REACHED) {                                            // :327 (pair always valid for GoalState)
      if (tStart.parent[startMotion] >= 0)                          // :332-335
        startMotion = tStart.parent[startMotion];
      else
        goalMotion = tGoal.parent[goalMotion];
      std::vector<std::vector<double>> p1, p;
      for (int32_t s = startMotion; s >= 0; s = tStart.parent[s])   // :340-346
        p1.emplace_back(tStart.at(s), tStart.at(s) + dim);
      for (int i = (int)p1.size() - 1; i >= 0; --i)                 // :358-359
        p.push_back(p1[i]);
      for (int32_t s = goalMotion; s >= 0; s = tGoal.parent[s])     // :348-354,360-361
        p.emplace_back(tGoal.at(s), tGoal.at(s) + dim);
      return finish(ST_EXACT, it, p);
    }
  }
  return finish(ST_TIMEOUT, it, {});
}

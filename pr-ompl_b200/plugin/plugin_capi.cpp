// Test driver for the OMPL-plugin mirror; see plugin_test_api.h.
#include "plugin_test_api.h"
#include <cmath>
#include <cstring>
#include <string>
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/geometric/PathGeometric.h>
#include <pr_ompl/RRTConnect.h>
#include <NNFrechet.hpp>
#include "pr_gpu/BatchedRRTConnect.h"
#include "pr_gpu/CounterSampler.h"
#include "pr_gpu/GpuWorld.h"

namespace ob = ompl::base;
namespace og = ompl::geometric;
static thread_local std::string g_err;
static uint32_t g_first_iters = 0;
static int g_clear_between = 0;

extern "C" {
const char *plg_last_error(void) { return g_err.c_str(); }
void plg_set_two_phase(uint32_t first_iters, int clear_between) {
  g_first_iters = first_iters;
  g_clear_between = clear_between;
}

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

static ob::SpaceInformationPtr makeSpaceInformation(const pr_gpu::GpuWorldPtr &world, int dim, const double *lo,
                                                    const double *hi, double resolution_fraction,
                                                    pr_gpu::SampleCounter *counter) {
  auto space = std::make_shared<ob::RealVectorStateSpace>(dim);
  ob::RealVectorBounds bounds(dim);
  for (int i = 0; i < dim; ++i) {
    bounds.setLow(i, lo[i]);
    bounds.setHigh(i, hi[i]);
  }
  space->setBounds(bounds);
  if (counter)
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

    pr_ompl::RRTConnect planner(si); // the drop-in name
    planner.params().setParam("range", std::to_string(range));
    planner.setRange(range);
    if (!planner.params().setParam("extension_type", extension_type)) {
      g_err = "Invalid extension type string.";
      return -2;
    }
    planner.setProblemDefinition(pdef);
    planner.setup();
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
    res->n_start = (uint32_t)planner.treeSize(0);
    res->n_goal = (uint32_t)planner.treeSize(1);
    bool overflow = false;
    for (int side = 0; side < 2; ++side) {
      double *states = side == 0 ? start_states : goal_states;
      int32_t *parent = side == 0 ? start_parent : goal_parent;
      for (std::size_t i = 0; i < planner.treeSize(side); ++i) {
        if (i >= max_nodes) {
          overflow = true;
          break;
        }
        std::memcpy(states + i * dim,
                    planner.treeState(side, i)->as<ob::RealVectorStateSpace::StateType>()->values,
                    sizeof(double) * dim);
        parent[i] = planner.treeParent(side, i);
      }
    }
    if (pdef->hasSolution()) {
      const og::PathGeometric *p = pdef->getSolutionPath()->as<og::PathGeometric>();
      res->path_len = (uint32_t)p->getStateCount();
      for (uint32_t i = 0; i < res->path_len; ++i) {
        if (i >= max_path) {
          overflow = true;
          break;
        }
        std::memcpy(path + (std::size_t)i * dim,
                    p->getState(i)->as<ob::RealVectorStateSpace::StateType>()->values, sizeof(double) * dim);
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

int plg_batched_rrtc_solve(void *worldp, int dim, const double *lo, const double *hi, uint32_t n,
                           const double *starts, const double *goals, double range, const char *extension_type,
                           double resolution_fraction, uint64_t seed, uint64_t query_id_base, uint32_t max_iters,
                           uint32_t max_nodes, int32_t *status, uint32_t *iterations, uint32_t *n_start,
                           uint32_t *n_goal, uint32_t *path_len, double *paths, uint32_t max_path) {
  try {
    const pr_gpu::GpuWorldPtr &world = *static_cast<pr_gpu::GpuWorldPtr *>(worldp);
    ob::RealVectorBounds bounds(dim);
    for (int i = 0; i < dim; ++i) {
      bounds.setLow(i, lo[i]);
      bounds.setHigh(i, hi[i]);
    }
    pr_gpu::BatchedRRTConnect batch(world, bounds, n);
    batch.setRange(range);
    batch.setStateValidityCheckingResolution(resolution_fraction);
    batch.setExtensionTypesString(extension_type);
    batch.setSampleStream(seed, query_id_base);
    batch.setLimits(max_iters, max_nodes);
    std::vector<pr_gpu::BatchedRRTConnect::Result> r = batch.solve(starts, goals);
    for (uint32_t q = 0; q < n; ++q) {
      status[q] = (int32_t)r[q].status;
      iterations[q] = r[q].iterations;
      n_start[q] = r[q].startTreeSize;
      n_goal[q] = r[q].goalTreeSize;
      std::vector<std::vector<double>> p = batch.path(q);
      path_len[q] = (uint32_t)p.size();
      for (std::size_t i = 0; i < p.size() && i < max_path; ++i)
        std::memcpy(paths + ((std::size_t)q * max_path + i) * dim, p[i].data(), sizeof(double) * dim);
    }
    return 0;
  } catch (std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

int plg_selftest_interfaces(void *worldp, const double *states, uint32_t n, uint8_t *is_valid_out,
                            uint8_t *motion_out, uint32_t *nearest_out) {
  try {
    const pr_gpu::GpuWorldPtr &world = *static_cast<pr_gpu::GpuWorldPtr *>(worldp);
    const int dim = world->dof();
    std::vector<double> lo(dim, -3.14159265358979323846), hi(dim, 3.14159265358979323846);
    auto si = makeSpaceInformation(world, dim, lo.data(), hi.data(), 0.01, nullptr);
    std::vector<ob::State *> st(n);
    for (uint32_t i = 0; i < n; ++i) {
      st[i] = si->allocState();
      std::memcpy(st[i]->as<ob::RealVectorStateSpace::StateType>()->values, states + (std::size_t)i * dim,
                  sizeof(double) * dim);
    }
    for (uint32_t i = 0; i < n; ++i)
      is_valid_out[i] = si->isValid(st[i]) ? 1 : 0;
    for (uint32_t i = 0; i + 1 < n; ++i)
      motion_out[i] = si->checkMotion(st[i], st[i + 1]) ? 1 : 0;
    // ompl::NearestNeighbors<State*> over the GPU: nearest of each state among the ones before it
    pr_gpu::GpuNearestNeighbors<ob::State *> nn;
    nn.setCoordinateAccessor(
        [](ob::State *const &s) { return (const double *)s->as<ob::RealVectorStateSpace::StateType>()->values; },
        dim);
    nn.setDistanceFunction([&si](ob::State *const &a, ob::State *const &b) { return si->distance(a, b); });
    for (uint32_t i = 0; i < n; ++i) {
      if (i == 0)
        nearest_out[i] = 0xFFFFFFFFu;
      else {
        ob::State *nb = nn.nearest(st[i]);
        uint32_t j = 0;
        while (st[j] != nb)
          ++j;
        nearest_out[i] = j;
      }
      nn.add(st[i]);
    }
    std::vector<ob::State *> all;
    nn.list(all);
    bool ok = all.size() == n && nn.size() == n;
    std::vector<ob::State *> nbh;
    nn.nearestK(st[0], 3, nbh);
    ok = ok && nbh.size() == std::min<std::size_t>(3, n) && nbh[0] == st[0];
    // the 3-argument checkMotion agrees with the 2-argument one and reports a lastValid fraction
    for (uint32_t i = 0; i + 1 < n && i < 16; ++i) {
      std::pair<ob::State *, double> lastValid(si->allocState(), -1.0);
      bool v3 = si->checkMotion(st[i], st[i + 1], lastValid);
      ok = ok && (v3 == (motion_out[i] != 0));
      if (!v3)
        ok = ok && lastValid.second >= 0.0 && lastValid.second < 1.0;
      si->freeState(lastValid.first);
    }
    nn.clear();
    ok = ok && nn.size() == 0;
    for (auto *s : st)
      si->freeState(s);
    if (!ok)
      g_err = "interface self-test failed";
    return ok ? 0 : 1;
  } catch (std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

namespace {
// metric_kind 0: Euclidean distance of the translations (the metric the device models); 1: translation
// distance + 0.25 x Frobenius norm of the rotation difference (a user metric the device has no model of)
double taskMetric(int kind, Eigen::Isometry3d &a, Eigen::Isometry3d &b) {
  double s = 0.0;
  for (int r = 0; r < 3; ++r)
    s += (a(r, 3) - b(r, 3)) * (a(r, 3) - b(r, 3));
  double d = std::sqrt(s);
  if (kind == 1) {
    double f = 0.0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        f += (a(r, c) - b(r, c)) * (a(r, c) - b(r, c));
    d = d + 0.25 * std::sqrt(f);
  }
  return d;
}
int nnfRun(void *worldp, int metric_kind, int with_world, int *used_host_form, int W, int M, int k, int D, int seed,
           const double *ref_pose12, uint32_t P, const double *ik_states, uint32_t n_ik_available,
           double solve_seconds, plg_nnf_result *res, uint32_t *layer_counts, uint32_t *path_vertices,
           uint32_t cap_path, double *path_states);
} // namespace

int plg_nnf_run(void *worldp, int W, int M, int k, int D, int seed, const double *ref_pose12, uint32_t P,
                const double *ik_states, uint32_t n_ik_available, double solve_seconds, plg_nnf_result *res,
                uint32_t *layer_counts, uint32_t *path_vertices, uint32_t cap_path, double *path_states) {
  return nnfRun(worldp, 0, 1, nullptr, W, M, k, D, seed, ref_pose12, P, ik_states, n_ik_available, solve_seconds, res,
                layer_counts, path_vertices, cap_path, path_states);
}
int plg_nnf_run_host(void *worldp, int metric_kind, int with_world, int *used_host_form, int W, int M, int k, int D,
                     int seed, const double *ref_pose12, uint32_t P, const double *ik_states,
                     uint32_t n_ik_available, double solve_seconds, plg_nnf_result *res, uint32_t *layer_counts,
                     uint32_t *path_vertices, uint32_t cap_path, double *path_states) {
  return nnfRun(worldp, metric_kind, with_world, used_host_form, W, M, k, D, seed, ref_pose12, P, ik_states,
                n_ik_available, solve_seconds, res, layer_counts, path_vertices, cap_path, path_states);
}
} // extern "C"

namespace {
int nnfRun(void *worldp, int metric_kind, int with_world, int *used_host_form, int W, int M, int k, int D, int seed,
           const double *ref_pose12, uint32_t P, const double *ik_states, uint32_t n_ik_available,
           double solve_seconds, plg_nnf_result *res, uint32_t *layer_counts, uint32_t *path_vertices,
           uint32_t cap_path, double *path_states) {
  try {
    const pr_gpu::GpuWorldPtr &world = *static_cast<pr_gpu::GpuWorldPtr *>(worldp);
    const int dim = world->dof();
    std::vector<double> lo(dim, -3.14159265358979323846), hi(dim, 3.14159265358979323846);
    auto si = makeSpaceInformation(world, dim, lo.data(), hi.data(), 0.01, nullptr);
    NNFrechet::NNFrechet planner(si, W, M, k, D, seed); // the drop-in name
    std::vector<Eigen::Isometry3d> ref(P);
    for (uint32_t i = 0; i < P; ++i)
      std::memcpy(ref[i].data(), ref_pose12 + (std::size_t)i * 12, 12 * sizeof(double));
    planner.setRefPath(ref);
    if (with_world)
      planner.setWorld(world);
    planner.setFKFunc([&](ob::State *s) {
      Eigen::Isometry3d out;
      double pose[12];
      pr_gpu::check(prgpu_fk_batch(world->handle(), s->as<ob::RealVectorStateSpace::StateType>()->values, 1, pose),
                    "fk callback");
      std::memcpy(out.data(), pose, sizeof(pose));
      return out;
    });
    uint32_t cursor = 0;
    planner.setIKFunc([&](Eigen::Isometry3d &, int n) {
      std::vector<ob::State *> out;
      for (int i = 0; i < n && cursor < n_ik_available; ++i, ++cursor) {
        ob::State *s = si->allocState();
        std::memcpy(s->as<ob::RealVectorStateSpace::StateType>()->values, ik_states + (std::size_t)cursor * dim,
                    sizeof(double) * dim);
        out.push_back(s);
      }
      return out;
    });
    planner.setDistanceFunc(
        [metric_kind](Eigen::Isometry3d &a, Eigen::Isometry3d &b) { return taskMetric(metric_kind, a, b); });
    auto pdef = std::make_shared<ob::ProblemDefinition>(si);
    planner.setProblemDefinition(pdef);
    planner.setup();
    if (used_host_form)
      *used_host_form = planner.usesHostCallbacks() ? 1 : 0;
    ob::PlannerStatus st = planner.solve(solve_seconds);
    std::memset(res, 0, sizeof(*res));
    res->status = (int32_t)(ob::PlannerStatus::StatusType)st;
    res->lazy_iters = planner.lazyIterations();
    res->n_evaluated = planner.evaluatedEdges();
    res->n_collisions = planner.collidingEdges();
    res->final_error = planner.finalError();
    res->goal_cost = planner.goalCost();
    res->R = (uint32_t)planner.layerCounts().size();
    for (std::size_t i = 0; i < planner.layerCounts().size(); ++i)
      layer_counts[i] = planner.layerCounts()[i];
    res->path_len = (uint32_t)planner.pathVertices().size();
    for (uint32_t i = 0; i < res->path_len && i < cap_path; ++i)
      path_vertices[i] = planner.pathVertices()[i];
    if (pdef->hasSolution()) {
      const og::PathGeometric *p = pdef->getSolutionPath()->as<og::PathGeometric>();
      for (uint32_t i = 0; i < p->getStateCount() && i < cap_path; ++i)
        std::memcpy(path_states + (std::size_t)i * dim,
                    p->getState(i)->as<ob::RealVectorStateSpace::StateType>()->values, sizeof(double) * dim);
    }
    return 0;
  } catch (std::exception &e) {
    g_err = e.what();
    return -1;
  }
}
} // namespace

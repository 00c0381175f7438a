// ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle.h.
// Drives the reference's OWN NNF planner, NNFrechet::NNFrechet, compiled UNMODIFIED from
// /root/reference/frechet/src/{NNFrechet,LPAStar,MinPriorityQueue,util}.cpp (never copied into this
// repo) over the ompl_min API stand-in, the Boost.Graph / boost::timer stand-ins of oracle/compat/ and
// the Eigen::Isometry3d value-type stand-in.  User callbacks (the reference takes FK / IK / task
// distance / validity as opaque std::functions, NNFrechet.hpp:43-59) are bound to the synthetic world:
//   FK        = SynthWorld::fk                                  (same call as oracle/nnf_port.cpp)
//   IK        = hands out the next `numSolutions` pre-drawn states of `ik_states` (the reference's
//               call order is first waypoint, last waypoint, then waypoints 1..R-2, NNFrechet.cpp:418-448)
//   distance  = Euclidean distance of the pose translations (ref_nnf_set_metric(1): plus a rotation term)
//   validity  = SynthChecker (clearance > 0)
// Built into oracle/_ref/libref_frechet.so (and libref_frechet_rev.so with the edge-list iteration
// order reversed) by oracle/Makefile; exports ref_nnf_run.
//
// The planner's result fields are private (NNFrechet.hpp:132-312, LPAStar.h:75-95).  This translation
// unit alone re-spells the access keywords while including the reference headers; the reference's
// own .cpp files are compiled with no such definition, and access control does not change class layout.
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>
#include <Eigen/Dense>
#include <boost/graph/adjacency_list.hpp>
#include <ompl/base/Planner.h>
#include <ompl/base/ProblemDefinition.h>
#include <ompl/base/SpaceInformation.h>
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/geometric/PathGeometric.h>
#define private public
#define protected public
#include "NNFrechet.hpp"
#undef private
#undef protected
#include "oracle.h"
#include "synth_checker.h"

namespace ob = ompl::base;
namespace og = ompl::geometric;

namespace {
Eigen::Isometry3d toIso(const double *p12) {
  Eigen::Isometry3d T;
  std::memcpy(T.data(), p12, 12 * sizeof(double));
  return T;
}
double translationDistance(Eigen::Isometry3d &a, Eigen::Isometry3d &b) {
  const double dx = a(0, 3) - b(0, 3), dy = a(1, 3) - b(1, 3), dz = a(2, 3) - b(2, 3);
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}
// a user metric with a rotation term (kind 1): translation distance + 0.25 x Frobenius norm of the
// rotation difference -- exercises the host-callback form of the product
int g_metric_kind = 0;
double taskMetric(Eigen::Isometry3d &a, Eigen::Isometry3d &b) {
  double d = translationDistance(a, b);
  if (g_metric_kind == 1) {
    double f = 0.0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        f += (a(r, c) - b(r, c)) * (a(r, c) - b(r, c));
    d = d + 0.25 * std::sqrt(f);
  }
  return d;
}
const double *values(const ob::State *s) { return s->as<ob::RealVectorStateSpace::StateType>()->values; }
} // namespace

extern "C" void ref_nnf_set_metric(int kind) { g_metric_kind = kind; }

extern "C" int ref_nnf_run(const void *world, uint32_t n_ref_full, const double *ref_full_pose12, int W, int M,
                           int k, int D, int seed, const double *ik_states, uint32_t n_ik_avail,
                           double solve_seconds, int search_only, ref_nnf_result *res,
                           double *ref_sub_pose12 /* cap_ref x 12 */, uint32_t cap_ref,
                           uint32_t *ik_call_counts /* cap_ref: requested per waypoint, in waypoint order */,
                           uint32_t *knn_out /* n_ik x k */, double *vertex_states, double *vertex_pose12,
                           uint32_t cap_vertices, uint32_t *nn_edges /* cap_edges x 2, creation order */,
                           uint8_t *nn_edge_flags /* bit0 evaluated, bit1 knocked out */, uint32_t cap_edges,
                           double *path_states, uint32_t cap_path) {
  std::memset(res, 0, sizeof(*res));
  try {
    const oracle::SynthWorld *w = static_cast<const oracle::SynthWorld *>(world);
    const int dof = w->dof;
    auto space = std::make_shared<ob::RealVectorStateSpace>(dof);
    ob::RealVectorBounds bounds(dof);
    bounds.setLow(-M_PI);
    bounds.setHigh(M_PI);
    space->setBounds(bounds);
    auto si = std::make_shared<ob::SpaceInformation>(space);
    si->setStateValidityChecker(ob::StateValidityCheckerPtr(new oracle::SynthChecker(si.get(), w)));
    // NNFrechet never calls si_->checkMotion; no motion validator is installed.
    auto pdef = std::make_shared<ob::ProblemDefinition>(si);

    std::vector<Eigen::Isometry3d> refPath;
    for (uint32_t i = 0; i < n_ref_full; ++i)
      refPath.push_back(toIso(ref_full_pose12 + 12 * (size_t)i));

    NNFrechet::NNFrechet planner(si, W, M, k, D, seed);
    planner.setProblemDefinition(pdef);
    planner.setRefPath(refPath);
    planner.setFKFunc([w](ob::State *s) {
      double p12[12];
      w->fk(values(s), p12);
      return toIso(p12);
    });
    size_t cursor = 0;
    bool ikUnderflow = false;
    std::vector<uint32_t> callCounts; // in call order: waypoint 0, waypoint R-1, waypoints 1..R-2
    planner.setIKFunc([&](Eigen::Isometry3d &, int numSolutions) {
      callCounts.push_back((uint32_t)numSolutions);
      std::vector<ob::State *> out;
      for (int i = 0; i < numSolutions; ++i) {
        if (cursor >= n_ik_avail) {
          ikUnderflow = true;
          break;
        }
        ob::State *s = space->allocState();
        std::memcpy(s->as<ob::RealVectorStateSpace::StateType>()->values, ik_states + (cursor++) * dof,
                    dof * sizeof(double));
        out.push_back(s);
      }
      return out;
    });
    planner.setDistanceFunc(&taskMetric);

    auto t0 = std::chrono::steady_clock::now();
    planner.setup();
    res->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (ikUnderflow)
      return -2;

    const Graph &nn = planner.mNNGraph;
    const Graph &tpg = planner.mTensorProductGraph;
    const uint32_t R = (uint32_t)planner.mReferencePath.size();
    res->R = R;
    res->n_vertices = (uint32_t)num_vertices(nn);
    res->n_edges = (uint32_t)num_edges(nn);
    res->n_tpg_vertices = (uint64_t)num_vertices(tpg);
    res->n_tpg_edges = (uint64_t)num_edges(tpg);
    res->n_ik = (uint32_t)cursor;
    if (R > cap_ref || res->n_vertices > cap_vertices || res->n_edges > cap_edges)
      res->overflow = 1;
    for (uint32_t r = 0; r < R && r < cap_ref; ++r) {
      std::memcpy(ref_sub_pose12 + 12 * (size_t)r, planner.mReferencePath[r].data(), 12 * sizeof(double));
      // call order -> waypoint order
      uint32_t call = r == 0 ? 0 : (r == R - 1 ? 1 : r + 1);
      ik_call_counts[r] = call < callCounts.size() ? callCounts[call] : 0xFFFFFFFFu;
    }
    for (uint32_t v = 0; v < res->n_vertices && v < cap_vertices; ++v) {
      const VProp &p = nn.m_vertices[v].prop;
      for (int i = 0; i < dof; ++i)
        vertex_states[(size_t)v * dof + i] = p.state ? values(p.state)[i] : 0.0;
      std::memcpy(vertex_pose12 + 12 * (size_t)v, p.poseEE.data(), 12 * sizeof(double));
    }
    // findKnnNodes choices, recovered from the graph: the j-th out-edge (creation order) of IK vertex
    // v leads through D interpolated vertices to the j-th chosen neighbour (NNFrechet.cpp:377-394)
    for (uint32_t i = 0; i < res->n_ik; ++i) {
      const std::vector<std::size_t> &out = nn.m_vertices[2 + i].out;
      for (int j = 0; j < k; ++j) {
        uint32_t nb = 0xFFFFFFFFu;
        if (j < (int)out.size()) {
          std::size_t v = nn.m_edges[out[j]].dst;
          if (v != planner.mNNGoalNode) {
            for (int s = 0; s < D; ++s)
              v = nn.m_edges[nn.m_vertices[v].out.at(0)].dst;
            nb = (uint32_t)v;
          }
        }
        knn_out[(size_t)i * k + j] = nb;
      }
    }

    t0 = std::chrono::steady_clock::now();
    ob::PlannerStatus st = ob::PlannerStatus::TIMEOUT;
    if (search_only) {
      std::vector<Vertex> p = planner.mLPAStar->computeShortestPath(planner.mTensorProductGraph);
      if (!p.empty())
        planner.extractNNPath(p);
    } else {
      st = planner.solve(solve_seconds);
    }
    res->solve_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    res->solved = (!search_only && st == ob::PlannerStatus::EXACT_SOLUTION) ? 1 : 0;
    res->final_error = planner.mFinalError;
    res->goal_cost = planner.mLPAStar->mDistanceMap[planner.mTensorGoalNode];

    // per NN edge: evaluated marker (NNFrechet.cpp:658-660) and whether its mapped TPG edges were set
    // to mInfVal (NNFrechet.cpp:692-693)
    for (uint32_t e = 0; e < res->n_edges; ++e) {
      const auto &rec = nn.m_edges[e];
      uint8_t flags = rec.prop.evaluated ? 1 : 0;
      const std::string key = nn.m_vertices[rec.src].prop.name + "_" + nn.m_vertices[rec.dst].prop.name;
      auto it = planner.mNNToTPGEdges.find(key);
      if (it != planner.mNNToTPGEdges.end() && !it->second.empty() &&
          tpg.m_edges[it->second.front().id].prop.length == DBL_MAX)
        flags |= 2;
      if (flags & 1)
        res->n_evaluated++;
      if (flags & 2)
        res->n_collisions++;
      if (e < cap_edges) {
        nn_edges[2 * (size_t)e] = (uint32_t)rec.src;
        nn_edges[2 * (size_t)e + 1] = (uint32_t)rec.dst;
        nn_edge_flags[e] = flags;
      }
    }
    if (res->solved && pdef->hasSolution()) {
      const og::PathGeometric *p = pdef->getSolutionPath()->as<og::PathGeometric>();
      res->path_len = (uint32_t)p->getStateCount();
      for (uint32_t i = 0; i < res->path_len; ++i) {
        if (i >= cap_path) {
          res->overflow = 1;
          break;
        }
        std::memcpy(path_states + (size_t)i * dof, values(p->getState(i)), dof * sizeof(double));
      }
    }
    return 0;
  } catch (std::exception &e) {
    std::fprintf(stderr, "ref_nnf_run: %s\n", e.what());
    return -1;
  }
}

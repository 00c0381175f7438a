// pr_gpu::NNFrechet: host side of the NNF planner over prgpu_nnf_* (see pr_gpu/NNFrechet.h).
// Behaviour follows frechet/src/NNFrechet.cpp (setters :23-84, setup :184-221, solve :95-178,
// subsampleRefPath :235-252 with util.cpp's linIntSpace :3-21, the IK budget :310-343 and :418-448).
#include "pr_gpu/NNFrechet.h"
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstring>
#include <ompl/base/spaces/RealVectorStateSpace.h>

namespace ob = ompl::base;
namespace pr_gpu {

NNFrechet::NNFrechet(const ob::SpaceInformationPtr &si) : ob::Planner(si, "NNFrechet"), space_(si->getStateSpace()) {
  rng_.seed(1);
}
NNFrechet::NNFrechet(const ob::SpaceInformationPtr &si, int numWaypoints, int ikMultiplier, int numNN,
                     int discretization, int seed)
    : ob::Planner(si, "NNFrechet"), space_(si->getStateSpace()), numWaypoints_(numWaypoints),
      ikMultiplier_(ikMultiplier), numNN_(numNN), discretization_(discretization) {
  rng_.seed(seed);
}
NNFrechet::~NNFrechet() { releaseGraph(); }

void NNFrechet::releaseGraph() {
  prgpu_nnf_destroy(handle_);
  handle_ = nullptr;
  for (ob::State *s : ikStates_)
    si_->freeState(s);
  ikStates_.clear();
  layerCounts_.clear();
  pathVertices_.clear();
  vertexStates_.clear();
  finalError_ = goalCost_ = 0.0;
  lazyIters_ = nEvaluated_ = nCollisions_ = 0;
}

void NNFrechet::clear() {
  ob::Planner::clear();
  releaseGraph();
  if (!originalRefPath_.empty())
    refPath_ = originalRefPath_; // setup() sub-samples in place (as the reference does, NNFrechet.cpp:201)
  setup_ = false;
}

void NNFrechet::setRefPath(std::vector<Eigen::Isometry3d> &referencePath) {
  if (referencePath.size() <= 1)
    throw ompl::Exception("referencePath is empty!");
  refPath_ = referencePath;
  originalRefPath_.clear();
}
void NNFrechet::setFKFunc(std::function<Eigen::Isometry3d(ob::State *)> fkFunc) {
  if (!fkFunc)
    throw ompl::Exception("Passed FK function is NULL!");
  fk_ = fkFunc;
}
void NNFrechet::setIKFunc(std::function<std::vector<ob::State *>(Eigen::Isometry3d &, int)> ikFunc) {
  if (!ikFunc)
    throw ompl::Exception("Passed IK function is NULL!");
  ik_ = ikFunc;
}
void NNFrechet::setDistanceFunc(std::function<double(Eigen::Isometry3d &, Eigen::Isometry3d &)> distanceFunc) {
  if (!distanceFunc)
    throw ompl::Exception("Passed task-space distance function is NULL!");
  dist_ = distanceFunc;
}
void NNFrechet::setNumWaypoints(int v) {
  if (v <= 0)
    throw ompl::Exception("numWaypoints must be positive!");
  numWaypoints_ = v;
}
void NNFrechet::setIKMultiplier(int v) {
  if (v <= 0)
    throw ompl::Exception("ikMultiplier must be positive!");
  ikMultiplier_ = v;
}
void NNFrechet::setNumNN(int v) {
  if (v <= 0)
    throw ompl::Exception("numNN must be positive!");
  numNN_ = v;
}
void NNFrechet::setDiscretization(int v) {
  if (v <= 0)
    throw ompl::Exception("discretization must be positive!");
  discretization_ = v;
}
void NNFrechet::setRandomSeed(int seed) { rng_.seed(seed); }

void NNFrechet::setProblemDefinition(const ob::ProblemDefinitionPtr &pdef) {
  ob::Planner::setProblemDefinition(pdef); // start/goal are not used by this planner
}

std::vector<Eigen::Isometry3d> NNFrechet::subsampleRefPath(const std::vector<Eigen::Isometry3d> &path) const {
  const int numPts = ((numWaypoints_ - 1) * (discretization_ + 1)) + 1;
  if (numPts > (int)path.size())
    throw ompl::Exception("Can't subsample ref path!");
  // linIntSpace(0, size-1, numPts): truncation of min + i*delta (may stop short of the last pose)
  const double mn = 0.0, mx = (double)(path.size() - 1);
  const double delta = (mx - mn) / double(numPts - 1);
  std::vector<Eigen::Isometry3d> out;
  for (int i = 0; i < numPts; i++)
    out.push_back(path[(int)(mn + i * delta)]);
  return out;
}

void NNFrechet::setup() {
  if (refPath_.size() == 0)
    throw ompl::Exception("Ref path not set!");
  if (!fk_)
    throw ompl::Exception("FK function not set!");
  if (!ik_)
    throw ompl::Exception("IK function not set!");
  if (!dist_)
    throw ompl::Exception("Task-space distance function not set!");
  requireRealVectorSpace(space_.get(), "NNFrechet");
  const auto t0 = std::chrono::steady_clock::now();
  const int dof = (int)space_->getDimension();
  // setup() may be called again (after clear() or a parameter change): start from nothing
  releaseGraph();
  if (world_ && world_->dof() != dof)
    throw ompl::Exception("NNFrechet", "state dimension differs from the world's dof");

  if (originalRefPath_.empty())
    originalRefPath_ = refPath_;
  refPath_ = subsampleRefPath(originalRefPath_);
  const int R = (int)refPath_.size();

  // IK budget: both end waypoints get ikMultiplier solutions, the rest of numWaypoints*ikMultiplier
  // is spread over the interior indices by uniform draws (a fresh distribution object per draw)
  std::vector<int> want(R, 0);
  want[0] = ikMultiplier_;
  want[R - 1] = ikMultiplier_;
  const int budget = numWaypoints_ * ikMultiplier_ - 2 * ikMultiplier_;
  for (int i = 0; i < budget; i++) {
    std::uniform_int_distribution<int> uniform(1, R - 2);
    want[uniform(rng_)]++;
  }
  // IK callback in the reference's call order: first waypoint, last waypoint, then 1..R-2
  layerCounts_.assign(R, 0);
  std::vector<double> ik;
  auto sample = [&](int r) {
    std::vector<ob::State *> sols = ik_(refPath_[r], want[r]);
    for (ob::State *s : sols) {
      const double *v = space_->getValueAddressAtIndex(s, 0);
      ik.insert(ik.end(), v, v + dof);
      ikStates_.push_back(s);
    }
    layerCounts_[r] = (unsigned)sols.size();
  };
  sample(0);
  sample(R - 1);
  for (int r = 1; r <= R - 2; ++r)
    sample(r);

  std::vector<double> ref12((size_t)R * 12);
  for (int r = 0; r < R; ++r)
    std::memcpy(ref12.data() + (size_t)r * 12, refPath_[r].data(), 12 * sizeof(double));

  prgpu_nnf_params p;
  p.num_nn = numNN_;
  p.discretization = discretization_;
  p.check_resolution = checkResolution_;
  // Device form: FK, the translation metric and edge validation all run on the device model -- taken only
  // when a world was given AND the user's callbacks agree with it (checked on every IK solution of the first
  // and the last layer).  Otherwise the host-callback form: the callbacks are what the reference calls.
  hostForm_ = !world_ || !callbacksMatchWorld();
  uint32_t nv = 0;
  if (!hostForm_) {
    check(prgpu_nnf_create(world_->handle(), &p, R, ref12.data(), layerCounts_.data(), ik.data(), &handle_),
          "NNFrechet::setup");
    check(prgpu_nnf_graph_info(handle_, &nv, nullptr, nullptr), "NNFrechet::setup");
    vertexStates_.resize((size_t)nv * dof);
    check(prgpu_nnf_get_vertices(handle_, vertexStates_.data(), nullptr), "NNFrechet::setup");
  } else {
    check(prgpu_nnf_create_graph(world_ ? world_->device() : device_, dof, &p, R, layerCounts_.data(), ik.data(),
                                 &handle_),
          "NNFrechet::setup");
    check(prgpu_nnf_graph_info(handle_, &nv, nullptr, nullptr), "NNFrechet::setup");
    vertexStates_.resize((size_t)nv * dof);
    check(prgpu_nnf_get_vertices(handle_, vertexStates_.data(), nullptr), "NNFrechet::setup");
    // node weights through the user's FK (NNFrechet.cpp:295, :372) and metric (:495); the two dummy
    // vertices carry the first / last reference pose (:409, :414)
    std::vector<double> weights((size_t)nv * R);
    ob::State *tmp = si_->allocState();
    for (uint32_t v = 0; v < nv; ++v) {
      Eigen::Isometry3d pose;
      if (v == 0)
        pose = refPath_.front();
      else if (v == 1)
        pose = refPath_.back();
      else {
        std::memcpy(space_->getValueAddressAtIndex(tmp, 0), vertexStates_.data() + (size_t)v * dof, dof * sizeof(double));
        pose = fk_(tmp);
      }
      for (int r = 0; r < R; ++r)
        weights[(size_t)v * R + r] = dist_(refPath_[r], pose);
    }
    si_->freeState(tmp);
    check(prgpu_nnf_set_weights(handle_, weights.data()), "NNFrechet::setup");
    check(prgpu_nnf_set_edge_evaluator(handle_, &NNFrechet::evaluateEdgesThunk, this), "NNFrechet::setup");
  }
  OMPL_INFORM("Tensor Product Graph has been built.");
  buildTime_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  setup_ = true;
}

// Do the user's FK and metric describe the device world (its arm model, the translation metric)?
bool NNFrechet::callbacksMatchWorld() {
  const int dof = world_->dof();
  const size_t nProbe = std::min<size_t>(ikStates_.size(), (size_t)layerCounts_.front() + layerCounts_.back());
  for (size_t i = 0; i < nProbe; ++i) {
    double pose[12];
    check(prgpu_fk_batch(world_->handle(), space_->getValueAddressAtIndex(ikStates_[i], 0), 1, pose), "NNFrechet::setup");
    (void)dof;
    Eigen::Isometry3d userPose = fk_(ikStates_[i]);
    double d2 = 0.0;
    for (int r = 0; r < 3; ++r) {
      const double diff = userPose(r, 3) - pose[4 * r + 3];
      d2 += diff * diff;
    }
    if (std::sqrt(d2) > 1e-9)
      return false;
    Eigen::Isometry3d a = refPath_[i % refPath_.size()], b = userPose;
    double t2 = 0.0;
    for (int r = 0; r < 3; ++r)
      t2 += (a(r, 3) - b(r, 3)) * (a(r, 3) - b(r, 3));
    if (std::fabs(dist_(a, b) - std::sqrt(t2)) > 1e-9 * (1.0 + std::sqrt(t2)))
      return false;
  }
  return nProbe > 0;
}

// evaluateEdge (NNFrechet.cpp:655-683) over the caller's validity checker: numQ = ceil(length / resolution)
// states at an alpha that ACCUMULATES 1/numQ (so the far endpoint may or may not be reached), first invalid
// state decides.
int NNFrechet::evaluateEdgesThunk(void *user, const double *from, const double *to, uint32_t n, uint8_t *freeOut) {
  NNFrechet *self = static_cast<NNFrechet *>(user);
  try {
    const int dof = (int)self->space_->getDimension();
    const ob::StateValidityCheckerPtr &checker = self->si_->getStateValidityChecker();
    ob::State *a = self->si_->allocState(), *b = self->si_->allocState(), *x = self->si_->allocState();
    for (uint32_t i = 0; i < n; ++i) {
      std::memcpy(self->space_->getValueAddressAtIndex(a, 0), from + (size_t)i * dof, dof * sizeof(double));
      std::memcpy(self->space_->getValueAddressAtIndex(b, 0), to + (size_t)i * dof, dof * sizeof(double));
      const int numQ = (int)std::ceil(self->space_->distance(a, b) / self->checkResolution_);
      const double step = 1.0 / numQ;
      bool free = true;
      for (double alpha = 0.0; alpha <= 1.0 && free; alpha += step) {
        self->space_->interpolate(a, b, alpha, x);
        free = checker->isValid(x);
      }
      freeOut[i] = free ? 1 : 0;
    }
    self->si_->freeState(a);
    self->si_->freeState(b);
    self->si_->freeState(x);
    return 0;
  } catch (...) {
    return 1;
  }
}

ob::PlannerStatus NNFrechet::solve(const ob::PlannerTerminationCondition &ptc) {
  if (!handle_)
    throw ompl::Exception("NNFrechet", "setup() has not been called");
  const auto t0 = std::chrono::steady_clock::now();
  bool solutionFound = false;
  std::vector<uint32_t> path(1u << 16);
  while (ptc == false) {
    prgpu_nnf_result res;
    check(prgpu_nnf_solve(handle_, 8, &res, path.data(), (uint32_t)path.size(), nullptr), "NNFrechet::solve");
    lazyIters_ += res.lazy_iters;
    nEvaluated_ = res.n_evaluated;
    nCollisions_ = res.n_collisions;
    goalCost_ = res.goal_cost;
    if (res.lazy_iters)
      finalError_ = res.final_error;
    if (res.solved) {
      solutionFound = true;
      pathVertices_.assign(path.begin(), path.begin() + res.path_len);
      break;
    }
    if (res.goal_cost >= DBL_MAX) // every path is in collision
      break;
  }
  searchTime_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (solutionFound) {
    const int dof = (int)space_->getDimension();
    auto *out = new ompl::geometric::PathGeometric(si_);
    ob::State *tmp = si_->allocState();
    for (unsigned v : pathVertices_) {
      std::memcpy(space_->getValueAddressAtIndex(tmp, 0), vertexStates_.data() + (size_t)v * dof, dof * sizeof(double));
      out->append(tmp);
    }
    si_->freeState(tmp);
    pdef_->addSolutionPath(ob::PathPtr(out));
    OMPL_INFORM("Solution Found!");
    OMPL_INFORM("Final Frechet Error:    %f", finalError_);
    OMPL_INFORM("Length of path:    %d", (int)pathVertices_.size());
    OMPL_INFORM("Total time to init structures:    %f seconds", buildTime_);
    OMPL_INFORM("Total time to search TPG:    %f seconds", searchTime_);
    return ob::PlannerStatus::EXACT_SOLUTION;
  }
  OMPL_INFORM("Solution NOT Found");
  return ob::PlannerStatus::TIMEOUT;
}

ob::PlannerStatus NNFrechet::solve(double solveTime) {
  return solve(ob::timedPlannerTerminationCondition(solveTime));
}

} // namespace pr_gpu

// pr_gpu::NNFrechet -- drop-in for NNFrechet::NNFrechet (frechet/include/NNFrechet.hpp:13): an
// ompl::base::Planner named "NNFrechet" with the reference's constructors, setters
// (setRefPath / setFKFunc / setIKFunc / setDistanceFunc / setNumWaypoints / setIKMultiplier /
// setNumNN / setDiscretization / setRandomSeed), the same ompl::Exception messages, setup(),
// solve(ptc) and solve(double).  The host keeps what the reference obtains through callbacks
// (reference-path sub-sampling, the IK budget per waypoint drawn exactly as
// frechet/src/NNFrechet.cpp:310-343, the IK callback itself); everything else -- kNN, sub-sampled
// edges, FK, tensor-product search, edge evaluation -- runs through prgpu_nnf_*.
// Two forms, chosen in setup():
//  * device form -- setWorld(GpuWorldPtr) was called and the user's FK / distance callbacks agree with the
//    device model (checked on the IK solutions of the end layers): FK, the translation metric and
//    evaluateEdge all run on the device;
//  * host-callback form -- no world, or callbacks the device has no model of (another robot, a metric
//    with a rotation term ...): the user's FK and metric fill the node-weight table on the host exactly as
//    the reference calls them (NNFrechet.cpp:295,372,495), evaluateEdge runs over
//    si_->getStateValidityChecker(), and the device does findKnnNodes, the NN graph and every bottleneck search.
#ifndef PR_GPU_NNFRECHET_H
#define PR_GPU_NNFRECHET_H
#include <functional>
#include <random>
#include <vector>
#include <Eigen/Dense>
#include <ompl/base/Planner.h>
#include <ompl/geometric/PathGeometric.h>
#include "pr_gpu/GpuWorld.h"
namespace pr_gpu {
class NNFrechet : public ompl::base::Planner {
public:
  explicit NNFrechet(const ompl::base::SpaceInformationPtr &si);
  NNFrechet(const ompl::base::SpaceInformationPtr &si, int numWaypoints, int ikMultiplier, int numNN,
            int discretization, int seed = 1);
  ~NNFrechet() override;

  void setRefPath(std::vector<Eigen::Isometry3d> &referencePath);
  void setFKFunc(std::function<Eigen::Isometry3d(ompl::base::State *)> fkFunc);
  void setIKFunc(std::function<std::vector<ompl::base::State *>(Eigen::Isometry3d &, int)> ikFunc);
  void setDistanceFunc(std::function<double(Eigen::Isometry3d &, Eigen::Isometry3d &)> distanceFunc);
  void setNumWaypoints(int numWaypoints);
  void setIKMultiplier(int ikMultiplier);
  void setNumNN(int numNN);
  void setDiscretization(int discretization);
  void setRandomSeed(int seed);
  /// device model used for FK and collision checking (optional; see the two forms above)
  void setWorld(const GpuWorldPtr &world) { world_ = world; }
  /// CUDA device of the host-callback form when no world is given (default 0)
  void setDevice(int device) { device_ = device; }
  bool usesHostCallbacks() const { return hostForm_; }

  void setProblemDefinition(const ompl::base::ProblemDefinitionPtr &pdef) override;
  ompl::base::PlannerStatus solve(const ompl::base::PlannerTerminationCondition &ptc) override;
  ompl::base::PlannerStatus solve(double solveTime);
  void setup() override;
  void clear() override;

  // statistics the reference only logs (frechet/src/NNFrechet.cpp:162-171)
  double finalError() const { return finalError_; }
  double goalCost() const { return goalCost_; }
  unsigned lazyIterations() const { return lazyIters_; }
  unsigned evaluatedEdges() const { return nEvaluated_; }
  unsigned collidingEdges() const { return nCollisions_; }
  const std::vector<unsigned> &layerCounts() const { return layerCounts_; }
  const std::vector<unsigned> &pathVertices() const { return pathVertices_; }

private:
  void releaseGraph();
  bool callbacksMatchWorld();
  static int evaluateEdgesThunk(void *user, const double *from, const double *to, uint32_t n, uint8_t *freeOut);
  bool hostForm_ = false;
  int device_ = 0;
  std::vector<Eigen::Isometry3d> subsampleRefPath(const std::vector<Eigen::Isometry3d> &path) const;

  ompl::base::StateSpacePtr space_;
  GpuWorldPtr world_;
  prgpu_nnf_t *handle_ = nullptr;
  int numWaypoints_ = 5, ikMultiplier_ = 5, numNN_ = 5, discretization_ = 3; // NNFrechet.hpp:262-274
  double checkResolution_ = 0.05;                                             // NNFrechet.hpp:277
  std::default_random_engine rng_;
  std::vector<Eigen::Isometry3d> refPath_, originalRefPath_;
  std::function<Eigen::Isometry3d(ompl::base::State *)> fk_;
  std::function<std::vector<ompl::base::State *>(Eigen::Isometry3d &, int)> ik_;
  std::function<double(Eigen::Isometry3d &, Eigen::Isometry3d &)> dist_;
  std::vector<ompl::base::State *> ikStates_; // owned (the reference never frees them)
  std::vector<unsigned> layerCounts_, pathVertices_;
  std::vector<double> vertexStates_;
  double finalError_ = 0.0, goalCost_ = 0.0, buildTime_ = 0.0, searchTime_ = 0.0;
  unsigned lazyIters_ = 0, nEvaluated_ = 0, nCollisions_ = 0;
};
} // namespace pr_gpu
#endif

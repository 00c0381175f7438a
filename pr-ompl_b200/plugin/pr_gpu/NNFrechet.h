// pr_gpu::NNFrechet -- drop-in for NNFrechet::NNFrechet (frechet/include/NNFrechet.hpp:13): an
// ompl::base::Planner named "NNFrechet" with the reference's constructors, setters
// (setRefPath / setFKFunc / setIKFunc / setDistanceFunc / setNumWaypoints / setIKMultiplier /
// setNumNN / setDiscretization / setRandomSeed), the same ompl::Exception messages, setup(),
// solve(ptc) and solve(double).  The host keeps what the reference obtains through callbacks
// (reference-path sub-sampling, the IK budget per waypoint drawn exactly as
// frechet/src/NNFrechet.cpp:310-343, the IK callback itself); everything else -- kNN, sub-sampled
// edges, FK, tensor-product search, edge evaluation -- runs through prgpu_nnf_*.
// The device needs a model, so one extra call is required: setWorld(GpuWorldPtr).  The FK and
// distance callbacks must still be supplied (as in the reference) and are used to verify that
// the device model agrees with them (FK of the first IK solution; translation metric).
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
  /// device model used for FK and collision checking (required before setup())
  void setWorld(const GpuWorldPtr &world) { world_ = world; }

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

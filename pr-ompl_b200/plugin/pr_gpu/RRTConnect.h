// pr_gpu::RRTConnect -- drop-in for pr_ompl::RRTConnect (pr_ompl/include/pr_ompl/RRTConnect.h:65):
// an ompl::base::Planner named "RRTConnect" with the same public surface (setRange/getRange,
// setExtensionTypes/getExtensionTypes, setNearestNeighbors<NN>(), params "range" and
// "extension_type" with values con-con / con-ext / ext-con / ext-ext, setup/solve/clear/
// getPlannerData, the same PlannerStatus results and the same std::runtime_error on a bad
// extension string), whose nearest-neighbour structure defaults to GpuNearestNeighbors and
// which recognises GpuMotionValidator to validate goal-tree edges in one launch.
// <pr_ompl/RRTConnect.h> in this directory aliases it into namespace pr_ompl.
#ifndef PR_GPU_RRTCONNECT_H
#define PR_GPU_RRTCONNECT_H
#include <deque>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include <ompl/datastructures/NearestNeighbors.h>
#include <ompl/geometric/planners/PlannerIncludes.h>
#include "pr_gpu/GpuNearestNeighbors.h"
namespace pr_gpu {
class RRTConnect : public ompl::base::Planner {
public:
  enum ExtensionType { EXTTYPE_EXTEND, EXTTYPE_CONNECT };
  typedef std::pair<ExtensionType, ExtensionType> ExtensionTypes;

  explicit RRTConnect(const ompl::base::SpaceInformationPtr &si);
  ~RRTConnect() override;

  void getPlannerData(ompl::base::PlannerData &data) const override;
  ompl::base::PlannerStatus solve(const ompl::base::PlannerTerminationCondition &ptc) override;
  using ompl::base::Planner::solve;
  void clear() override;
  void setup() override;

  void setRange(double distance) { range_ = distance; }
  double getRange() const { return range_; }
  void setExtensionTypes(const ExtensionTypes &types) { ext_ = types; }
  ExtensionTypes getExtensionTypes() const { return ext_; }

  /// use a different nearest-neighbour structure (must be default constructible)
  template <template <typename T> class NN> void setNearestNeighbors() {
    trees_[0].nn.reset(new NN<Node *>());
    trees_[1].nn.reset(new NN<Node *>());
  }

  /// tree nodes in insertion order: state values and parent index (-1 for roots); side 0 = start
  std::size_t treeSize(int side) const { return trees_[side].nodes.size(); }
  const ompl::base::State *treeState(int side, std::size_t i) const { return trees_[side].nodes[i].state; }
  int treeParent(int side, std::size_t i) const {
    const Node *p = trees_[side].nodes[i].parent;
    return p ? (int)p->index : -1;
  }

protected:
  struct Node {
    ompl::base::State *state = nullptr;
    const double *coords = nullptr; // contiguous reals of `state` (GPU coordinate accessor)
    Node *parent = nullptr;
    const ompl::base::State *root = nullptr;
    std::size_t index = 0;
  };
  struct Tree {
    std::shared_ptr<ompl::NearestNeighbors<Node *>> nn;
    std::deque<Node> nodes; // stable addresses, insertion order
  };
  enum Growth { TRAPPED, ADVANCED, REACHED };

  void setExtensionTypesString(const std::string &str);
  std::string getExtensionTypesString() const;
  static const char *extensionName(ExtensionType t);

  Node *insert(Tree &tree, const ompl::base::State *state, Node *parent);
  Growth extend(Tree &tree, bool fromStart, ompl::base::State *target, ompl::base::State *scratch,
                Node *&added);
  void release();

  ompl::base::StateSamplerPtr sampler_;
  Tree trees_[2]; // [0] rooted at the start states, [1] at the goal states
  double range_;
  ExtensionTypes ext_;
  std::pair<ompl::base::State *, ompl::base::State *> connection_;
};
} // namespace pr_gpu
#endif

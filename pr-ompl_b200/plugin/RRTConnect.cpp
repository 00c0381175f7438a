// pr_gpu::RRTConnect: host control flow of the bidirectional planner over the OMPL interfaces.
// Behaviour follows pr_ompl/src/RRTConnect.cpp (solve :219-376, growTree :178-217, setup
// :125-137, clear :166-176, getPlannerData :378-415, extension strings :78-123); the code is
// organised around index-stable node arrays instead of heap-allocated Motion objects.
#include "pr_gpu/RRTConnect.h"
#include <sstream>
#include <stdexcept>
#include <ompl/tools/config/SelfConfig.h>
#include "pr_gpu/GpuWorld.h"

namespace ob = ompl::base;
namespace og = ompl::geometric;

namespace pr_gpu {

RRTConnect::RRTConnect(const ob::SpaceInformationPtr &si)
    : ob::Planner(si, "RRTConnect"), range_(0.0), ext_(EXTTYPE_EXTEND, EXTTYPE_CONNECT),
      connection_(nullptr, nullptr) {
  specs_.recognizedGoal = ob::GOAL_SAMPLEABLE_REGION;
  specs_.directed = true;
  declareParam<double>("range", this, &RRTConnect::setRange, &RRTConnect::getRange, "0.:1.:10000.");
  declareParam<std::string>("extension_type", this, &RRTConnect::setExtensionTypesString,
                            &RRTConnect::getExtensionTypesString, "con-con,con-ext,ext-con,ext-ext");
}

RRTConnect::~RRTConnect() { release(); }

const char *RRTConnect::extensionName(ExtensionType t) {
  switch (t) {
  case EXTTYPE_CONNECT:
    return "con";
  case EXTTYPE_EXTEND:
    return "ext";
  }
  throw std::runtime_error("Unknown extension type.");
}

void RRTConnect::setExtensionTypesString(const std::string &str) {
  static const struct {
    const char *name;
    ExtensionType first, second;
  } table[] = {{"con-con", EXTTYPE_CONNECT, EXTTYPE_CONNECT},
               {"con-ext", EXTTYPE_CONNECT, EXTTYPE_EXTEND},
               {"ext-con", EXTTYPE_EXTEND, EXTTYPE_CONNECT},
               {"ext-ext", EXTTYPE_EXTEND, EXTTYPE_EXTEND}};
  for (const auto &row : table)
    if (str == row.name) {
      ext_ = ExtensionTypes(row.first, row.second);
      return;
    }
  throw std::runtime_error("Invalid extension type string.");
}

std::string RRTConnect::getExtensionTypesString() const {
  return std::string(extensionName(ext_.first)) + "-" + extensionName(ext_.second);
}

void RRTConnect::setup() {
  ob::Planner::setup();
  // nearest() is Euclidean over `dim` contiguous doubles on the device while the step clamp uses
  // si_->distance / interpolate: the two agree for RealVectorStateSpace only
  requireRealVectorSpace(si_->getStateSpace().get(), "RRTConnect");
  ompl::tools::SelfConfig(si_, getName()).configurePlannerRange(range_);
  const int dim = (int)si_->getStateDimension();
  for (Tree &t : trees_) {
    if (!t.nn)
      t.nn.reset(new GpuNearestNeighbors<Node *>());
    t.nn->setDistanceFunction(
        [this](Node *const &a, Node *const &b) { return si_->distance(a->state, b->state); });
    if (auto *gpu = dynamic_cast<GpuNearestNeighbors<Node *> *>(t.nn.get()))
      gpu->setCoordinateAccessor([](Node *const &n) { return n->coords; }, dim);
  }
}

void RRTConnect::release() {
  for (Tree &t : trees_) {
    for (Node &n : t.nodes)
      if (n.state)
        si_->freeState(n.state);
    t.nodes.clear();
  }
}

void RRTConnect::clear() {
  ob::Planner::clear();
  sampler_.reset();
  release();
  for (Tree &t : trees_)
    if (t.nn)
      t.nn->clear();
  connection_ = std::make_pair<ob::State *, ob::State *>(nullptr, nullptr);
}

RRTConnect::Node *RRTConnect::insert(Tree &tree, const ob::State *state, Node *parent) {
  tree.nodes.emplace_back();
  Node &n = tree.nodes.back();
  n.state = si_->cloneState(state);
  n.coords = si_->getStateSpace()->getValueAddressAtIndex(n.state, 0);
  n.parent = parent;
  n.root = parent ? parent->root : n.state;
  n.index = tree.nodes.size() - 1;
  tree.nn->add(&n);
  return &n;
}

// One step of a tree toward `target`.
RRTConnect::Growth RRTConnect::extend(Tree &tree, bool fromStart, ob::State *target, ob::State *scratch,
                                      Node *&added) {
  Node probe;
  probe.state = target;
  probe.coords = si_->getStateSpace()->getValueAddressAtIndex(target, 0);
  Node *near = tree.nn->nearest(&probe);

  ob::State *goalOfStep = target;
  const double d = si_->distance(near->state, target);
  const bool reaches = !(d > range_);
  if (!reaches) {
    si_->getStateSpace()->interpolate(near->state, target, range_ / d, scratch);
    goalOfStep = scratch;
  }
  // The start tree validates tree -> new state.  The goal tree validates new state -> tree, and
  // because checkMotion takes its first argument on trust, the new state is tested first.
  bool ok;
  if (fromStart)
    ok = si_->checkMotion(near->state, goalOfStep);
  else if (auto *gmv = dynamic_cast<const GpuMotionValidator *>(si_->getMotionValidator().get()))
    ok = gmv->checkMotionWithFirst(goalOfStep, near->state);
  else
    ok = si_->getStateValidityChecker()->isValid(goalOfStep) && si_->checkMotion(goalOfStep, near->state);
  if (!ok)
    return TRAPPED;
  added = insert(tree, goalOfStep, near);
  return reaches ? REACHED : ADVANCED;
}

ob::PlannerStatus RRTConnect::solve(const ob::PlannerTerminationCondition &ptc) {
  checkValidity();
  auto *goal = dynamic_cast<ob::GoalSampleableRegion *>(pdef_->getGoal().get());
  if (!goal) {
    OMPL_ERROR("Unknown type of goal (or goal undefined)");
    return ob::PlannerStatus::UNRECOGNIZED_GOAL_TYPE;
  }
  Tree &startTree = trees_[0], &goalTree = trees_[1];
  while (const ob::State *st = pis_.nextStart())
    insert(startTree, st, nullptr);
  if (startTree.nn->size() == 0) {
    OMPL_ERROR("Motion planning start tree could not be initialized!");
    return ob::PlannerStatus::INVALID_START;
  }
  if (!goal->couldSample()) {
    OMPL_ERROR("Insufficient states in sampleable goal region");
    return ob::PlannerStatus::INVALID_GOAL;
  }
  if (!sampler_)
    sampler_ = si_->allocStateSampler();
  OMPL_INFORM("Starting with %d states", (int)(startTree.nn->size() + goalTree.nn->size()));

  ob::State *scratch = si_->allocState();
  ob::State *target = si_->allocState();
  bool solved = false;
  int side = 0; // tree grown first in this iteration; alternates every iteration

  while (ptc == false) {
    Tree &first = trees_[side];
    Tree &second = trees_[1 - side];
    const bool firstIsStart = (side == 0);
    side = 1 - side;

    // keep the goal tree seeded
    if (goalTree.nn->size() == 0 || pis_.getSampledGoalsCount() < goalTree.nn->size() / 2) {
      const ob::State *st = goalTree.nn->size() == 0 ? pis_.nextGoal(ptc) : pis_.nextGoal();
      if (st)
        insert(goalTree, st, nullptr);
      if (goalTree.nn->size() == 0) {
        OMPL_ERROR("Unable to sample any valid states for goal tree");
        break;
      }
    }

    sampler_->sampleUniform(target);

    // phase 1: grow `first` toward the sample (repeatedly when its extension type is CONNECT)
    Node *lastAdded = nullptr;
    Growth g;
    do
      g = extend(first, firstIsStart, target, scratch, lastAdded);
    while (g == ADVANCED && ext_.first == EXTTYPE_CONNECT);
    if (!lastAdded)
      continue;

    // phase 2: grow `second` toward the node phase 1 ended on
    if (g != REACHED)
      si_->copyState(target, lastAdded->state);
    Node *bridge = nullptr;
    do
      g = extend(second, !firstIsStart, target, scratch, bridge);
    while (g == ADVANCED && ext_.second == EXTTYPE_CONNECT);

    Node *onStart = firstIsStart ? lastAdded : bridge;
    Node *onGoal = firstIsStart ? bridge : lastAdded;
    if (g == REACHED && goal->isStartGoalPairValid(onStart->root, onGoal->root)) {
      // both nodes hold the same state: drop one copy from the path
      if (onStart->parent)
        onStart = onStart->parent;
      else
        onGoal = onGoal->parent;
      connection_ = std::make_pair(onStart->state, onGoal->state);

      std::vector<const Node *> toRoot;
      for (const Node *n = onStart; n; n = n->parent)
        toRoot.push_back(n);
      auto *path = new og::PathGeometric(si_);
      for (auto it = toRoot.rbegin(); it != toRoot.rend(); ++it)
        path->append((*it)->state);
      for (const Node *n = onGoal; n; n = n->parent)
        path->append(n->state);
      pdef_->addSolutionPath(ob::PathPtr(path), false, 0.0);
      solved = true;
      break;
    }
  }
  si_->freeState(scratch);
  si_->freeState(target);
  OMPL_INFORM("Created %u states (%u start + %u goal)", (unsigned)(startTree.nn->size() + goalTree.nn->size()),
              (unsigned)startTree.nn->size(), (unsigned)goalTree.nn->size());
  return solved ? ob::PlannerStatus::EXACT_SOLUTION : ob::PlannerStatus::TIMEOUT;
}

void RRTConnect::getPlannerData(ob::PlannerData &data) const {
  ob::Planner::getPlannerData(data);
  for (const Node &n : trees_[0].nodes) {
    if (!n.parent)
      data.addStartVertex(ob::PlannerDataVertex(n.state, 1));
    else
      data.addEdge(ob::PlannerDataVertex(n.parent->state, 1), ob::PlannerDataVertex(n.state, 1));
  }
  for (const Node &n : trees_[1].nodes) {
    if (!n.parent)
      data.addGoalVertex(ob::PlannerDataVertex(n.state, 2));
    else // goal-tree edges point toward the goal, consistent with the start tree
      data.addEdge(ob::PlannerDataVertex(n.state, 2), ob::PlannerDataVertex(n.parent->state, 2));
  }
  if (connection_.first && connection_.second) // the reference dereferences these unconditionally (:414)
    data.addEdge(data.vertexIndex(connection_.first), data.vertexIndex(connection_.second));
}

} // namespace pr_gpu

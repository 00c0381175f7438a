// ompl_min: ompl::base::ProblemDefinition.
#ifndef OMPL_MIN_BASE_PROBLEMDEFINITION_H
#define OMPL_MIN_BASE_PROBLEMDEFINITION_H
#include <algorithm>
#include <vector>
#include "ompl/base/Goal.h"
#include "ompl/base/Path.h"
#include "ompl/base/goals/GoalState.h"
namespace ompl {
namespace base {
struct PlannerSolution {
  PlannerSolution(const PathPtr &path, bool approximate = false, double difference = -1.0)
      : index_(-1), path_(path), length_(path ? path->length() : 0.0), approximate_(approximate),
        difference_(difference) {}
  bool operator<(const PlannerSolution &b) const {
    if (!approximate_ && b.approximate_)
      return true;
    if (approximate_ && !b.approximate_)
      return false;
    if (approximate_ && b.approximate_)
      return difference_ < b.difference_;
    return length_ < b.length_;
  }
  int index_;
  PathPtr path_;
  double length_;
  bool approximate_;
  double difference_;
};

class ProblemDefinition {
public:
  explicit ProblemDefinition(const SpaceInformationPtr &si) : si_(si) {}
  virtual ~ProblemDefinition() { clearStartStates(); }
  ProblemDefinition(const ProblemDefinition &) = delete;
  ProblemDefinition &operator=(const ProblemDefinition &) = delete;

  const SpaceInformationPtr &getSpaceInformation() const { return si_; }
  void addStartState(const State *state) { startStates_.push_back(si_->cloneState(state)); }
  void clearStartStates() {
    for (auto *s : startStates_)
      si_->freeState(s);
    startStates_.clear();
  }
  unsigned int getStartStateCount() const { return (unsigned int)startStates_.size(); }
  const State *getStartState(unsigned int index) const { return startStates_.at(index); }
  State *getStartState(unsigned int index) { return startStates_.at(index); }

  void setGoal(const GoalPtr &goal) { goal_ = goal; }
  void clearGoal() { goal_.reset(); }
  const GoalPtr &getGoal() const { return goal_; }
  void setGoalState(const State *goal,
                    double threshold = std::numeric_limits<double>::epsilon()) {
    clearGoal();
    GoalState *gs = new GoalState(si_);
    gs->setState(goal);
    gs->setThreshold(threshold);
    setGoal(GoalPtr(gs));
  }
  void setStartAndGoalStates(const State *start, const State *goal,
                             double threshold = std::numeric_limits<double>::epsilon()) {
    clearStartStates();
    addStartState(start);
    setGoalState(goal, threshold);
  }

  bool hasSolution() const { return !solutions_.empty(); }
  bool hasExactSolution() const { return hasSolution() && !hasApproximateSolution(); }
  bool hasApproximateSolution() const {
    return !solutions_.empty() && solutions_.front().approximate_;
  }
  double getSolutionDifference() const {
    return solutions_.empty() ? -1.0 : solutions_.front().difference_;
  }
  PathPtr getSolutionPath() const {
    return solutions_.empty() ? PathPtr() : solutions_.front().path_;
  }
  void addSolutionPath(const PathPtr &path, bool approximate = false,
                       double difference = -1.0) const {
    PlannerSolution sol(path, approximate, difference);
    sol.index_ = (int)solutions_.size();
    solutions_.push_back(sol);
    std::stable_sort(solutions_.begin(), solutions_.end());
  }
  std::size_t getSolutionCount() const { return solutions_.size(); }
  const std::vector<PlannerSolution> &getSolutions() const { return solutions_; }
  void clearSolutionPaths() const { solutions_.clear(); }

protected:
  SpaceInformationPtr si_;
  std::vector<State *> startStates_;
  GoalPtr goal_;
  mutable std::vector<PlannerSolution> solutions_;
};
typedef std::shared_ptr<ProblemDefinition> ProblemDefinitionPtr;
} // namespace base
} // namespace ompl
#endif

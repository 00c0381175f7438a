// ompl_min: ompl::base::GoalStates (finite set of goal configurations, sampled round-robin).
#ifndef OMPL_MIN_BASE_GOALS_GOALSTATES_H
#define OMPL_MIN_BASE_GOALS_GOALSTATES_H
#include <vector>
#include "ompl/base/goals/GoalSampleableRegion.h"
namespace ompl {
namespace base {
class GoalStates : public GoalSampleableRegion {
public:
  explicit GoalStates(const SpaceInformationPtr &si) : GoalSampleableRegion(si), samplePosition_(0) {
    type_ = GOAL_STATES;
  }
  ~GoalStates() override { clear(); }
  void sampleGoal(State *st) const override {
    if (states_.empty())
      throw Exception("There are no goals to sample");
    si_->copyState(st, states_[samplePosition_]);
    samplePosition_ = (samplePosition_ + 1) % states_.size();
  }
  unsigned int maxSampleCount() const override { return (unsigned int)states_.size(); }
  double distanceGoal(const State *st) const override {
    double dist = std::numeric_limits<double>::infinity();
    for (std::size_t i = 0; i < states_.size(); ++i) {
      double d = si_->distance(st, states_[i]);
      if (d < dist)
        dist = d;
    }
    return dist;
  }
  virtual void addState(const State *st) { states_.push_back(si_->cloneState(st)); }
  virtual void clear() {
    for (std::size_t i = 0; i < states_.size(); ++i)
      si_->freeState(states_[i]);
    states_.clear();
    samplePosition_ = 0;
  }
  virtual bool hasStates() const { return !states_.empty(); }
  virtual const State *getState(unsigned int index) const { return states_.at(index); }
  virtual std::size_t getStateCount() const { return states_.size(); }

protected:
  std::vector<State *> states_;

private:
  mutable unsigned int samplePosition_;
};
} // namespace base
} // namespace ompl
#endif

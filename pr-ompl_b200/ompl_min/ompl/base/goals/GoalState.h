// ompl_min: ompl::base::GoalState (single goal configuration; maxSampleCount()==1).
#ifndef OMPL_MIN_BASE_GOALS_GOALSTATE_H
#define OMPL_MIN_BASE_GOALS_GOALSTATE_H
#include "ompl/base/goals/GoalSampleableRegion.h"
namespace ompl {
namespace base {
class GoalState : public GoalSampleableRegion {
public:
  explicit GoalState(const SpaceInformationPtr &si) : GoalSampleableRegion(si), state_(nullptr) {
    type_ = GOAL_STATE;
  }
  ~GoalState() override {
    if (state_)
      si_->freeState(state_);
  }
  void sampleGoal(State *st) const override { si_->copyState(st, state_); }
  unsigned int maxSampleCount() const override { return 1; }
  double distanceGoal(const State *st) const override { return si_->distance(st, state_); }
  void setState(const State *st) {
    if (state_)
      si_->freeState(state_);
    state_ = si_->cloneState(st);
  }
  const State *getState() const { return state_; }
  State *getState() { return state_; }

protected:
  State *state_;
};
} // namespace base
} // namespace ompl
#endif

// ompl_min: ompl::base::GoalSampleableRegion.
#ifndef OMPL_MIN_BASE_GOALS_GOALSAMPLEABLEREGION_H
#define OMPL_MIN_BASE_GOALS_GOALSAMPLEABLEREGION_H
#include "ompl/base/Goal.h"
namespace ompl {
namespace base {
class GoalSampleableRegion : public GoalRegion {
public:
  explicit GoalSampleableRegion(const SpaceInformationPtr &si) : GoalRegion(si) {
    type_ = GOAL_SAMPLEABLE_REGION;
  }
  virtual void sampleGoal(State *st) const = 0;
  virtual unsigned int maxSampleCount() const = 0;
  bool canSample() const { return maxSampleCount() > 0; }
  virtual bool couldSample() const { return canSample(); }
};
} // namespace base
} // namespace ompl
#endif

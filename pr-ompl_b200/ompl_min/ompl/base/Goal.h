// ompl_min: ompl::base::Goal and GoalRegion.
#ifndef OMPL_MIN_BASE_GOAL_H
#define OMPL_MIN_BASE_GOAL_H
#include <limits>
#include "ompl/base/SpaceInformation.h"
namespace ompl {
namespace base {
enum GoalType {
  GOAL_ANY = 1,
  GOAL_REGION = GOAL_ANY + 2,
  GOAL_SAMPLEABLE_REGION = GOAL_REGION + 4,
  GOAL_STATE = GOAL_SAMPLEABLE_REGION + 8,
  GOAL_STATES = GOAL_SAMPLEABLE_REGION + 16,
  GOAL_LAZY_SAMPLES = GOAL_STATES + 32
};

class Goal {
public:
  explicit Goal(const SpaceInformationPtr &si) : type_(GOAL_ANY), si_(si) {}
  virtual ~Goal() {}
  Goal(const Goal &) = delete;
  Goal &operator=(const Goal &) = delete;
  template <class T> T *as() { return static_cast<T *>(this); }
  template <class T> const T *as() const { return static_cast<const T *>(this); }
  GoalType getType() const { return type_; }
  bool hasType(GoalType type) const { return (type_ & type) == type; }
  const SpaceInformationPtr &getSpaceInformation() const { return si_; }
  virtual bool isSatisfied(const State *st) const = 0;
  virtual bool isSatisfied(const State *st, double *distance) const {
    if (distance)
      *distance = std::numeric_limits<double>::max();
    return isSatisfied(st);
  }
  virtual bool isStartGoalPairValid(const State * /*start*/, const State * /*goal*/) const {
    return true;
  }

protected:
  GoalType type_;
  SpaceInformationPtr si_;
};
typedef std::shared_ptr<Goal> GoalPtr;

class GoalRegion : public Goal {
public:
  explicit GoalRegion(const SpaceInformationPtr &si)
      : Goal(si), threshold_(std::numeric_limits<double>::epsilon()) {
    type_ = GOAL_REGION;
  }
  bool isSatisfied(const State *st) const override { return distanceGoal(st) <= threshold_; }
  bool isSatisfied(const State *st, double *distance) const override {
    double d2g = distanceGoal(st);
    if (distance)
      *distance = d2g;
    return d2g <= threshold_;
  }
  virtual double distanceGoal(const State *st) const = 0;
  void setThreshold(double threshold) { threshold_ = threshold; }
  double getThreshold() const { return threshold_; }

protected:
  double threshold_;
};
} // namespace base
} // namespace ompl
#endif

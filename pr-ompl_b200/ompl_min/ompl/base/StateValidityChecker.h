// ompl_min: ompl::base::StateValidityChecker.
#ifndef OMPL_MIN_BASE_STATEVALIDITYCHECKER_H
#define OMPL_MIN_BASE_STATEVALIDITYCHECKER_H
#include <functional>
#include <memory>
#include "ompl/base/State.h"
namespace ompl {
namespace base {
class SpaceInformation;
typedef std::shared_ptr<SpaceInformation> SpaceInformationPtr;

class StateValidityChecker {
public:
  explicit StateValidityChecker(SpaceInformation *si) : si_(si) {}
  explicit StateValidityChecker(const SpaceInformationPtr &si) : si_(si.get()) {}
  virtual ~StateValidityChecker() {}
  virtual bool isValid(const State *state) const = 0;
  virtual bool isValid(const State *state, double &dist) const {
    dist = clearance(state);
    return isValid(state);
  }
  virtual double clearance(const State *) const { return 0.0; }

protected:
  SpaceInformation *si_;
};
typedef std::shared_ptr<StateValidityChecker> StateValidityCheckerPtr;
typedef std::function<bool(const State *)> StateValidityCheckerFn;

class AllValidStateValidityChecker : public StateValidityChecker {
public:
  using StateValidityChecker::StateValidityChecker;
  bool isValid(const State *) const override { return true; }
};
} // namespace base
} // namespace ompl
#endif

// ompl_min: ompl::base::SpaceInformation.  Forwards distance/alloc/copy to the
// space, isValid to the StateValidityChecker and checkMotion to the installed
// MotionValidator.  Unlike upstream, setup() does NOT silently create a CPU
// DiscreteMotionValidator: a validator must be installed explicitly (the GPU
// one from plugin/, or the oracle's CPU one in tests).
#ifndef OMPL_MIN_BASE_SPACEINFORMATION_H
#define OMPL_MIN_BASE_SPACEINFORMATION_H
#include "ompl/base/MotionValidator.h"
#include "ompl/base/StateSpace.h"
#include "ompl/base/StateValidityChecker.h"
#include "ompl/util/Console.h"
namespace ompl {
namespace base {
class SpaceInformation {
public:
  explicit SpaceInformation(const StateSpacePtr &space) : stateSpace_(space), setup_(false) {
    if (!stateSpace_)
      throw Exception("Invalid space definition");
  }
  virtual ~SpaceInformation() {}
  SpaceInformation(const SpaceInformation &) = delete;
  SpaceInformation &operator=(const SpaceInformation &) = delete;

  const StateSpacePtr &getStateSpace() const { return stateSpace_; }
  unsigned int getStateDimension() const { return stateSpace_->getDimension(); }
  double getMaximumExtent() const { return stateSpace_->getMaximumExtent(); }

  bool isValid(const State *state) const { return stateValidityChecker_->isValid(state); }
  bool satisfiesBounds(const State *state) const { return stateSpace_->satisfiesBounds(state); }
  bool equalStates(const State *a, const State *b) const { return stateSpace_->equalStates(a, b); }
  double distance(const State *a, const State *b) const { return stateSpace_->distance(a, b); }
  void enforceBounds(State *state) const { stateSpace_->enforceBounds(state); }

  void setStateValidityChecker(const StateValidityCheckerPtr &svc) {
    stateValidityChecker_ = svc;
    setup_ = false;
  }
  void setStateValidityChecker(const StateValidityCheckerFn &svc) {
    class FnChecker : public StateValidityChecker {
    public:
      FnChecker(SpaceInformation *si, const StateValidityCheckerFn &fn)
          : StateValidityChecker(si), fn_(fn) {}
      bool isValid(const State *s) const override { return fn_(s); }
      StateValidityCheckerFn fn_;
    };
    if (!svc)
      throw Exception("Invalid function definition for state validity checking");
    setStateValidityChecker(StateValidityCheckerPtr(new FnChecker(this, svc)));
  }
  const StateValidityCheckerPtr &getStateValidityChecker() const { return stateValidityChecker_; }

  void setMotionValidator(const MotionValidatorPtr &mv) {
    motionValidator_ = mv;
    setup_ = false;
  }
  const MotionValidatorPtr &getMotionValidator() const { return motionValidator_; }

  void setStateValidityCheckingResolution(double resolution) {
    stateSpace_->setLongestValidSegmentFraction(resolution);
    setup_ = false;
  }
  double getStateValidityCheckingResolution() const {
    return stateSpace_->getLongestValidSegmentFraction();
  }

  bool checkMotion(const State *s1, const State *s2) const {
    return motionValidator_->checkMotion(s1, s2);
  }
  bool checkMotion(const State *s1, const State *s2, std::pair<State *, double> &lastValid) const {
    return motionValidator_->checkMotion(s1, s2, lastValid);
  }

  State *allocState() const { return stateSpace_->allocState(); }
  void allocStates(std::vector<State *> &states) const {
    for (auto &s : states)
      s = stateSpace_->allocState();
  }
  void freeState(State *state) const { stateSpace_->freeState(state); }
  void freeStates(std::vector<State *> &states) const {
    for (auto &s : states)
      stateSpace_->freeState(s);
  }
  void copyState(State *destination, const State *source) const {
    stateSpace_->copyState(destination, source);
  }
  State *cloneState(const State *source) const {
    State *copy = stateSpace_->allocState();
    stateSpace_->copyState(copy, source);
    return copy;
  }
  StateSamplerPtr allocStateSampler() const { return stateSpace_->allocStateSampler(); }

  virtual void setup() {
    if (!stateValidityChecker_) {
      stateValidityChecker_.reset(new AllValidStateValidityChecker(this));
      OMPL_WARN("State validity checker not set! No collision checking is performed");
    }
    if (!motionValidator_)
      throw Exception("SpaceInformation",
                      "no MotionValidator installed: ompl_min ships none; install "
                      "pr_gpu::GpuMotionValidator (plugin/) or, in tests, the oracle's "
                      "DiscreteMotionValidator");
    stateSpace_->setup();
    if (stateSpace_->getDimension() <= 0)
      throw Exception("The dimension of the state space we plan in must be > 0");
    setup_ = true;
  }
  bool isSetup() const { return setup_; }

protected:
  StateSpacePtr stateSpace_;
  StateValidityCheckerPtr stateValidityChecker_;
  MotionValidatorPtr motionValidator_;
  bool setup_;
};
} // namespace base
} // namespace ompl
#endif

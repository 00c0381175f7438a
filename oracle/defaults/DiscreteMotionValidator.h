// ORACLE / TEST INFRASTRUCTURE ONLY -- never included by the product path.
// CPU restatement of OMPL's DiscreteMotionValidator::checkMotion (third-party
// dependency of the reference, unpinned; semantics UPSTREAM-RECALLED, SURVEY.md
// Appendix B.3).  This is what si_->checkMotion reaches in the reference
// (pr_ompl/src/RRTConnect.cpp:198):  s1 assumed valid; s2 tested first; then
// the nd-1 interior points j/nd in FIFO-bisection order with early exit.
// Parity unpinned upstream: the reference has no tests.
#ifndef ORACLE_DISCRETEMOTIONVALIDATOR_H
#define ORACLE_DISCRETEMOTIONVALIDATOR_H
#include <queue>
#include "ompl/base/MotionValidator.h"
#include "ompl/base/SpaceInformation.h"
namespace ompl {
namespace base {
class DiscreteMotionValidator : public MotionValidator {
public:
  explicit DiscreteMotionValidator(SpaceInformation *si) : MotionValidator(si) { defaultSettings(); }
  explicit DiscreteMotionValidator(const SpaceInformationPtr &si) : MotionValidator(si) {
    defaultSettings();
  }
  ~DiscreteMotionValidator() override {}

  bool checkMotion(const State *s1, const State *s2) const override {
    /* assume motion starts in a valid configuration so s1 is valid */
    if (!si_->isValid(s2)) {
      invalid_++;
      return false;
    }
    bool result = true;
    int nd = (int)stateSpace_->validSegmentCount(s1, s2);

    /* initialize the queue of test positions */
    std::queue<std::pair<int, int>> pos;
    if (nd >= 2) {
      pos.push(std::make_pair(1, nd - 1));
      /* temporary storage for the checked state */
      State *test = si_->allocState();
      /* repeatedly subdivide the path segment in the middle (and check the middle) */
      while (!pos.empty()) {
        std::pair<int, int> x = pos.front();
        int mid = (x.first + x.second) / 2;
        stateSpace_->interpolate(s1, s2, (double)mid / (double)nd, test);
        if (!si_->isValid(test)) {
          result = false;
          break;
        }
        pos.pop();
        if (x.first < mid)
          pos.push(std::make_pair(x.first, mid - 1));
        if (x.second > mid)
          pos.push(std::make_pair(mid + 1, x.second));
      }
      si_->freeState(test);
    }
    if (result)
      valid_++;
    else
      invalid_++;
    return result;
  }

  bool checkMotion(const State *s1, const State *s2,
                   std::pair<State *, double> &lastValid) const override {
    /* assume motion starts in a valid configuration so s1 is valid */
    bool result = true;
    int nd = (int)stateSpace_->validSegmentCount(s1, s2);
    if (nd > 1) {
      /* temporary storage for the checked state */
      State *test = si_->allocState();
      for (int j = 1; j < nd; ++j) {
        stateSpace_->interpolate(s1, s2, (double)j / (double)nd, test);
        if (!si_->isValid(test)) {
          lastValid.second = (double)(j - 1) / (double)nd;
          if (lastValid.first)
            stateSpace_->interpolate(s1, s2, lastValid.second, lastValid.first);
          result = false;
          break;
        }
      }
      si_->freeState(test);
    }
    if (result)
      if (!si_->isValid(s2)) {
        lastValid.second = (double)(nd - 1) / (double)nd;
        if (lastValid.first)
          stateSpace_->interpolate(s1, s2, lastValid.second, lastValid.first);
        result = false;
      }
    if (result)
      valid_++;
    else
      invalid_++;
    return result;
  }

private:
  void defaultSettings() {
    stateSpace_ = si_->getStateSpace().get();
    if (!stateSpace_)
      throw Exception("No state space for motion validator");
  }
  StateSpace *stateSpace_;
};
} // namespace base
} // namespace ompl
#endif

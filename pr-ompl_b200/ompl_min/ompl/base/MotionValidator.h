// ompl_min: ompl::base::MotionValidator (abstract).  No concrete validator is
// shipped here on purpose: the product installs plugin/GpuMotionValidator.h;
// the CPU DiscreteMotionValidator restatement lives under /oracle.
#ifndef OMPL_MIN_BASE_MOTIONVALIDATOR_H
#define OMPL_MIN_BASE_MOTIONVALIDATOR_H
#include <memory>
#include <utility>
#include "ompl/base/State.h"
namespace ompl {
namespace base {
class SpaceInformation;
typedef std::shared_ptr<SpaceInformation> SpaceInformationPtr;

class MotionValidator {
public:
  explicit MotionValidator(SpaceInformation *si) : si_(si), valid_(0), invalid_(0) {}
  explicit MotionValidator(const SpaceInformationPtr &si) : si_(si.get()), valid_(0), invalid_(0) {}
  virtual ~MotionValidator() {}
  /// s1 is assumed valid.
  virtual bool checkMotion(const State *s1, const State *s2) const = 0;
  virtual bool checkMotion(const State *s1, const State *s2,
                           std::pair<State *, double> &lastValid) const = 0;
  unsigned int getValidMotionCount() const { return valid_; }
  unsigned int getInvalidMotionCount() const { return invalid_; }
  unsigned int getCheckedMotionCount() const { return valid_ + invalid_; }
  void resetMotionCounter() { valid_ = invalid_ = 0; }

protected:
  SpaceInformation *si_;
  mutable unsigned int valid_;
  mutable unsigned int invalid_;
};
typedef std::shared_ptr<MotionValidator> MotionValidatorPtr;
} // namespace base
} // namespace ompl
#endif

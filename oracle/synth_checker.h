// ORACLE / TEST INFRASTRUCTURE ONLY.
// ompl::base::StateValidityChecker over the synthetic world: plays the role of the
// user-supplied validity callback the reference calls at pr_ompl/src/RRTConnect.cpp:198
// and frechet/src/NNFrechet.cpp:665,675.
#ifndef ORACLE_SYNTH_CHECKER_H
#define ORACLE_SYNTH_CHECKER_H
#include <limits>
#include "ompl/base/SpaceInformation.h"
#include "ompl/base/spaces/RealVectorStateSpace.h"
#include "synth_world.h"
namespace oracle {
namespace ob = ompl::base;
class SynthChecker : public ob::StateValidityChecker {
public:
  SynthChecker(ob::SpaceInformation *si, const SynthWorld *w) : ob::StateValidityChecker(si), w_(w) {}
  bool isValid(const ob::State *s) const override {
    double c = w_->clearance(s->as<ob::RealVectorStateSpace::StateType>()->values);
    if (track_) {
      if (c < minClear_)
        minClear_ = c;
      ++count_;
    }
    return c > 0.0;
  }
  double clearance(const ob::State *s) const override {
    return w_->clearance(s->as<ob::RealVectorStateSpace::StateType>()->values);
  }
  const SynthWorld *w_;
  mutable bool track_ = false;
  mutable double minClear_ = std::numeric_limits<double>::infinity();
  mutable unsigned count_ = 0;
};
} // namespace oracle
#endif

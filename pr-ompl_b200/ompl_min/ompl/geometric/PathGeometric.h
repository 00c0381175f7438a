// ompl_min: ompl::geometric::PathGeometric (owns copies of appended states).
#ifndef OMPL_MIN_GEOMETRIC_PATHGEOMETRIC_H
#define OMPL_MIN_GEOMETRIC_PATHGEOMETRIC_H
#include <vector>
#include "ompl/base/Path.h"
namespace ompl {
namespace geometric {
class PathGeometric : public base::Path {
public:
  explicit PathGeometric(const base::SpaceInformationPtr &si) : base::Path(si) {}
  ~PathGeometric() override { clear(); }
  void append(const base::State *state) { states_.push_back(si_->cloneState(state)); }
  std::vector<base::State *> &getStates() { return states_; }
  const std::vector<base::State *> &getStates() const { return states_; }
  base::State *getState(unsigned int index) { return states_[index]; }
  const base::State *getState(unsigned int index) const { return states_[index]; }
  std::size_t getStateCount() const { return states_.size(); }
  void clear() {
    for (auto *s : states_)
      si_->freeState(s);
    states_.clear();
  }
  double length() const override {
    double L = 0.0;
    for (std::size_t i = 1; i < states_.size(); ++i)
      L += si_->distance(states_[i - 1], states_[i]);
    return L;
  }
  bool check() const override {
    if (states_.empty())
      return true;
    if (!si_->isValid(states_[0]))
      return false;
    for (std::size_t i = 1; i < states_.size(); ++i)
      if (!si_->checkMotion(states_[i - 1], states_[i]))
        return false;
    return true;
  }

protected:
  std::vector<base::State *> states_;
};
} // namespace geometric
} // namespace ompl
#endif

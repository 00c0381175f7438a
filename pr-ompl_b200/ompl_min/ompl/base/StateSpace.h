// ompl_min: ompl::base::StateSpace / StateSampler abstract interfaces.
// Semantics follow upstream OMPL 1.x as recalled in SURVEY.md Appendix B.
#ifndef OMPL_MIN_BASE_STATESPACE_H
#define OMPL_MIN_BASE_STATESPACE_H
#include <cmath>
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include "ompl/base/State.h"
#include "ompl/util/Exception.h"
#include "ompl/util/RandomNumbers.h"
namespace ompl {
namespace base {
class StateSpace;
typedef std::shared_ptr<StateSpace> StateSpacePtr;

class StateSampler {
public:
  explicit StateSampler(const StateSpace *space) : space_(space) {}
  virtual ~StateSampler() {}
  virtual void sampleUniform(State *state) = 0;
  virtual void sampleUniformNear(State *state, const State *near, double distance) = 0;
  virtual void sampleGaussian(State *state, const State *mean, double stdDev) = 0;

protected:
  const StateSpace *space_;
  RNG rng_;
};
typedef std::shared_ptr<StateSampler> StateSamplerPtr;
typedef std::function<StateSamplerPtr(const StateSpace *)> StateSamplerAllocator;

/// upstream ompl/base/StateSpaceTypes.h (the values the planners' drop-in classes test against)
enum StateSpaceType {
  STATE_SPACE_UNKNOWN = 0,
  STATE_SPACE_REAL_VECTOR = 1,
  STATE_SPACE_SO2 = 2,
  STATE_SPACE_SO3 = 3,
  STATE_SPACE_SE2 = 4,
  STATE_SPACE_SE3 = 5,
  STATE_SPACE_TIME = 6,
  STATE_SPACE_DISCRETE = 7,
  STATE_SPACE_TYPE_COUNT
};

class StateSpace {
public:
  typedef State StateType;
  StateSpace() : name_("Space"), type_(STATE_SPACE_UNKNOWN) {}
  int getType() const { return type_; }
  virtual ~StateSpace() {}
  StateSpace(const StateSpace &) = delete;
  StateSpace &operator=(const StateSpace &) = delete;

  template <class T> T *as() { return static_cast<T *>(this); }
  template <class T> const T *as() const { return static_cast<const T *>(this); }

  const std::string &getName() const { return name_; }
  void setName(const std::string &name) { name_ = name; }

  virtual bool isMetricSpace() const { return true; }
  virtual unsigned int getDimension() const = 0;
  virtual double getMaximumExtent() const = 0;
  virtual void enforceBounds(State *state) const = 0;
  virtual bool satisfiesBounds(const State *state) const = 0;
  virtual void copyState(State *destination, const State *source) const = 0;
  virtual double distance(const State *state1, const State *state2) const = 0;
  virtual bool equalStates(const State *state1, const State *state2) const = 0;
  virtual void interpolate(const State *from, const State *to, double t, State *state) const = 0;
  virtual StateSamplerPtr allocDefaultStateSampler() const = 0;
  virtual State *allocState() const = 0;
  virtual void freeState(State *state) const = 0;
  /// address of the index-th real value of a state, or NULL (upstream name)
  virtual double *getValueAddressAtIndex(State *state, unsigned int index) const = 0;
  const double *getValueAddressAtIndex(const State *state, unsigned int index) const {
    return const_cast<StateSpace *>(this)->getValueAddressAtIndex(const_cast<State *>(state), index);
  }
  virtual void copyToReals(std::vector<double> &reals, const State *source) const {
    reals.resize(getDimension());
    for (unsigned int i = 0; i < getDimension(); ++i)
      reals[i] = *getValueAddressAtIndex(source, i);
  }
  virtual void copyFromReals(State *destination, const std::vector<double> &reals) const {
    for (unsigned int i = 0; i < getDimension(); ++i)
      *getValueAddressAtIndex(destination, i) = reals[i];
  }

  virtual StateSamplerPtr allocStateSampler() const {
    return ssa_ ? ssa_(this) : allocDefaultStateSampler();
  }
  void setStateSamplerAllocator(const StateSamplerAllocator &ssa) { ssa_ = ssa; }
  void clearStateSamplerAllocator() { ssa_ = StateSamplerAllocator(); }

  // --- discretisation of motions (used by DiscreteMotionValidator upstream) ---
  virtual void setLongestValidSegmentFraction(double segmentFraction) {
    if (segmentFraction < 1e-300 || segmentFraction > 1.0 - 1e-16)
      throw Exception("The fraction of the extent must be larger than 0 and less than 1");
    longestValidSegmentFraction_ = segmentFraction;
    longestValidSegment_ = getMaximumExtent() * longestValidSegmentFraction_;
  }
  virtual double getLongestValidSegmentFraction() const { return longestValidSegmentFraction_; }
  void setValidSegmentCountFactor(unsigned int factor) {
    if (factor < 1)
      throw Exception("The multiplicative factor for the valid segment count must be >= 1");
    longestValidSegmentCountFactor_ = factor;
  }
  unsigned int getValidSegmentCountFactor() const { return longestValidSegmentCountFactor_; }
  double getLongestValidSegmentLength() const { return longestValidSegment_; }
  /// nd = factor * (unsigned)ceil(distance / longestValidSegment)   (Appendix B.3)
  virtual unsigned int validSegmentCount(const State *state1, const State *state2) const {
    return longestValidSegmentCountFactor_ *
           (unsigned int)std::ceil(distance(state1, state2) / longestValidSegment_);
  }
  virtual void setup() { longestValidSegment_ = getMaximumExtent() * longestValidSegmentFraction_; }

protected:
  std::string name_;
  int type_;
  StateSamplerAllocator ssa_;
  double longestValidSegmentFraction_ = 0.01;
  double longestValidSegment_ = 0.0;
  unsigned int longestValidSegmentCountFactor_ = 1;
};
} // namespace base
} // namespace ompl
#endif

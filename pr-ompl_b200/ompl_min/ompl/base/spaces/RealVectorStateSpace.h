// ompl_min: ompl::base::RealVectorStateSpace.  distance / interpolate restate
// upstream OMPL exactly: sequential sum of
// squared differences then sqrt; from + (to-from)*t per dimension.  Compile
// host code with -ffp-contract=off so no FMA is contracted into these loops.
#ifndef OMPL_MIN_BASE_SPACES_REALVECTORSTATESPACE_H
#define OMPL_MIN_BASE_SPACES_REALVECTORSTATESPACE_H
#include <cstring>
#include <limits>
#include "ompl/base/StateSpace.h"
#include "ompl/base/spaces/RealVectorBounds.h"
namespace ompl {
namespace base {
class RealVectorStateSampler : public StateSampler {
public:
  explicit RealVectorStateSampler(const StateSpace *space) : StateSampler(space) {}
  void sampleUniform(State *state) override;
  void sampleUniformNear(State *state, const State *near, double distance) override;
  void sampleGaussian(State *state, const State *mean, double stdDev) override;
};

class RealVectorStateSpace : public StateSpace {
public:
  class StateType : public State {
  public:
    StateType() : values(nullptr) {}
    double operator[](unsigned int i) const { return values[i]; }
    double &operator[](unsigned int i) { return values[i]; }
    double *values;
  };

  explicit RealVectorStateSpace(unsigned int dim = 0)
      : dimension_(dim), bounds_(dim), stateBytes_(dim * sizeof(double)) {
    setName("RealVector" + getName());
    type_ = STATE_SPACE_REAL_VECTOR;
  }
  void addDimension(double minBound = 0.0, double maxBound = 0.0) {
    dimension_++;
    stateBytes_ = dimension_ * sizeof(double);
    bounds_.low.push_back(minBound);
    bounds_.high.push_back(maxBound);
  }
  void setBounds(const RealVectorBounds &bounds) {
    bounds.check();
    if (bounds.low.size() != dimension_)
      throw Exception("Bounds do not match dimension of state space");
    bounds_ = bounds;
  }
  void setBounds(double low, double high) {
    RealVectorBounds b(dimension_);
    b.setLow(low);
    b.setHigh(high);
    setBounds(b);
  }
  const RealVectorBounds &getBounds() const { return bounds_; }

  unsigned int getDimension() const override { return dimension_; }
  double getMaximumExtent() const override {
    double e = 0.0;
    for (unsigned int i = 0; i < dimension_; ++i) {
      double d = bounds_.high[i] - bounds_.low[i];
      e += d * d;
    }
    return std::sqrt(e);
  }
  void enforceBounds(State *state) const override {
    StateType *s = static_cast<StateType *>(state);
    for (unsigned int i = 0; i < dimension_; ++i) {
      if (s->values[i] > bounds_.high[i])
        s->values[i] = bounds_.high[i];
      else if (s->values[i] < bounds_.low[i])
        s->values[i] = bounds_.low[i];
    }
  }
  bool satisfiesBounds(const State *state) const override {
    const StateType *s = static_cast<const StateType *>(state);
    for (unsigned int i = 0; i < dimension_; ++i)
      if (s->values[i] - std::numeric_limits<double>::epsilon() > bounds_.high[i] ||
          s->values[i] + std::numeric_limits<double>::epsilon() < bounds_.low[i])
        return false;
    return true;
  }
  void copyState(State *destination, const State *source) const override {
    std::memcpy(static_cast<StateType *>(destination)->values,
                static_cast<const StateType *>(source)->values, stateBytes_);
  }
  double distance(const State *state1, const State *state2) const override {
    double dist = 0.0;
    const double *s1 = static_cast<const StateType *>(state1)->values;
    const double *s2 = static_cast<const StateType *>(state2)->values;
    for (unsigned int i = 0; i < dimension_; ++i) {
      double diff = (*s1++) - (*s2++);
      dist += diff * diff;
    }
    return std::sqrt(dist);
  }
  bool equalStates(const State *state1, const State *state2) const override {
    const double *s1 = static_cast<const StateType *>(state1)->values;
    const double *s2 = static_cast<const StateType *>(state2)->values;
    for (unsigned int i = 0; i < dimension_; ++i) {
      double diff = (*s1++) - (*s2++);
      if (std::fabs(diff) > std::numeric_limits<double>::epsilon() * 2.0)
        return false;
    }
    return true;
  }
  void interpolate(const State *from, const State *to, double t, State *state) const override {
    const StateType *rfrom = static_cast<const StateType *>(from);
    const StateType *rto = static_cast<const StateType *>(to);
    const StateType *rstate = static_cast<StateType *>(state);
    for (unsigned int i = 0; i < dimension_; ++i)
      rstate->values[i] = rfrom->values[i] + (rto->values[i] - rfrom->values[i]) * t;
  }
  StateSamplerPtr allocDefaultStateSampler() const override {
    return StateSamplerPtr(new RealVectorStateSampler(this));
  }
  State *allocState() const override {
    StateType *rstate = new StateType();
    rstate->values = new double[dimension_];
    return rstate;
  }
  void freeState(State *state) const override {
    StateType *rstate = static_cast<StateType *>(state);
    delete[] rstate->values;
    delete rstate;
  }
  double *getValueAddressAtIndex(State *state, unsigned int index) const override {
    return index < dimension_ ? static_cast<StateType *>(state)->values + index : nullptr;
  }
  void setup() override {
    bounds_.check();
    StateSpace::setup();
  }

protected:
  unsigned int dimension_;
  RealVectorBounds bounds_;
  std::size_t stateBytes_;
};

inline void RealVectorStateSampler::sampleUniform(State *state) {
  const RealVectorStateSpace *sp = static_cast<const RealVectorStateSpace *>(space_);
  const RealVectorBounds &b = sp->getBounds();
  RealVectorStateSpace::StateType *s = static_cast<RealVectorStateSpace::StateType *>(state);
  for (unsigned int i = 0; i < sp->getDimension(); ++i)
    s->values[i] = rng_.uniformReal(b.low[i], b.high[i]);
}
inline void RealVectorStateSampler::sampleUniformNear(State *state, const State *near,
                                                      double distance) {
  const RealVectorStateSpace *sp = static_cast<const RealVectorStateSpace *>(space_);
  const RealVectorBounds &b = sp->getBounds();
  RealVectorStateSpace::StateType *s = static_cast<RealVectorStateSpace::StateType *>(state);
  const RealVectorStateSpace::StateType *n =
      static_cast<const RealVectorStateSpace::StateType *>(near);
  for (unsigned int i = 0; i < sp->getDimension(); ++i)
    s->values[i] = rng_.uniformReal(std::max(b.low[i], n->values[i] - distance),
                                    std::min(b.high[i], n->values[i] + distance));
}
inline void RealVectorStateSampler::sampleGaussian(State *state, const State *mean, double stdDev) {
  const RealVectorStateSpace *sp = static_cast<const RealVectorStateSpace *>(space_);
  const RealVectorBounds &b = sp->getBounds();
  RealVectorStateSpace::StateType *s = static_cast<RealVectorStateSpace::StateType *>(state);
  const RealVectorStateSpace::StateType *m =
      static_cast<const RealVectorStateSpace::StateType *>(mean);
  for (unsigned int i = 0; i < sp->getDimension(); ++i) {
    double v = rng_.gaussian(m->values[i], stdDev);
    if (v < b.low[i])
      v = b.low[i];
    else if (v > b.high[i])
      v = b.high[i];
    s->values[i] = v;
  }
}
} // namespace base
} // namespace ompl
#endif

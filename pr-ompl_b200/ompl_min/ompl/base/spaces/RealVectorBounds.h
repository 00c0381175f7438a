// ompl_min: ompl::base::RealVectorBounds.
#ifndef OMPL_MIN_BASE_SPACES_REALVECTORBOUNDS_H
#define OMPL_MIN_BASE_SPACES_REALVECTORBOUNDS_H
#include <vector>
#include "ompl/util/Exception.h"
namespace ompl {
namespace base {
class RealVectorBounds {
public:
  explicit RealVectorBounds(unsigned int dim) { resize(dim); }
  void setLow(double value) { low.assign(low.size(), value); }
  void setHigh(double value) { high.assign(high.size(), value); }
  void setLow(unsigned int index, double value) { low.at(index) = value; }
  void setHigh(unsigned int index, double value) { high.at(index) = value; }
  void resize(std::size_t size) {
    low.resize(size, 0.0);
    high.resize(size, 0.0);
  }
  std::vector<double> getDifference() const {
    std::vector<double> d(low.size());
    for (std::size_t i = 0; i < low.size(); ++i)
      d[i] = high[i] - low[i];
    return d;
  }
  void check() const {
    if (low.size() != high.size())
      throw Exception("Lower and upper bounds are not of same dimension");
    for (std::size_t i = 0; i < low.size(); ++i)
      if (low[i] > high[i])
        throw Exception("Bounds for real vector space seem to be incorrect (lower bound must be "
                        "stricly less than upper bound).");
  }
  std::vector<double> low;
  std::vector<double> high;
};
} // namespace base
} // namespace ompl
#endif

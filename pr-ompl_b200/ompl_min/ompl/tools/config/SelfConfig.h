// ompl_min: ompl::tools::SelfConfig (configurePlannerRange, getDefaultNearestNeighbors).
// The default NN structure is supplied by the build: define
// OMPL_MIN_DEFAULT_NN_HEADER to a header that provides
//   template <typename T> ompl::NearestNeighbors<T>* ompl_min_default_nn(const ompl::base::StateSpacePtr&);
// The product build points it at plugin/GpuNearestNeighbors.h; the oracle build
// at oracle/defaults/NearestNeighborsLinear.h.  Without it the call throws.
#ifndef OMPL_MIN_TOOLS_CONFIG_SELFCONFIG_H
#define OMPL_MIN_TOOLS_CONFIG_SELFCONFIG_H
#include <limits>
#include <string>
#include "ompl/base/SpaceInformation.h"
#include "ompl/datastructures/NearestNeighbors.h"
#ifdef OMPL_MIN_DEFAULT_NN_HEADER
#include OMPL_MIN_DEFAULT_NN_HEADER
#endif
namespace ompl {
namespace tools {
class SelfConfig {
public:
  SelfConfig(const base::SpaceInformationPtr &si, const std::string &context = std::string())
      : si_(si), context_(context) {}
  /// range < eps  ->  0.2 * maximum extent
  void configurePlannerRange(double &range) {
    if (range < std::numeric_limits<double>::epsilon()) {
      range = si_->getMaximumExtent() * 0.2;
      OMPL_DEBUG("%sPlanner range detected to be %lf", context_.empty() ? "" : (context_ + ": ").c_str(),
                 range);
    }
  }
  void configureValidStateSamplingAttempts(unsigned int &attempts) {
    if (attempts == 0)
      attempts = 100;
  }
  template <typename _T>
  static NearestNeighbors<_T> *getDefaultNearestNeighbors(const base::StateSpacePtr &space) {
#ifdef OMPL_MIN_DEFAULT_NN_HEADER
    return ompl_min_default_nn<_T>(space);
#else
    (void)space;
    throw Exception("SelfConfig", "no default NearestNeighbors compiled in (OMPL_MIN_DEFAULT_NN_HEADER)");
#endif
  }

private:
  base::SpaceInformationPtr si_;
  std::string context_;
};
} // namespace tools
} // namespace ompl
#endif

// pr_gpu: error plumbing between the prgpu C ABI and OMPL exceptions.
#ifndef PR_GPU_ERROR_H
#define PR_GPU_ERROR_H
#include <string>
#include <ompl/base/StateSpace.h>
#include <ompl/util/Exception.h>
#include "prgpu.h"
namespace pr_gpu {
/// The device classes read a state as `dim` contiguous doubles and measure Euclidean distance on them:
/// that is RealVectorStateSpace and nothing else (SO2 wrap-around, compound spaces: not representable).
/// For any other space keep OMPL's own NearestNeighbors / validators.
inline void requireRealVectorSpace(const ompl::base::StateSpace *space, const char *who) {
  if (!space || space->getType() != ompl::base::STATE_SPACE_REAL_VECTOR)
    throw ompl::Exception(std::string(who), "the GPU path supports ompl::base::RealVectorStateSpace only (space '" +
                                                (space ? space->getName() : std::string("null")) + "')");
}
inline void check(int rc, const char *what) {
  if (rc != PRGPU_OK)
    throw ompl::Exception(std::string(what), std::string(prgpu_last_error()) + " (prgpu code " +
                                                 std::to_string(rc) + "; there is no CPU fallback)");
}
} // namespace pr_gpu
#endif

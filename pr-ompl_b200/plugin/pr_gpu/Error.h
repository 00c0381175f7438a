// pr_gpu: error plumbing between the prgpu C ABI and OMPL exceptions.
#ifndef PR_GPU_ERROR_H
#define PR_GPU_ERROR_H
#include <string>
#include <ompl/util/Exception.h>
#include "prgpu.h"
namespace pr_gpu {
inline void check(int rc, const char *what) {
  if (rc != PRGPU_OK)
    throw ompl::Exception(std::string(what), std::string(prgpu_last_error()) + " (prgpu code " +
                                                 std::to_string(rc) + "; there is no CPU fallback)");
}
} // namespace pr_gpu
#endif

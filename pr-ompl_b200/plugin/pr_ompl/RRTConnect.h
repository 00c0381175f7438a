// Same include path and class name as the reference (pr_ompl/include/pr_ompl/RRTConnect.h):
// code written against pr_ompl::RRTConnect compiles unchanged against the GPU-backed planner.
#ifndef PR_OMPL_RRT_CONNECT_H_
#define PR_OMPL_RRT_CONNECT_H_
#include "pr_gpu/RRTConnect.h"
namespace pr_ompl {
typedef pr_gpu::RRTConnect RRTConnect;
}
#endif
